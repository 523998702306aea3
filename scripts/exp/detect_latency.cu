// Experiment: how long after one warp stores a word in shared memory does a warp that spins on it see it?
// Warp 0 publishes {count, clock} every PERIOD cycles; NW waiter warps (on the other schedulers) take turns: waiter w waits
// for counts = w (mod NW) with the decoders' poll loop and records clock - stamp.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gpurun_out/detect_latency scripts/exp/detect_latency.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ void ld64v(uint32_t a, uint32_t &x, uint32_t &y) { asm volatile("ld.volatile.shared.v2.u32 {%0, %1}, [%2];" : "=r"(x), "=r"(y) : "r"(a) : "memory"); }
__device__ __forceinline__ void st64v(uint32_t a, uint32_t x, uint32_t y) { asm volatile("st.volatile.shared.v2.u32 [%0], {%1, %2};" ::"r"(a), "r"(x), "r"(y) : "memory"); }
template <int MODE>
__global__ void k(int nw, int period, int n, unsigned long long *out) {
  __shared__ __align__(16) uint32_t ring[8][2];
  const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t base = (uint32_t)__cvta_generic_to_shared(&ring[0][0]);
  if (threadIdx.x < 16) (&ring[0][0])[threadIdx.x] = 0;
  __syncthreads();
  if (warp == 0) {
    uint32_t next = (uint32_t)clock() + 2000;
    for (int c = 1; c <= n; c++) {
      while ((int32_t)((uint32_t)clock() - next) < 0) { }
      st64v(base + (c & 7) * 8, (uint32_t)clock(), (uint32_t)c);
      next += period;
    }
    return;
  }
  if ((warp & 3) == 0) return;              // scheduler 0 is the publisher's
  const int w = warp - 1 - (warp >> 2);     // waiter index
  if (w >= nw) return;
  unsigned long long sum = 0, worst = 0, cnt = 0;
  for (int c = w + 1; c <= n; c += nw) {
    uint32_t stamp, t;
    const uint32_t a = base + (c & 7) * 8;
    ld64v(a, stamp, t);
    if (t != (uint32_t)c) {
      uint32_t spins = 0;
      do ld64v(a, stamp, t);
      while (t != (uint32_t)c && ++spins < (1u << 24));
    }
    const uint32_t d = (uint32_t)clock() - stamp;
    if (MODE == 1) { // busy work after each detection, like a model warp's lookup: ~100 dependent instructions
      uint32_t x = stamp;
      for (int i = 0; i < 100; i++) x = x * 1664525u + 1013904223u;
      if (x == 12345u) sum += 1;
    }
    sum += d; cnt++; if (d > worst) worst = d;
  }
  if (lane == 0) { out[w * 3] = sum; out[w * 3 + 1] = cnt; out[w * 3 + 2] = worst; }
}
int main() {
  unsigned long long *d, h[24];
  cudaMalloc(&d, sizeof h);
  for (int mode = 0; mode < 2; mode++)
    for (int nw : {1, 3, 8})
      for (int period : {150, 300, 1000}) {
        cudaMemset(d, 0, sizeof h);
        if (mode == 0) k<0><<<1, 352>>>(nw, period, 200000, d); else k<1><<<1, 352>>>(nw, period, 200000, d);
        cudaDeviceSynchronize();
        cudaMemcpy(h, d, sizeof h, cudaMemcpyDeviceToHost);
        printf("mode %d waiters %d period %4d:", mode, nw, period);
        for (int w = 0; w < nw; w++) printf("  w%d avg %.0f worst %llu", w, (double)h[w * 3] / (double)(h[w * 3 + 1] ? h[w * 3 + 1] : 1), h[w * 3 + 2]);
        printf("\n");
      }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
