// microbench.cu — cycles per coder step of a lone warp, for variants of the
// per-bit loop.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/microbench tools/microbench.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../weath3rb0i_b200/csrc/w3_core.cuh"

#define N_STEPS (1 << 18)
#define FULL 0xFFFFFFFFu

__device__ __forceinline__ uint32_t lds(uint32_t a) { uint32_t v; asm volatile("ld.volatile.shared.u32 %0, [%1];" : "=r"(v) : "r"(a) : "memory"); return v; }
__device__ __forceinline__ uint32_t lds_nv(uint32_t a) { uint32_t v; asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a)); return v; }
__device__ __forceinline__ void sts(uint32_t a, uint32_t v) { asm volatile("st.volatile.shared.u32 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }
__device__ __forceinline__ uint32_t bswap(uint32_t v) { return __byte_perm(v, 0, 0x0123); }

// ring of (p<<16)|bit entries, 2048 of them, text-like: p far from 1/2 most of the time
__device__ void fill_ring(uint32_t *ring, int n) {
  uint32_t s = 12345u + threadIdx.x;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    s = s * 1664525u + 1013904223u;
    uint32_t r = s >> 8;
    uint32_t p = (r & 7) == 0 ? 20000 + (r >> 4) % 25000 : ((r & 8) ? 60000 + (r >> 4) % 5000 : 500 + (r >> 4) % 5000);
    uint32_t bit = ((s >> 3) % 65536) < p ? 1u : 0u;
    ring[i] = (p << 16) | bit;
  }
}

template <int V>
__global__ void k_enc(uint32_t *out, long long *cycles, uint32_t *sink) {
  __shared__ uint32_t ring[2048];
  fill_ring(ring, 2048);
  __syncthreads();
  const uint32_t ra = (uint32_t)__cvta_generic_to_shared(ring);
  const uint32_t lane = threadIdx.x & 31;
  uint32_t x1 = 0, x2 = 0xFFFFFFFFu;
  w3_coder cw = w3_coder_new();
  sts(ra + 8192 - 16, 0xFFFFFFFFu);
  const uint32_t ones = lds(ra + 8192 - 16); /* opaque to ptxas */
  uint64_t acc = 0; uint32_t nb = 0, words = 0, rev = 0, trouble = 0;
  uint32_t *dst = out + blockIdx.x * (N_STEPS / 8 + 64);
  long long t0 = clock64();
  for (int base = 0; base < N_STEPS; base += 32) {
    const uint32_t ga = ra + (base & 2047) * 4;
    uint32_t v0 = lds(ga), v1 = lds(ga + 4);
#pragma unroll
    for (int j = 0; j < 32; j++) {
      uint32_t vj;
      if (V == 0) { vj = (0x6000u + ((base + j) & 7) * 0x1000u) << 16 | ((base + j) & 1); }
      else { vj = v0; v0 = v1; if (j + 2 < 32) v1 = lds(ga + (j + 2) * 4); }
      if (V == 7) { /* as the kernel's coder warp: state only, publish the bare bounds */
        uint32_t flip, x1p, x2p;
        (void)w3_coder_step<false>(cw, vj & 1u, vj & 0xFFFF0000u, flip, x1p, x2p, ones);
        asm volatile("st.volatile.shared.v2.u32 [%0], {%1, %2};" ::"r"(ra + 8192 - 8), "r"(x1p), "r"(x2p) : "memory");
        continue;
      }
      if (V >= 5) { /* width-carrying coder: one clz per step */
        uint32_t flip, x1p, x2p;
        const int n = w3_coder_step(cw, vj & 1u, vj & 0xFFFF0000u, flip, x1p, x2p);
        if (V == 6) { /* what the writer warp needs: bounds' common prefix, its length, the total */
          const int s6 = w3_clz32_nz((x1p ^ x2p) | 1u);
          sts(ra + 8192 - 8, x1p); sts(ra + 8192 - 4, (uint32_t)s6 | ((uint32_t)n << 8));
        }
        continue;
      }
      w3_ac_narrow(x1, x2, vj & 1u, vj & 0xFFFF0000u);
      const uint32_t pre = x1;
      int s, k;
      w3_ac_renorm(x1, x2, s, k);
      if (V >= 2) {
        trouble |= (rev != 0u && s != 0) ? 1u : 0u;
        acc = (acc << s) | __funnelshift_l(pre, 0u, (uint32_t)s);
        nb += (uint32_t)s;
        rev = s ? (uint32_t)k : rev + (uint32_t)k;
        if (V == 2) {
          const uint32_t full = nb >= 32u ? 1u : 0u;
          nb -= full << 5;
          const uint32_t word = bswap((uint32_t)(acc >> nb));
          asm volatile("{ .reg .pred p; setp.ne.u32 p, %0, 0; @p st.global.u32 [%1], %2; }" ::"r"(full & (lane == 0)), "l"(dst + words), "r"(word) : "memory");
          words += full;
        } else if (V == 3) {
          if (nb >= 32u) { nb -= 32u; if (lane == 0) dst[words] = bswap((uint32_t)(acc >> nb)); words++; }
        } else if (V == 4) {  // publish (pre, s, k) to smem for a writer warp
          sts(ra + 8192 - 8, pre); sts(ra + 8192 - 4, (uint32_t)s | ((uint32_t)k << 8));
        }
      }
    }
  }
  long long t1 = clock64();
  if (threadIdx.x == 0) { cycles[blockIdx.x] = t1 - t0; sink[blockIdx.x] = x1 ^ x2 ^ (uint32_t)acc ^ nb ^ words ^ trouble ^ rev ^ cw.x1 ^ cw.w; }
}

// decode-like step: pair select, pub store, stream bits from smem ring, candidates from a smem table
template <int V>
__global__ void k_dec(long long *cycles, uint32_t *sink) {
  __shared__ uint32_t ring[2048];
  __shared__ uint32_t cand[16 * 32];
  __shared__ uint2 sring[128];
  __shared__ uint32_t pub[4];
  fill_ring(ring, 2048);
  for (int i = threadIdx.x; i < 512; i += blockDim.x) cand[i] = ring[i] & 0xFFFF0000u;
  for (int i = threadIdx.x; i < 128; i += blockDim.x) sring[i] = make_uint2(ring[i] * 2654435761u, ring[i + 1] * 2654435761u);
  __syncthreads();
  const uint32_t ca = (uint32_t)__cvta_generic_to_shared(cand), sa = (uint32_t)__cvta_generic_to_shared(sring), pa = (uint32_t)__cvta_generic_to_shared(pub);
  uint32_t x1 = 0, x2 = 0xFFFFFFFFu, x = 0x12345678u, bp = 32, hist = 0, rhist = 0, flag_next = 2;
  w3_coder cw = w3_coder_new();
  uint2 pair = make_uint2(0x80000000u, 0x80000000u);
  long long t0 = clock64();
  for (uint32_t t = 0; t < N_STEPS; t += 16) {
#pragma unroll
    for (uint32_t i = 0; i < 16; i++) {
      const uint32_t P = (hist & 1u) ? pair.y : pair.x;
      const uint32_t s1 = (i + 1) & 15;
      if (V >= 1) { if (flag_next == 0xdeadbeefu) break; }
      uint2 pair_n; uint2 sw;
      asm volatile("ld.volatile.shared.v2.u32 {%0, %1}, [%2];" : "=r"(pair_n.x), "=r"(pair_n.y) : "r"(ca + s1 * 128 + (hist & 15u) * 8u) : "memory");
      if (V >= 1) flag_next = lds(pa + 8);
      asm volatile("ld.volatile.shared.v2.u32 {%0, %1}, [%2];" : "=r"(sw.x), "=r"(sw.y) : "r"(sa + ((bp >> 5) & 127u) * 8u) : "memory");
      if (V == 3) { /* width-carrying decoder step */
        const uint32_t m = __umulhi(cw.w, P);
        const uint32_t bit = (x - cw.x1) <= m ? 1u : 0u; /* x <= x1 + m, x >= x1 always */
        hist = (hist << 1) | bit;
        rhist = __funnelshift_r(rhist, bit, 1);
        sts(pa, __byte_perm(rhist, t + i + 1u, 0x5432));
        uint32_t flip;
        const int nsh = w3_coder_renorm(cw, bit ? cw.x1 : cw.x1 + m + 1u, bit ? m : cw.w - m - 1u, flip);
        const uint32_t fresh = __funnelshift_l(sw.y, sw.x, bp & 31u);
        x = __funnelshift_l(fresh, x, (uint32_t)nsh) ^ flip;
        bp += (uint32_t)nsh;
        pair = pair_n;
        x = cw.x1 + (cw.w >> 1) + (bp & 1); /* keep x inside the range (synthetic stream) */
        continue;
      }
      const uint32_t xmid = x1 + __umulhi(x2 - x1, P);
      const uint32_t bit = x <= xmid ? 1u : 0u;
      if (bit) x2 = xmid; else x1 = xmid + 1u;
      hist = (hist << 1) | bit;
      rhist = __funnelshift_r(rhist, bit, 1);
      sts(pa, __byte_perm(rhist, t + i + 1u, 0x5432));
      int s, k;
      const int nsh = w3_ac_renorm(x1, x2, s, k);
      const uint32_t fresh = __funnelshift_l(sw.y, sw.x, bp & 31u);
      x = __funnelshift_l(fresh, x, (uint32_t)nsh) ^ ((0u - (uint32_t)k) & 0x80000000u);
      bp += (uint32_t)nsh;
      pair = pair_n;
      if (V == 2) { x = x1 + ((x2 - x1) >> 1) + (bp & 1); }  // keep x inside the range (synthetic stream)
    }
  }
  long long t1 = clock64();
  if (threadIdx.x == 0) { cycles[blockIdx.x] = t1 - t0; sink[blockIdx.x] = x1 ^ x2 ^ x ^ bp ^ hist ^ flag_next ^ cw.x1 ^ cw.w; }
}

template <class F> void run(const char *name, F launch, long long *d_cyc, int blocks) {
  launch();
  cudaDeviceSynchronize();
  launch();
  cudaError_t e = cudaDeviceSynchronize();
  long long h[256];
  cudaMemcpy(h, d_cyc, sizeof(long long) * blocks, cudaMemcpyDeviceToHost);
  long long mx = 0, mn = 1ll << 62;
  for (int i = 0; i < blocks; i++) { if (h[i] > mx) mx = h[i]; if (h[i] < mn) mn = h[i]; }
  printf("%-44s blocks %3d  cycles/step min %.1f max %.1f  (%s)\n", name, blocks, (double)mn / N_STEPS, (double)mx / N_STEPS, cudaGetErrorString(e));
}

int main() {
  uint32_t *d_out, *d_sink; long long *d_cyc;
  cudaMalloc(&d_out, 256 * (N_STEPS / 8 + 64) * 4); cudaMalloc(&d_sink, 1024); cudaMalloc(&d_cyc, 2048);
  for (int blocks : {1, 64}) {
    run("enc V0 chain only (synthetic P)", [&] { k_enc<0><<<blocks, 32>>>(d_out, d_cyc, d_sink); }, d_cyc, blocks);
    run("enc V1 + LDS ring entries", [&] { k_enc<1><<<blocks, 32>>>(d_out, d_cyc, d_sink); }, d_cyc, blocks);
    run("enc V2 + writer, predicated st.global", [&] { k_enc<2><<<blocks, 32>>>(d_out, d_cyc, d_sink); }, d_cyc, blocks);
    run("enc V3 + writer, branch on word full", [&] { k_enc<3><<<blocks, 32>>>(d_out, d_cyc, d_sink); }, d_cyc, blocks);
    run("enc V4 + publish (pre,s,k) to smem", [&] { k_enc<4><<<blocks, 32>>>(d_out, d_cyc, d_sink); }, d_cyc, blocks);
    run("enc V5 width-carrying step + LDS ring entries", [&] { k_enc<5><<<blocks, 32>>>(d_out, d_cyc, d_sink); }, d_cyc, blocks);
    run("enc V6 width-carrying step + publish to smem", [&] { k_enc<6><<<blocks, 32>>>(d_out, d_cyc, d_sink); }, d_cyc, blocks);
    run("enc V7 as the kernel: state only + publish bounds", [&] { k_enc<7><<<blocks, 32>>>(d_out, d_cyc, d_sink); }, d_cyc, blocks);
    run("dec V0 step, no flag check", [&] { k_dec<0><<<blocks, 32>>>(d_cyc, d_sink); }, d_cyc, blocks);
    run("dec V1 step + flag check branch", [&] { k_dec<1><<<blocks, 32>>>(d_cyc, d_sink); }, d_cyc, blocks);
    run("dec V2 step, x kept in range", [&] { k_dec<2><<<blocks, 32>>>(d_cyc, d_sink); }, d_cyc, blocks);
    run("dec V3 width-carrying step, x kept in range", [&] { k_dec<3><<<blocks, 32>>>(d_cyc, d_sink); }, d_cyc, blocks);
  }
  return 0;
}
