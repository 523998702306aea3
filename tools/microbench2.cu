// microbench2.cu — (1) dependent-issue latency of the ops on the coder chain,
// (2) the encode coder loop variants fed with a real (p,bit) trace.
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cuda_runtime.h>
#include "../weath3rb0i_b200/csrc/w3_core.cuh"

#define N_STEPS (1 << 18)
__device__ __forceinline__ uint32_t lds(uint32_t a) { uint32_t v; asm volatile("ld.volatile.shared.u32 %0, [%1];" : "=r"(v) : "r"(a) : "memory"); return v; }
__device__ __forceinline__ uint32_t bswap(uint32_t v) { return __byte_perm(v, 0, 0x0123); }

// ---- latency chains ---------------------------------------------------------
template <int OP>
__global__ void k_lat(long long *cycles, uint32_t *sink, uint32_t seed) {
  __shared__ uint32_t sm[1024];
  for (int i = threadIdx.x; i < 1024; i += 32) sm[i] = (i * 4 + 64) & 4095;
  __syncthreads();
  uint32_t a = seed | 1, b = seed * 3 + 7;
  const uint32_t sa = (uint32_t)__cvta_generic_to_shared(sm);
  long long t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < 4096; i++) {
#pragma unroll
    for (int j = 0; j < 16; j++) {
      if (OP == 0) { asm volatile("bfind.shiftamt.u32 %0, %1;" : "=r"(a) : "r"(a | 0x100)); }
      if (OP == 1) { a = __umulhi(a, b) | 0x80000000u; }
      if (OP == 2) { asm volatile("shl.b32 %0, %1, %2;" : "=r"(a) : "r"(a | 1), "r"(b & 3)); }
      if (OP == 3) { asm volatile("xor.b32 %0, %1, %2;" : "=r"(a) : "r"(a), "r"(b)); }
      if (OP == 4) { asm volatile("{.reg .pred p; setp.gt.u32 p, %1, %2; selp.u32 %0, %1, %2, p;}" : "=r"(a) : "r"(a), "r"(b)); b += 3; }
      if (OP == 5) { a = lds(sa + (a & 4092)); }
      if (OP == 6) { asm volatile("add.u32 %0, %1, %2;" : "=r"(a) : "r"(a), "r"(b)); }
      if (OP == 7) { asm volatile("brev.b32 %0, %1;" : "=r"(a) : "r"(a)); }
      if (OP == 8) { asm volatile("popc.b32 %0, %1;" : "=r"(a) : "r"(a | 0x10000)); }
      if (OP == 9) { float f; asm volatile("cvt.rn.f32.u32 %0, %1;" : "=f"(f) : "r"(a)); a = __float_as_uint(f); }
      if (OP == 10) { a = __shfl_sync(0xFFFFFFFFu, a, (b + j) & 31); }
    }
  }
  long long t1 = clock64();
  if (threadIdx.x == 0) { cycles[0] = t1 - t0; sink[0] = a; }
}

// ---- coder loop variants ----------------------------------------------------
struct wstate { uint64_t acc; uint32_t nb, words, rev; };

__device__ __noinline__ wstate writer_slow(wstate w, uint32_t pre, int s, int k, uint32_t *dst, uint32_t lane) {
  w3_writer wr; wr.acc = w.acc; wr.nb = w.nb; wr.words = w.words; wr.rev = w.rev;
  auto store = [&](uint32_t idx, uint32_t v) { if (lane == 0) dst[idx] = bswap(v); };
  w3_writer_step(wr, pre, s, k, store);
  w.acc = wr.acc; w.nb = wr.nb; w.words = wr.words; w.rev = wr.rev;
  return w;
}

// V: 0 = inline general writer (w3_writer_step), 1 = fast/slow split with noinline slow,
//    2 = fast only + trouble flag, redo per 32 with general writer, 3 = like 1 but plain (non-volatile) ring loads,
//    4 = like 1 but ring entries pre-split: P word and bit word (no masking)
template <int V>
__global__ void k_coder(const uint32_t *trace, uint32_t *out, long long *cycles, uint32_t *sink) {
  __shared__ uint32_t ring[2048];
  for (int i = threadIdx.x; i < 2048; i += blockDim.x) ring[i] = trace[i];
  __syncthreads();
  const uint32_t ra = (uint32_t)__cvta_generic_to_shared(ring);
  const uint32_t lane = threadIdx.x & 31;
  uint32_t x1 = 0, x2 = 0xFFFFFFFFu;
  wstate w; w.acc = 0; w.nb = 0; w.words = 0; w.rev = 0;
  uint32_t *dst = out + blockIdx.x * (N_STEPS + 64);
  uint32_t redo = 0;
  long long t0 = clock64();
#pragma unroll 1
  for (int base = 0; base < N_STEPS; base += 32) {
    const uint32_t ga = ra + (base & 2047) * 4;
    if (w.words > N_STEPS - 64) w.words = 0;
    if (V == 2) {
      const uint32_t cx1 = x1, cx2 = x2; const wstate cw = w; uint32_t trouble = 0;
      uint32_t v0 = lds(ga), v1 = lds(ga + 4);
#pragma unroll
      for (int j = 0; j < 32; j++) {
        const uint32_t vj = v0; v0 = v1; if (j + 2 < 32) v1 = lds(ga + (j + 2) * 4);
        w3_ac_narrow(x1, x2, vj & 1u, vj & 0xFFFF0000u);
        const uint32_t pre = x1; int s, k; w3_ac_renorm(x1, x2, s, k);
        trouble |= (w.rev != 0u && s != 0) ? 1u : 0u;
        w.acc = (w.acc << s) | __funnelshift_l(pre, 0u, (uint32_t)s); w.nb += s; w.rev += k;
        if (w.nb >= 32u) { w.nb -= 32u; if (lane == 0) dst[w.words] = bswap((uint32_t)(w.acc >> w.nb)); w.words++; }
      }
      if (trouble) {
        redo++; x1 = cx1; x2 = cx2; w = cw;
        for (int j = 0; j < 32; j++) { const uint32_t vj = lds(ga + j * 4); w3_ac_narrow(x1, x2, vj & 1u, vj & 0xFFFF0000u); const uint32_t pre = x1; int s, k; w3_ac_renorm(x1, x2, s, k); w = writer_slow(w, pre, s, k, dst, lane); }
      }
    } else {
      uint32_t v0, v1;
      if (V == 3) { v0 = ring[(base & 2047)]; v1 = ring[(base & 2047) + 1]; } else { v0 = lds(ga); v1 = lds(ga + 4); }
#pragma unroll
      for (int j = 0; j < 32; j++) {
        const uint32_t vj = v0; v0 = v1;
        if (j + 2 < 32) { if (V == 3) v1 = ring[(base & 2047) + j + 2]; else v1 = lds(ga + (j + 2) * 4); }
        w3_ac_narrow(x1, x2, vj & 1u, vj & 0xFFFF0000u);
        const uint32_t pre = x1; int s, k; w3_ac_renorm(x1, x2, s, k);
        if (V == 0) {
          w3_writer wr; wr.acc = w.acc; wr.nb = w.nb; wr.words = w.words; wr.rev = w.rev;
          auto store = [&](uint32_t idx, uint32_t v) { if (lane == 0) dst[idx] = bswap(v); };
          w3_writer_step(wr, pre, s, k, store);
          w.acc = wr.acc; w.nb = wr.nb; w.words = wr.words; w.rev = wr.rev;
        } else {
          if (w.rev != 0u && s != 0) { w = writer_slow(w, pre, s, k, dst, lane); redo++; }
          else {
            w.acc = (w.acc << s) | __funnelshift_l(pre, 0u, (uint32_t)s); w.nb += s; w.rev += k;
            if (w.nb >= 32u) { w.nb -= 32u; if (lane == 0) dst[w.words] = bswap((uint32_t)(w.acc >> w.nb)); w.words++; }
          }
        }
      }
    }
  }
  long long t1 = clock64();
  if (threadIdx.x == 0) { cycles[blockIdx.x] = t1 - t0; sink[blockIdx.x * 2] = x1 ^ x2 ^ (uint32_t)w.acc ^ w.nb ^ w.words ^ w.rev; sink[blockIdx.x * 2 + 1] = redo; }
}

int main() {
  uint32_t *d_out, *d_sink, *d_trace; long long *d_cyc;
  cudaMalloc(&d_out, 64ull * (N_STEPS + 64) * 4); cudaMalloc(&d_sink, 4096); cudaMalloc(&d_cyc, 2048); cudaMalloc(&d_trace, 8192);
  uint32_t h_trace[2048];
  FILE *f = fopen("tools/trace2048.bin", "rb");
  if (!f || fread(h_trace, 4, 2048, f) != 2048) { printf("no trace\n"); return 1; }
  cudaMemcpy(d_trace, h_trace, 8192, cudaMemcpyHostToDevice);
  const char *names[] = {"FLO.SH (bfind.shiftamt)", "IMAD.HI (umulhi)", "SHF (shl var)", "LOP3 (xor)", "ISETP+SEL", "LDS volatile", "IADD", "BREV", "POPC", "I2F (cvt.rn.f32.u32)", "SHFL.IDX"};
  long long h; uint32_t hs[4];
#define LAT(OP) { k_lat<OP><<<1, 32>>>(d_cyc, d_sink, 12345); cudaDeviceSynchronize(); k_lat<OP><<<1, 32>>>(d_cyc, d_sink, 777); cudaDeviceSynchronize(); cudaMemcpy(&h, d_cyc, 8, cudaMemcpyDeviceToHost); printf("lat %-28s %.2f cycles/op\n", names[OP], (double)h / (4096 * 16)); }
  LAT(0) LAT(1) LAT(2) LAT(3) LAT(4) LAT(5) LAT(6) LAT(7) LAT(8) LAT(9) LAT(10)
#define COD(V, name) { k_coder<V><<<1, 32>>>(d_trace, d_out, d_cyc, d_sink); cudaDeviceSynchronize(); k_coder<V><<<1, 32>>>(d_trace, d_out, d_cyc, d_sink); cudaError_t e = cudaDeviceSynchronize(); cudaMemcpy(&h, d_cyc, 8, cudaMemcpyDeviceToHost); cudaMemcpy(hs, d_sink, 8, cudaMemcpyDeviceToHost); printf("coder %-52s %.1f cycles/step  state %08x slow/redo %u (%s)\n", name, (double)h / N_STEPS, hs[0], hs[1], cudaGetErrorString(e)); }
  COD(0, "V0 inline general writer (w3_writer_step)")
  COD(1, "V1 fast/slow split, noinline slow path")
  COD(2, "V2 fast only + trouble flag, redo per 32 steps")
  COD(3, "V3 like V1, plain ld.shared ring loads")
  return 0;
}
