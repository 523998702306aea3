// microbench3.cu — more op latencies; renorm with cvt.rz-based clz; coder loops with a rev-aware branch-free writer
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../weath3rb0i_b200/csrc/w3_core.cuh"
#define N_STEPS (1 << 18)
__device__ __forceinline__ uint32_t lds(uint32_t a) { uint32_t v; asm volatile("ld.volatile.shared.u32 %0, [%1];" : "=r"(v) : "r"(a) : "memory"); return v; }
__device__ __forceinline__ uint32_t bswap(uint32_t v) { return __byte_perm(v, 0, 0x0123); }

template <int OP>
__global__ void k_lat(long long *cycles, uint32_t *sink, uint32_t seed) {
  uint32_t a = seed | 1, b = seed * 3 + 7;
  long long t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < 4096; i++) {
#pragma unroll
    for (int j = 0; j < 16; j++) {
      if (OP == 0) { asm volatile("mad.lo.u32 %0, %1, %2, %3;" : "=r"(a) : "r"(a), "r"(b), "r"(b)); }
      if (OP == 1) { uint64_t w; asm volatile("mul.wide.u32 %0, %1, %2;" : "=l"(w) : "r"(a), "r"(b)); a = (uint32_t)(w >> 32) | 0x80000000u; }
      if (OP == 2) { float f; asm volatile("cvt.rz.f32.u32 %0, %1;" : "=f"(f) : "r"(a)); a = __float_as_uint(f); }
      if (OP == 3) { asm volatile("shr.u32 %0, %1, 23;" : "=r"(a) : "r"(a | 0x80000000u)); }
      if (OP == 4) { a = __funnelshift_l(b, a, a & 7); }
      if (OP == 5) { a = __byte_perm(a, b, 0x5432); }
      if (OP == 6) { asm volatile("shl.b32 %0, %1, %2;" : "=r"(a) : "r"(a), "r"(b)); a |= 1; }
      if (OP == 7) { asm volatile("xor.b32 %0, %1, %2;" : "=r"(a) : "r"(a), "r"(b)); asm volatile("add.u32 %0, %1, %2;" : "=r"(a) : "r"(a), "r"(b)); }
      if (OP == 8) { asm volatile("{.reg .pred p; setp.ge.u32 p, %1, %2; @p add.u32 %0, %1, 1; @!p sub.u32 %0, %1, 3;}" : "=r"(a) : "r"(a), "r"(b)); }
    }
  }
  long long t1 = clock64();
  if (threadIdx.x == 0) { cycles[0] = t1 - t0; sink[0] = a; }
}

// clz for non-zero v via round-toward-zero int->float: exponent is exact
__device__ __forceinline__ int clz_f(uint32_t v) {
  float f; asm("cvt.rz.f32.u32 %0, %1;" : "=f"(f) : "r"(v));
  return 158 - (int)(__float_as_uint(f) >> 23);
}
__device__ __forceinline__ int renorm_f(uint32_t &x1, uint32_t &x2, int &s, int &k) {
  s = clz_f(x1 ^ x2);
  uint32_t h = (((~x1 | x2) << 1) | 1u) << s;
  k = clz_f(h);
  int n = s + k;
  x1 = (x1 << n) & 0x7FFFFFFFu;
  x2 = (x2 << n) | ~(0xFFFFFFFFu << n) | 0x80000000u;
  return n;
}

struct wstate { uint64_t acc; uint32_t nb, words, rev; };
__device__ __noinline__ wstate writer_slow(wstate w, uint32_t pre, int s, int k, uint32_t *dst, uint32_t lane) {
  w3_writer wr; wr.acc = w.acc; wr.nb = w.nb; wr.words = w.words; wr.rev = w.rev;
  auto store = [&](uint32_t idx, uint32_t v) { if (lane == 0) dst[idx] = bswap(v); };
  w3_writer_step(wr, pre, s, k, store);
  w.acc = wr.acc; w.nb = wr.nb; w.words = wr.words; w.rev = wr.rev;
  return w;
}

// R: 0 = FLO renorm, 1 = cvt.rz renorm.  W: 0 = none (chain only), 1 = fast/slow split (noinline slow),
// 2 = rev-aware branch-free put + branch on word flush (slow only when s+rev > 32), 3 = same, flush by predicated plain store
template <int R, int W>
__global__ void k_coder(const uint32_t *trace, uint32_t *out, long long *cycles, uint32_t *sink) {
  __shared__ uint32_t ring[2048];
  for (int i = threadIdx.x; i < 2048; i += blockDim.x) ring[i] = trace[i];
  __syncthreads();
  const uint32_t lane = threadIdx.x & 31;
  uint32_t x1 = 0, x2 = 0xFFFFFFFFu;
  wstate w; w.acc = 0; w.nb = 0; w.words = 0; w.rev = 0;
  uint32_t *dst = out + blockIdx.x * (N_STEPS + 64);
  uint32_t slow = 0;
  long long t0 = clock64();
#pragma unroll 1
  for (int base = 0; base < N_STEPS; base += 32) {
    const uint32_t *gp = ring + (base & 2047);
    if (w.words > N_STEPS - 64) w.words = 0;
#pragma unroll
    for (int j = 0; j < 32; j++) {
      const uint32_t vj = gp[j];
      w3_ac_narrow(x1, x2, vj & 1u, vj & 0xFFFF0000u);
      const uint32_t pre = x1; int s, k;
      if (R == 0) w3_ac_renorm(x1, x2, s, k); else renorm_f(x1, x2, s, k);
      if (W == 1) {
        if (w.rev != 0u && s != 0) { w = writer_slow(w, pre, s, k, dst, lane); slow++; }
        else {
          w.acc = (w.acc << s) | __funnelshift_l(pre, 0u, (uint32_t)s); w.nb += s; w.rev += k;
          if (w.nb >= 32u) { w.nb -= 32u; if (lane == 0) dst[w.words] = bswap((uint32_t)(w.acc >> w.nb)); w.words++; }
        }
      } else if (W >= 2) {
        const uint32_t has = s != 0 ? 1u : 0u;
        const uint32_t ruse = has ? w.rev : 0u;       // pending bits resolved now
        const uint32_t c = (uint32_t)s + ruse;        // bits appended this step
        if (c > 32u) { w = writer_slow(w, pre, s, k, dst, lane); slow++; }
        else {
          const uint32_t top = __funnelshift_l(pre, 0u, (uint32_t)s);      // s bits, first = b
          const uint32_t b = pre >> 31;
          const uint32_t sm1 = (uint32_t)s - has;                          // s-1, or 0 when s == 0
          const uint32_t low = top & ~(0xFFFFFFFFu << sm1);                 // the s-1 bits after the first
          const uint64_t fill = b ? 0ull : ~(~0ull << ruse);                // ruse copies of !b
          const uint64_t V = ((uint64_t)(b & has) << (sm1 + ruse)) | (fill << sm1) | low;
          w.acc = (w.acc << c) | V; w.nb += c;
          w.rev = (has ? 0u : w.rev) + (uint32_t)k;
          if (W == 2) {
            if (w.nb >= 32u) { w.nb -= 32u; if (lane == 0) dst[w.words] = bswap((uint32_t)(w.acc >> w.nb)); w.words++; }
          } else {
            const bool full = w.nb >= 32u;
            w.nb -= full ? 32u : 0u;
            if (full && lane == 0) dst[w.words] = bswap((uint32_t)(w.acc >> w.nb));
            w.words += full ? 1u : 0u;
          }
        }
      }
    }
  }
  long long t1 = clock64();
  if (threadIdx.x == 0) { cycles[blockIdx.x] = t1 - t0; sink[blockIdx.x * 2] = x1 ^ x2 ^ (uint32_t)w.acc ^ w.nb ^ w.words ^ w.rev; sink[blockIdx.x * 2 + 1] = slow; }
}

int main() {
  uint32_t *d_out, *d_sink, *d_trace; long long *d_cyc;
  cudaMalloc(&d_out, 4ull * (N_STEPS + 64) * 4); cudaMalloc(&d_sink, 4096); cudaMalloc(&d_cyc, 2048); cudaMalloc(&d_trace, 8192);
  uint32_t h_trace[2048];
  FILE *f = fopen("tools/trace2048.bin", "rb");
  if (!f || fread(h_trace, 4, 2048, f) != 2048) { printf("no trace\n"); return 1; }
  cudaMemcpy(d_trace, h_trace, 8192, cudaMemcpyHostToDevice);
  const char *names[] = {"IMAD (mad.lo)", "IMAD.WIDE hi", "cvt.rz.f32.u32", "SHF const", "funnelshift var", "PRMT", "SHL var + LOP", "LOP + IADD", "ISETP + 2 predicated"};
  long long h; uint32_t hs[4];
#define LAT(OP) { k_lat<OP><<<1, 32>>>(d_cyc, d_sink, 12345); cudaDeviceSynchronize(); k_lat<OP><<<1, 32>>>(d_cyc, d_sink, 777); cudaDeviceSynchronize(); cudaMemcpy(&h, d_cyc, 8, cudaMemcpyDeviceToHost); printf("lat %-28s %.2f cycles/op\n", names[OP], (double)h / (4096 * 16)); }
  LAT(0) LAT(1) LAT(2) LAT(3) LAT(4) LAT(5) LAT(6) LAT(7) LAT(8)
#define COD(R, W, name) { k_coder<R, W><<<1, 32>>>(d_trace, d_out, d_cyc, d_sink); cudaDeviceSynchronize(); k_coder<R, W><<<1, 32>>>(d_trace, d_out, d_cyc, d_sink); cudaError_t e = cudaDeviceSynchronize(); cudaMemcpy(&h, d_cyc, 8, cudaMemcpyDeviceToHost); cudaMemcpy(hs, d_sink, 8, cudaMemcpyDeviceToHost); printf("coder %-58s %.1f cycles/step  state %08x slow %u (%s)\n", name, (double)h / N_STEPS, hs[0], hs[1], cudaGetErrorString(e)); }
  COD(0, 0, "R0 FLO renorm, chain only")
  COD(1, 0, "R1 cvt.rz renorm, chain only")
  COD(0, 1, "R0 + fast/slow split writer")
  COD(1, 1, "R1 + fast/slow split writer")
  COD(1, 2, "R1 + rev-aware put, branch on flush")
  COD(1, 3, "R1 + rev-aware put, predicated flush")
  return 0;
}
