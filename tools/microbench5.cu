// microbench5.cu — the team decoder's step in isolation: candidates always valid (tags
// pre-filled for the positions visited), stream ring pre-filled, no other warps (or
// spinners only), to separate the decoder's own cost from waiting and contention.
#include <cstdio>
#include "../weath3rb0i_b200/csrc/w3_kernels.cu"

/* the out-of-line candidate wait the kernel used before its decoder went branch-free */
__device__ __noinline__ uint2 w3_pair_spin(uint32_t pa, uint32_t tag, uint32_t abort_a) {
  uint2 pr = make_uint2(0u, 0u);
  for (uint32_t spins = 0; spins < W3_SPIN_LIMIT; spins++) {
    pr = w3_lds64(pa);
    if (((pr.x ^ tag) & 0xFFFFu) == 0u) return pr;
    if ((spins & 255u) == 255u && w3_lds(abort_a)) break;
  }
  w3_sts(abort_a, 1u);
  return pr;
}

template <int V>
__global__ void __launch_bounds__(384) k_dec_alone(long long *cycles, uint32_t *sink, int spinners) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  w3_dec_team *tm = reinterpret_cast<w3_dec_team *>(smem_raw);
  const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31u;
  for (uint32_t i = threadIdx.x; i < W3_SRING; i += blockDim.x) tm->sring[i] = make_uint2(i * 2654435761u, (i + 1) * 2654435761u);
  if (threadIdx.x == 0) { tm->pub = 0; tm->abort_flag = 0; tm->fed = 1u << 30; tm->cons = 0; }
  __syncthreads();
  const uint32_t cand_a = w3_saddr(&tm->cand[0][0]), pub_a = w3_saddr(&tm->pub), abort_a = w3_saddr(&tm->abort_flag);
  const uint32_t sring_a = w3_saddr(&tm->sring[0]), fed_a = w3_saddr(&tm->fed), cons_a = w3_saddr(&tm->cons);
  if (warp != 0) {
    if ((int)warp > spinners) return;
    if (warp == 1) {
      // keep the candidate ring valid: rewrite tags for positions ahead of the decoder
      uint32_t done = 0, t = 0;
      for (;;) {
        const uint32_t pv = w3_lds(pub_a);
        if (pv == 0xFFFFFFFFu) break;
        t += ((pv >> 16) - t) & 0xFFFFu;
        while ((int32_t)(done - (t + 12u)) < 0 || done < 16u) {
          const uint32_t tag = (done + 1u) & 0xFFFFu;
          w3_sts(cand_a + (done & 15u) * 128u + lane * 4u, ((0x4000u + lane * 977u + done * 131u) << 16) | tag);
          done++;
        }
      }
      return;
    }
    uint32_t n = 0;
    while (w3_lds(pub_a) != 0xFFFFFFFFu) n++;
    if (lane == 0) sink[warp] = n;
    return;
  }
  uint32_t x1 = 0, x2 = 0xFFFFFFFFu, x = 0x12345678u, bp = 32, rhist = 0, xmid = 0;
  bool dead = false;
  auto next_mid = [&](uint2 pr, uint32_t bit) -> uint32_t { return x1 + __umulhi(x2 - x1, (bit ? pr.y : pr.x) & 0xFFFF0000u); };
  auto step = [&](uint32_t tt, uint32_t slot) {
    const uint32_t s1 = (slot + 1u) & (W3_CAND_RING - 1u);
    uint2 pair_n = w3_lds64(cand_a + s1 * (W3_NGUESS * 4u) + (rhist >> (33u - W3_SPEC_D)) * 8u);
    const uint2 sw = w3_lds64(sring_a + ((bp >> 5) & (W3_SRING - 1u)) * 8u);
    const uint32_t bit = x <= xmid ? 1u : 0u;
    if (bit) x2 = xmid; else x1 = xmid + 1u;
    rhist = __funnelshift_r(rhist, bit, 1);
    w3_sts(pub_a, __byte_perm(rhist, tt + 1u, 0x5432));
    int s, k;
    const int nsh = w3_ac_renorm(x1, x2, s, k);
    const uint32_t fresh = __funnelshift_l(sw.y, sw.x, bp & 31u);
    x = __funnelshift_l(fresh, x, (uint32_t)nsh) ^ ((0u - (uint32_t)k) & 0x80000000u);
    if (V == 1) x = x1 + ((x2 - x1) >> 1) + (bp & 3);
    bp += (uint32_t)nsh;
    xmid = next_mid(pair_n, bit);
    if (V != 3 && __builtin_expect(__any_sync(FULL, ((pair_n.x ^ (tt + 2u)) & 0xFFFFu) != 0u), 0)) {
      pair_n = w3_pair_spin(cand_a + s1 * (W3_NGUESS * 4u) + ((rhist << 1) >> (33u - W3_SPEC_D)) * 8u, tt + 2u, abort_a);
      xmid = next_mid(pair_n, bit);
    }
  };
  const uint32_t nbits = 1u << 18;
  xmid = next_mid(w3_pair_spin(cand_a, 1u, abort_a), 0u);
  if (V == 4) {
    /* branch-free attempts: an attempt decodes bit t iff the split point computed by the
     * previous attempt was built from valid candidates; otherwise it only re-fetches */
    uint32_t t = 0, ok = 1;
    long long t0b = clock64();
    while (t < nbits) {
#pragma unroll
      for (uint32_t i = 0; i < 16u; i++) {
        const uint32_t pos_n = t + ok;                               /* position whose candidates we need next */
        const uint32_t idx = ok ? (rhist >> (33u - W3_SPEC_D)) : ((rhist << 1) >> (33u - W3_SPEC_D));
        const uint2 pair_n = w3_lds64(cand_a + (pos_n & (W3_CAND_RING - 1u)) * (W3_NGUESS * 4u) + idx * 8u);
        const uint2 sw = w3_lds64(sring_a + ((bp >> 5) & (W3_SRING - 1u)) * 8u);
        const uint32_t bit = x <= xmid ? 1u : 0u;
        if (ok && bit) x2 = xmid;
        if (ok && !bit) x1 = xmid + 1u;
        rhist = ok ? __funnelshift_r(rhist, bit, 1) : rhist;
        t += ok;
        w3_sts(pub_a, __byte_perm(rhist, t, 0x5432));
        int s, k;
        const int nsh = w3_ac_renorm(x1, x2, s, k);                  /* a normalised pair renormalises to itself */
        const uint32_t fresh = __funnelshift_l(sw.y, sw.x, bp & 31u);
        x = __funnelshift_l(fresh, x, (uint32_t)nsh) ^ ((0u - (uint32_t)k) & 0x80000000u);
        bp += (uint32_t)nsh;
        xmid = x1 + __umulhi(x2 - x1, (((int32_t)rhist < 0) ? pair_n.y : pair_n.x) & 0xFFFF0000u);
        ok = (((pair_n.x ^ (t + 1u)) & 0xFFFFu) == 0u) ? 1u : 0u;
      }
    }
    long long t1b = clock64();
    w3_sts(pub_a, 0xFFFFFFFFu);
    if (lane == 0) { cycles[0] = t1b - t0b; sink[0] = x1 ^ x2 ^ x ^ bp ^ rhist ^ ok; }
    return;
  }
  long long t0 = clock64();
  for (uint32_t t = 0; t + 16u <= nbits && !dead; t += 16u) {
    if (V != 2 && V != 3 && (t & 63u) == 0u) { w3_sts(cons_a, bp >> 5); if ((int32_t)(w3_lds(fed_a) - ((bp >> 5) + 66u)) < 0) dead = true; }
#pragma unroll
    for (uint32_t i = 0; i < 16u; i++) step(t + i, i);
  }
  long long t1 = clock64();
  w3_sts(pub_a, 0xFFFFFFFFu);
  if (lane == 0) { cycles[0] = t1 - t0; sink[0] = x1 ^ x2 ^ x ^ bp ^ rhist ^ (uint32_t)dead; }
}

int main() {
  long long *d_cyc; uint32_t *d_sink; long long hc;
  cudaMalloc(&d_cyc, 64); cudaMalloc(&d_sink, 256);
  size_t smem = sizeof(w3_dec_team);
  for (int spinners : {1, 10}) {
#define RUN(V, name) { cudaFuncSetAttribute(k_dec_alone<V>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); k_dec_alone<V><<<1, 384, smem>>>(d_cyc, d_sink, spinners); cudaError_t e = cudaDeviceSynchronize(); cudaMemcpy(&hc, d_cyc, 8, cudaMemcpyDeviceToHost); printf("%-58s other warps %2d: %.1f cycles/step (%s)\n", name, spinners, (double)hc / (1 << 18), cudaGetErrorString(e)); }
    RUN(0, "decoder step as in the kernel (random stream)")
    RUN(1, "decoder step, x kept inside the range")
    RUN(2, "decoder step, x in range, no per-trip stream check")
    RUN(3, "decoder step, no tag check at all (no branch in the step)")
    RUN(4, "branch-free attempts (decode iff previous candidates were valid)")
  }
  return 0;
}
