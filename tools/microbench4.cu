// microbench4.cu — the speculator's lookup in isolation (one warp, real table), with and
// without other warps spin-polling shared memory in the same CTA.
#include <cstdio>
#include "../weath3rb0i_b200/csrc/w3_kernels.cu"

template <int V>
__global__ void __launch_bounds__(384) k_spec(const w3_hashtab *tab, long long *cycles, uint32_t *sink, int spinners) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  w3_dec_tab *st = reinterpret_cast<w3_dec_tab *>(smem_raw);
  w3_dec_team *tm = reinterpret_cast<w3_dec_team *>(smem_raw + sizeof(w3_dec_tab));
  const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31u;
  {
    const uint4 *gt = reinterpret_cast<const uint4 *>(&tab->thr[0][0]);
    uint4 *dt = reinterpret_cast<uint4 *>(st->thr);
    for (uint32_t i = threadIdx.x; i < 8u * 256u * 8u / 16u; i += blockDim.x) dt[i] = gt[i];
    if (threadIdx.x < 8) st->thr[8 * 256 + threadIdx.x] = ~(uint64_t)0;
    const uint16_t *gc = &tab->cell2[0][0];
    for (uint32_t i = threadIdx.x; i < 8u * W3_NCELL2; i += blockDim.x)
      st->cell2[(i & ~(W3_NCELL2 - 1u)) | w3_cell2_swizzle(i & (W3_NCELL2 - 1u))] = __ldg(gc + i);
  }
  for (uint32_t i = threadIdx.x; i < W3_NCTX + 8u; i += blockDim.x) tm->probs[i] = 0x80000000u ^ (i << 16);
  if (threadIdx.x == 0) { tm->pub = 0; tm->abort_flag = 0; }
  __syncthreads();
  const uint32_t pub_a = w3_saddr(&tm->pub), probs_a = w3_saddr(&tm->probs[0]), cand_a = w3_saddr(&tm->cand[0][0]);
  if (warp != 0) {
    if ((int)warp > spinners) return;
    // spin-poll like a waiting speculator until warp 0 is done
    uint32_t n = 0;
    while (w3_lds(pub_a) != 0xFFFFFFFFu) n++;
    if (lane == 0) sink[warp] = n;
    return;
  }
  const uint64_t gk0 = (uint64_t)(__brev(lane) >> 26) << 58;
  const uint64_t gk1 = gk0 | ((uint64_t)1 << 58);
  uint64_t K = 0x6f6c6c6568206f77ull;  // some text-like history
  uint32_t acc = 0;
  const int N = 20000;
  long long t0 = clock64();
  for (int it = 0; it < N; it++) {
    const uint32_t pos = (uint32_t)it;
    const uint64_t kb = K >> 6;
    uint32_t r0, r1;
    if (V == 0) {  // fixed lookups only
      r0 = w3_rank_lookup_fixed<true>(kb | gk0, pos & 7u, st->thr, st->cell2);
      r1 = w3_rank_lookup_fixed<true>(kb | gk1, pos & 7u, st->thr, st->cell2);
    } else {       // + slot read + publish like the kernel
      r0 = w3_rank_lookup_fixed<true>(kb | gk0, pos & 7u, st->thr, st->cell2);
      r1 = w3_rank_lookup_fixed<true>(kb | gk1, pos & 7u, st->thr, st->cell2);
      const uint32_t ctx0 = ((r0 & 255u) << 3) | (pos & 7u), ctx1 = ((r1 & 255u) << 3) | (pos & 7u);
      const uint32_t P0 = w3_lds(probs_a + w3_dec_slot(ctx0) * 4u);
      const uint32_t P1 = w3_lds(probs_a + w3_dec_slot(ctx1) * 4u);
      const uint32_t slot = pos & (W3_CAND_RING - 1u);
      w3_sts(cand_a + slot * 128u + lane * 4u, P0);
      w3_sts(cand_a + slot * 128u + lane * 4u, P1);
      __syncwarp();
      if (lane == 0) { if (V == 1) __threadfence_block(); w3_sts(pub_a + 4u, pos + 1u); }
    }
    acc += r0 + r1;
    // next history depends on the result (serialises iterations like the real wait does)
    K = (K >> 1) | ((uint64_t)((__shfl_sync(0xFFFFFFFFu, acc, 0) >> 3) & 1u) << 63) ^ ((uint64_t)it * 0x9E3779B97F4A7C15ull >> 40);
  }
  long long t1 = clock64();
  w3_sts(pub_a, 0xFFFFFFFFu);
  if (lane == 0) { cycles[0] = t1 - t0; sink[0] = acc; }
}

int main() {
  w3_hashtab *d_tab; long long *d_cyc; uint32_t *d_sink;
  cudaMalloc(&d_tab, sizeof(w3_hashtab)); cudaMalloc(&d_cyc, 64); cudaMalloc(&d_sink, 256);
  w3_hashtab *h = (w3_hashtab *)calloc(1, sizeof(w3_hashtab));
  const uint16_t book1[8] = W3_BOOK1_TABLE;
  for (int i = 0; i < 8; i++) h->table[i] = book1[i];
  h->max_bits = 8;
  cudaMemcpy(d_tab, h, sizeof(w3_hashtab), cudaMemcpyHostToDevice);
  w3_launch_build_table(d_tab, 0);
  cudaDeviceSynchronize();
  size_t smem = sizeof(w3_dec_tab) + sizeof(w3_dec_team);
  cudaFuncSetAttribute(k_spec<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  cudaFuncSetAttribute(k_spec<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  cudaFuncSetAttribute(k_spec<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  long long hc;
  for (int spinners : {0, 3, 11}) {
#define RUN(V, name) { k_spec<V><<<1, 384, smem>>>(d_tab, d_cyc, d_sink, spinners); cudaError_t e = cudaDeviceSynchronize(); cudaMemcpy(&hc, d_cyc, 8, cudaMemcpyDeviceToHost); printf("%-52s spinners %2d: %.1f cycles/position (%s)\n", name, spinners, (double)hc / 20000, cudaGetErrorString(e)); }
    RUN(0, "two fixed lookups")
    RUN(1, "two lookups + slot reads + publish (fence)")
    RUN(2, "two lookups + slot reads + publish (no fence)")
  }
  return 0;
}
