/*
 * w3_lane.cuh — "lane" decode: one LANE per block, for the many-block regime
 * (1 GiB in 1 MiB blocks on one GPU is 7 blocks per SM; 256 KiB blocks, 28).
 *
 * decompress() of the reference (src/main.rs:131-140) is one dependent chain
 * per stream: bit -> history -> hash -> context -> probability -> bit.  With
 * one WARP per block every instruction of that chain is issued for 32 lanes
 * that all hold the same values (round 1: 250 warp-instructions per coded bit,
 * 678 cycles per bit at 7 blocks per SM, profiles/r01_1gib_*).  The integer
 * pipes of an SM sub-partition retire one warp-instruction per two cycles
 * whatever the number of active lanes, so the instruction, not the lane, is
 * the unit of cost: here every lane runs its own block, and one instruction
 * stream decodes up to 32 blocks at once.  Lanes never talk to each other, so
 * the per-lane code is plain sequential code — `__host__ __device__`, and
 * tests/hostsim runs exactly these statements against the oracle on the CPU.
 *
 * Shared memory of a CTA: one copy of the lookup tables (w3_lane_tab, 72 KiB)
 * and 8 KiB of counters per resident block (Counter — src/models/counter.rs;
 * the probability is recomputed from the counts at predict time: one
 * reciprocal, fewer instructions than keeping a second table current).  HBM
 * sees the compressed streams (read once, through L1) and the output words.
 *
 * The entropy hash (ACHistory::hash — src/history/ac_history.rs:28-46) is the
 * rank of the bit-reversed history among <= 255 thresholds per alignment
 * (w3_core.cuh).  Lookup: a byte per (alignment, top 13 key bits) gives the
 * number of thresholds below the key's cell; the next W3_LANE_FIX thresholds'
 * high words are loaded at once and compared (thresholds ascend, so those past
 * the cell compare greater by themselves).  A tie on a high word, or all
 * W3_LANE_FIX compares true, sends that lane through the general search
 * (w3_rank_lookup, w3_achash_exact) over the tables in global memory.
 */
#pragma once
#include "w3_core.cuh"

#define W3_LANE_PB W3_PB2
#define W3_LANE_NCELL (1u << W3_LANE_PB)
#define W3_LANE_FIX 6
#define W3_LANE_ROW 264u /* 255 thresholds + ~0 padding: a lookup reads W3_LANE_FIX entries from any base <= 255 */
static_assert(255u + W3_LANE_FIX <= W3_LANE_ROW, "threshold rows must cover the read-ahead");

struct w3_lane_tab {
  uint32_t thi[8 * W3_LANE_ROW];    /* high words of thr[a][j] */
  uint8_t cellb[8 * W3_LANE_NCELL]; /* #{j : thr[a][j] < first key of the cell} */
};
static_assert(sizeof(w3_lane_tab) % 16 == 0, "counters follow the tables at 16-byte alignment");

/* entry i of the staged tables, from the context's table in global memory */
W3_HD uint32_t w3_lane_tab_thi(const w3_hashtab *g, uint32_t i) {
  const uint32_t a = i / W3_LANE_ROW, j = i % W3_LANE_ROW;
  return j < 255u ? (uint32_t)(g->thr[a][j] >> 32) : 0xFFFFFFFFu;
}
W3_HD uint8_t w3_lane_tab_cellb(const w3_hashtab *g, uint32_t i) {
  return (uint8_t)(g->cell2[i / W3_LANE_NCELL][i % W3_LANE_NCELL] & 0xFFu);
}

/* rank of a key among the thresholds of alignment a, from the key's high word */
W3_HD uint32_t w3_lane_rank(uint32_t khi, uint32_t a, const uint32_t *thi, const uint8_t *cellb, bool &bad) {
  const uint32_t base = cellb[a * W3_LANE_NCELL + (khi >> (32u - W3_LANE_PB))];
  const uint32_t *t = thi + a * W3_LANE_ROW + base;
  uint32_t b[W3_LANE_FIX];
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
  for (int i = 0; i < W3_LANE_FIX; i++) b[i] = t[i];
  uint32_t r = base;
  bool tie = false;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
  for (int i = 0; i < W3_LANE_FIX; i++) {
    r += khi > b[i] ? 1u : 0u;
    tie |= khi == b[i];
  }
  bad = tie || khi > b[W3_LANE_FIX - 1]; /* all compares true: there may be more thresholds below the key */
  return r;
}
/* the general search, for the lane the fixed lookup declined (rare) */
#if defined(__CUDACC__)
__host__ __device__ __noinline__
#else
static
#endif
uint32_t w3_lane_rank_slow(uint32_t khi, uint32_t klo, uint32_t a, const w3_hashtab *g) {
  const uint64_t key = ((uint64_t)khi << 32) | klo;
  uint32_t r = w3_rank_lookup(key, a, &g->thr[0][0], &g->cell[0][0]);
  if (r == 0xFFFFFFFFu)
    r = w3_rank_of_hash(w3_achash_exact(w3_key_of_history(key), a, g->table, (int)g->max_bits).hash);
  return r;
}

/* the next 32 stream bits at bit position bp of the big-endian bit stream that
 * starts at the aligned word pointer ab; bits at or past end_bits read as zero
 * (ACReader — src/entropy_coding/io.rs:18-48).  kCareful = false: the caller
 * knows bp + 64 <= end_bits. */
template <bool kCareful>
W3_HD uint32_t w3_lane_fresh(const uint32_t *ab, uint32_t bp, uint32_t end_bits) {
  const uint32_t wi = bp >> 5;
  uint32_t w0, w1;
  if (!kCareful) {
    w0 = w3_bswap32(w3_ldg32(ab + wi));
    w1 = w3_bswap32(w3_ldg32(ab + wi + 1u));
  } else {
    const uint32_t s0 = wi << 5;
    w0 = w1 = 0u;
    if (s0 < end_bits) {
      w0 = w3_bswap32(w3_ldg32(ab + wi));
      const uint32_t valid = end_bits - s0;
      if (valid < 32u) w0 &= 0xFFFFFFFFu << (32u - valid);
      if (valid > 32u) {
        w1 = w3_bswap32(w3_ldg32(ab + wi + 1u));
        if (valid < 64u) w1 &= 0xFFFFFFFFu << (64u - valid);
      }
    }
  }
  return w3_fsl(w1, w0, bp);
}

struct w3_lane_dec {
  uint32_t x1, x2, x; /* ArithmeticCoder — arithmetic_coder.rs:11-18 */
  uint32_t bp;        /* stream bits consumed, counted from the aligned base */
  uint32_t khi, klo;  /* ACHistory.bits, newest bit at bit 63 (w3_key_of_history) */
  uint32_t ci;        /* counter of the current context: (rank << 3) | alignment */
};

/* one pass of the decompress loop (main.rs:134-137) for the bit whose successor
 * has alignment a1; cnt = this block's 2048 counters */
template <bool kCareful>
W3_HD uint32_t w3_lane_dec_step(w3_lane_dec &s, uint32_t a1, uint32_t *cnt, const w3_lane_tab *lt,
                                const w3_hashtab *g, const uint32_t *ab, uint32_t end_bits, uint32_t ones) {
  const uint32_t c = cnt[s.ci];
  const uint32_t P = w3_counter_p(c & 0xFFFFu, c >> 16) << 16;  /* predict — main.rs:134 */
  const uint32_t m = w3_umulhi(s.x2 - s.x1, P);                 /* lerp — arithmetic_coder.rs:81 */
  const uint32_t bit = (s.x - s.x1) <= m ? 1u : 0u;             /* :82; x >= x1 always */
  if (bit) s.x2 = s.x1 + m;                                     /* :85-88 */
  else s.x1 = s.x1 + m + 1u;
  cnt[s.ci] = w3_counter_update(c, bit);                        /* adapt — main.rs:136, ordern_entropy.rs:32-34 */
  s.klo = w3_fsr(s.klo, s.khi, 1u);                             /* history — ac_history.rs:23-26 */
  s.khi = (s.khi >> 1) | (bit << 31);
  bool bad;
  uint32_t r = w3_lane_rank(s.khi, a1, lt->thi, lt->cellb, bad); /* hash — ordern_entropy.rs:36-45 */
  if (bad) r = w3_lane_rank_slow(s.khi, s.klo, a1, g);
  s.ci = (r << W3_ALIGN_BITS) | a1;
  /* both renormalisation loops (:91-102) in closed form; of the E3 flips of x only the last survives */
  const uint32_t fresh = w3_lane_fresh<kCareful>(ab, s.bp, end_bits);
  int sh_s, sh_k;
  const int n = w3_ac_renorm_ones(s.x1, s.x2, sh_s, sh_k, ones);
  s.x = w3_fsl(fresh, s.x, (uint32_t)n) ^ ((0u - (uint32_t)sh_k) & 0x80000000u);
  s.bp += (uint32_t)n;
  if (kCareful) { /* everything from end_bits on reads as zero: keep bp from running away on a cut stream */
    const uint32_t cap = end_bits + 64u;
    s.bp = s.bp < cap ? s.bp : cap;
  }
  return bit;
}

/* stream bits the eight steps of one byte can consume at most, plus the two words a step looks at */
#define W3_LANE_FAST_MARGIN (8u * 31u + 64u)

/* decompress one block: stream sp[0..slen) -> dst[0..n).  slen * 8 + 1024 must fit in 32 bits. */
W3_HD void w3_lane_decode_block(const uint8_t *sp, uint32_t slen, uint32_t n, uint8_t *dst, uint32_t *cnt,
                                const w3_lane_tab *lt, const w3_hashtab *g, uint32_t ones) {
  const uint32_t sh = (uint32_t)((uintptr_t)sp & 3u);
  const uint32_t *ab = (const uint32_t *)(sp - sh);
  const uint32_t bp0 = sh * 8u, end_bits = bp0 + slen * 8u;
  w3_lane_dec s;
  s.x1 = 0u;
  s.x2 = 0xFFFFFFFFu;
  s.x = w3_lane_fresh<true>(ab, bp0, end_bits); /* new_decoder — arithmetic_coder.rs:75-78 */
  s.bp = bp0 + 32u;
  s.khi = s.klo = 0u;
  s.ci = W3_CTX0_RANKED;
  uint32_t word = 0;
  for (uint32_t i = 0; i < n; i++) {
    uint32_t byte = 0;
    if (s.bp + W3_LANE_FAST_MARGIN <= end_bits) {
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
      for (uint32_t a = 0; a < 8u; a++)
        byte = (byte << 1) | w3_lane_dec_step<false>(s, (a + 1u) & 7u, cnt, lt, g, ab, end_bits, ones);
    } else {
#if defined(__CUDA_ARCH__)
#pragma unroll 1
#endif
      for (uint32_t a = 0; a < 8u; a++)
        byte = (byte << 1) | w3_lane_dec_step<true>(s, (a + 1u) & 7u, cnt, lt, g, ab, end_bits, ones);
    }
    word |= byte << (8u * (i & 3u));
    if ((i & 3u) == 3u) {
      *(uint32_t *)(dst + (i - 3u)) = word;
      word = 0;
    }
  }
  for (uint32_t b = 0; b < (n & 3u); b++) dst[(n & ~3u) + b] = (uint8_t)(word >> (8u * b));
}

#if defined(__CUDACC__)
/* ------------------------------------------------------------------------ */
/* the kernel: blockDim.x / 32 warps of `lanes` active lanes; a CTA decodes   */
/* warps * lanes blocks per round                                             */

#define W3_LANE_MAX_WARPS 8u

__global__ void __launch_bounds__(W3_LANE_MAX_WARPS * 32u)
w3_decode_blocks_lane(const uint8_t *__restrict__ comp, const uint64_t *__restrict__ offsets, uint32_t n_blocks,
                      uint32_t block_size, uint64_t n_total, uint8_t *__restrict__ out, uint32_t *__restrict__ err,
                      const w3_hashtab *__restrict__ tab, uint32_t lanes) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  w3_lane_tab *lt = reinterpret_cast<w3_lane_tab *>(smem_raw);
  uint32_t *counts = reinterpret_cast<uint32_t *>(smem_raw + sizeof(w3_lane_tab));
  const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31u;
  const uint32_t per_cta = (blockDim.x >> 5) * lanes;
  const uint32_t slot = warp * lanes + lane;

  for (uint32_t i = threadIdx.x; i < 8u * W3_LANE_ROW; i += blockDim.x) lt->thi[i] = w3_lane_tab_thi(tab, i);
  for (uint32_t i = threadIdx.x; i < 8u * W3_LANE_NCELL; i += blockDim.x) lt->cellb[i] = w3_lane_tab_cellb(tab, i);
  uint32_t ones;
  asm volatile("mov.u32 %0, 0xFFFFFFFF;" : "=r"(ones)); /* opaque: ptxas would fold the constant and spend a move per use */

  for (uint32_t first = blockIdx.x * per_cta; first < n_blocks; first += gridDim.x * per_cta) {
    __syncthreads(); /* tables staged / the previous round's lanes are done with their counters */
    for (uint32_t i = threadIdx.x; i < per_cta * W3_NCTX; i += blockDim.x) counts[i] = 0u; /* ordern_entropy.rs:17 */
    __syncthreads();
    const uint32_t blk = first + slot;
    if (lane < lanes && blk < n_blocks) {
      const uint64_t off = (uint64_t)blk * block_size;
      const uint32_t n = (uint32_t)(n_total - off < block_size ? n_total - off : block_size);
      const uint64_t s0 = offsets[blk], slen = offsets[blk + 1] - s0;
      if (slen > 0x1FFFFF00ull) atomicOr(err, W3_ERR_ABORT); /* stream bits must fit 32 bits (the launcher checks too) */
      else w3_lane_decode_block(comp + s0, (uint32_t)slen, n, out + off, counts + slot * W3_NCTX, lt, tab, ones);
    }
  }
}
#endif
