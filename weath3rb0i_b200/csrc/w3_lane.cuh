/*
 * w3_lane.cuh — "lane" decode for the many-block regime: a block is decoded by
 * one lane (or a pair of lanes) of a warp instead of a whole warp or CTA
 * (1 GiB in 1 MiB blocks on one GPU is 7 blocks per SM; 256 KiB blocks, 28).
 *
 * decompress() of the reference (src/main.rs:131-140) is one dependent chain
 * per stream: bit -> history -> hash -> context -> probability -> bit.  With
 * one WARP per block every instruction of that chain is issued for 32 lanes
 * that all hold the same values (round 1: 250 warp-instructions per coded bit,
 * 678 cycles per bit at 7 blocks per SM, profiles/r01_1gib_*).  The integer
 * pipes of an SM sub-partition retire one warp-instruction per two cycles
 * whatever the number of active lanes, so the instruction, not the lane, is
 * the unit of cost: here every lane (pair) runs its own block, and one
 * instruction stream decodes up to 32 (16) blocks at once.
 *
 *   lane form   one lane per block; lanes never talk to each other.  Chosen
 *               when there are so many blocks per SM that shared memory
 *               (8 KiB of counters per resident block) is what limits.
 *   pair form   two lanes per block.  Both run the coder step on the same
 *               values; lane g additionally prepares context and probability
 *               of the NEXT position for the case that the bit being decoded
 *               is g, so the table lookup and Counter::p overlap the coder
 *               arithmetic instead of following it; one shuffle pair picks the
 *               survivor.  Chosen when an SM holds more than 8 blocks.
 *   split form  coder and model of a block in different warps: one coder
 *               warp (lane = block) and eight model warps (four lanes per
 *               block, guessing two bits ahead) per CTA of up to 8 blocks.
 *               Chosen when an SM holds 3 to 8 blocks (1 GiB in 1 MiB blocks).
 *               Described where it is defined, below.
 *
 * The per-lane code is plain sequential code — `__host__ __device__` — and
 * tests/hostsim runs exactly these statements against the oracle on the CPU
 * (the pair form with both lanes in lockstep and the pick done by hand, the
 * split form's warps as independent actors in a random order).
 *
 * Shared memory of a CTA: one copy of the lookup tables (w3_lane_tab, 72 KiB)
 * and 8 KiB of counters per resident block (Counter — src/models/counter.rs;
 * lane and pair form recompute the probability from the counts at predict
 * time — one reciprocal, fewer instructions than keeping a second table
 * current; the split form keeps that table, 8 KiB more per block, because
 * there the division would sit on the round trip between the warps).  HBM
 * sees the compressed streams (read once, through L1) and the output words.
 *
 * The entropy hash (ACHistory::hash — src/history/ac_history.rs:28-46) is the
 * rank of the bit-reversed history among <= 255 thresholds per alignment
 * (w3_core.cuh).  Lookup: a byte per (alignment, top 13 key bits) gives the
 * number of thresholds below the key's cell; the next W3_LANE_FIX thresholds
 * are loaded at once and compared on the top 31 bits of their high words by
 * subtraction (the sign bits add up to the rank; thresholds ascend, so those
 * past the cell compare greater by themselves).  A tie on those 31 bits, or
 * all W3_LANE_FIX thresholds below the key (0.18 % of lookups on the bench
 * corpus), sends that lane through w3_lane_rank_decl: the following
 * thresholds settle 93 % of those in shared memory, a real tie goes to the
 * general search (w3_rank_lookup, w3_achash_exact) over the tables in global
 * memory.
 */
#pragma once
#include <type_traits>

#include "w3_core.cuh"

#define W3_LANE_PB W3_PB2
#define W3_LANE_NCELL (1u << W3_LANE_PB)
#ifndef W3_LANE_FIX
#define W3_LANE_FIX 4
#endif
#ifndef W3_PAIR_EARLY_REPAIR
/* 1: repair a declined lookup where it is made (a branch inside B); 0: just before the pick, which
 * keeps B and the next F in one basic block.  Measured on 1 GiB / 1024 blocks
 * (profiles/r02_pair_variants.txt): early 1334 ms, late 1684 ms without the probability table;
 * with it 1671 / 1436.  The early form without the table is the default. */
#define W3_PAIR_EARLY_REPAIR 1
#endif
#ifndef W3_LANE_MINB
#define W3_LANE_MINB 1 /* __launch_bounds__ minimum CTAs per SM of the lane kernels (register budget) */
#endif
#define W3_LANE_ROW 264u /* 255 thresholds + padding: a lookup reads W3_LANE_FIX entries from any base <= 255 */
static_assert(255u + W3_LANE_FIX <= W3_LANE_ROW, "threshold rows must cover the read-ahead");

struct w3_lane_tab {
  uint32_t thi31[8 * W3_LANE_ROW];  /* thr[a][j] >> 33: the top 31 bits; padding 0x7FFFFFFF */
  uint8_t cellb[8 * W3_LANE_NCELL]; /* #{j : thr[a][j] < first key of the cell} */
  uint32_t ones;                    /* 0xFFFFFFFF, read back into a register ptxas cannot fold */
  uint32_t pad[3];
};
static_assert(sizeof(w3_lane_tab) % 16 == 0, "counters follow the tables at 16-byte alignment");

/* entry i of the staged tables, from the context's table in global memory */
W3_HD uint32_t w3_lane_tab_thi31(const w3_hashtab *g, uint32_t i) {
  const uint32_t a = i / W3_LANE_ROW, j = i % W3_LANE_ROW;
  return j < 255u ? (uint32_t)(g->thr[a][j] >> 33) : 0x7FFFFFFFu;
}
W3_HD uint8_t w3_lane_tab_cellb(const w3_hashtab *g, uint32_t i) {
  return (uint8_t)(g->cell2[i / W3_LANE_NCELL][i % W3_LANE_NCELL] & 0xFFu);
}

/* ---- shared-memory accessors ------------------------------------------------
 * On the device a block's counters and the tables are addressed by 32-bit
 * shared-window addresses held in registers: left to generic pointers the
 * compiler re-derives the window base (S2UR SR_CgaCtaId, ULEA) inside the
 * unrolled loop.  On the host the same type is a plain address. */
#if defined(__CUDA_ARCH__)
typedef uint32_t w3_sptr;
W3_HD uint32_t w3_s_ld32c(w3_sptr a) { /* constant tables: schedulable */
  uint32_t v;
  asm("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a));
  return v;
}
W3_HD uint32_t w3_s_ld8c(w3_sptr a) {
  uint32_t v;
  asm("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(a));
  return v;
}
W3_HD uint32_t w3_s_ld32(w3_sptr a) { /* counters: volatile asm statements keep their order */
  uint32_t v;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a));
  return v;
}
W3_HD void w3_s_st32(w3_sptr a, uint32_t v) { asm volatile("st.shared.u32 [%0], %1;" ::"r"(a), "r"(v)); }
/* ring words shared between warps: volatile, and a compiler barrier around them */
W3_HD void w3_s_ld64v(w3_sptr a, uint32_t &x, uint32_t &y) {
  asm volatile("ld.volatile.shared.v2.u32 {%0, %1}, [%2];" : "=r"(x), "=r"(y) : "r"(a) : "memory");
}
W3_HD void w3_s_st64v(w3_sptr a, uint32_t x, uint32_t y) {
  asm volatile("st.volatile.shared.v2.u32 [%0], {%1, %2};" ::"r"(a), "r"(x), "r"(y) : "memory");
}
W3_HD void w3_s_st32v(w3_sptr a, uint32_t v) { asm volatile("st.volatile.shared.u32 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }
W3_HD uint32_t w3_s_ld32v(w3_sptr a) {
  uint32_t v;
  asm volatile("ld.volatile.shared.u32 %0, [%1];" : "=r"(v) : "r"(a) : "memory");
  return v;
}
#else
typedef uintptr_t w3_sptr;
W3_HD uint32_t w3_s_ld32v(w3_sptr a) { return *(const volatile uint32_t *)a; }
W3_HD void w3_s_ld64v(w3_sptr a, uint32_t &x, uint32_t &y) { x = ((const volatile uint32_t *)a)[0]; y = ((const volatile uint32_t *)a)[1]; }
W3_HD void w3_s_st64v(w3_sptr a, uint32_t x, uint32_t y) { ((volatile uint32_t *)a)[0] = x; ((volatile uint32_t *)a)[1] = y; }
W3_HD void w3_s_st32v(w3_sptr a, uint32_t v) { *(volatile uint32_t *)a = v; }
W3_HD uint32_t w3_s_ld32c(w3_sptr a) { return *(const uint32_t *)a; }
W3_HD uint32_t w3_s_ld8c(w3_sptr a) { return *(const uint8_t *)a; }
W3_HD uint32_t w3_s_ld32(w3_sptr a) { return *(const uint32_t *)a; }
W3_HD void w3_s_st32(w3_sptr a, uint32_t v) { *(uint32_t *)a = v; }
#endif
/* what a lane needs of its CTA's shared memory */
struct w3_lane_mem {
  w3_sptr thi31, cellb; /* w3_lane_tab */
  w3_sptr cnt;          /* this block's 2048 counters */
  const w3_hashtab *g;  /* the full tables in global memory (general search) */
  uint32_t ones;
};

/* rank of a key among the thresholds of alignment a, from the top 31 bits of the key; kFix thresholds looked at */
template <int kFix>
W3_HD uint32_t w3_lane_rank_n(uint32_t khi, uint32_t a, const w3_lane_mem &m, bool &bad) {
  static_assert(255u + kFix <= W3_LANE_ROW, "threshold rows must cover the read-ahead");
  const uint32_t base = w3_s_ld8c(m.cellb + a * W3_LANE_NCELL + (khi >> (32u - W3_LANE_PB)));
  const w3_sptr t = m.thi31 + (a * W3_LANE_ROW + base) * 4u;
  uint32_t b[kFix];
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
  for (int i = 0; i < kFix; i++) b[i] = w3_s_ld32c(t + 4u * (uint32_t)i);
  /* z = key31 - threshold31 fits a signed word: > 0 the threshold is below the key, < 0 above, 0 undecided */
  const uint32_t k31 = khi >> 1;
  uint32_t above = 0;
  bool tie = false;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
  for (int i = 0; i < kFix; i++) {
    const uint32_t z = k31 - b[i];
    above += z >> 31;
    tie |= z == 0u;
  }
  bad = tie || above == 0u; /* none above: there may be more thresholds below the key than were looked at */
  return base + (uint32_t)kFix - above;
}
W3_HD uint32_t w3_lane_rank(uint32_t khi, uint32_t a, const w3_lane_mem &m, bool &bad) { return w3_lane_rank_n<W3_LANE_FIX>(khi, a, m, bad); }
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

/* what a lane does whose fixed lookup declined.  Measured on the bench corpus (scripts/exp/bad_split.cpp):
 * 0.18 % of lookups decline, 93 % of those because all W3_LANE_FIX thresholds lie below the key — the next
 * few settle it (rows end in padding that compares above every key) — and only a tie on the 31 bits needs
 * the 64-bit thresholds in global memory. */
#if defined(__CUDACC__)
__host__ __device__ __noinline__
#else
static
#endif
uint32_t w3_lane_rank_decl(uint32_t khi, uint32_t klo, uint32_t a, const w3_lane_mem &m) {
  const uint32_t base = w3_s_ld8c(m.cellb + a * W3_LANE_NCELL + (khi >> (32u - W3_LANE_PB)));
  const uint32_t k31 = khi >> 1;
  for (uint32_t j = base; j < 256u; j += (uint32_t)W3_LANE_FIX) {
    const w3_sptr t = m.thi31 + (a * W3_LANE_ROW + j) * 4u;
    uint32_t above = 0;
    bool tie = false;
    for (int i = 0; i < W3_LANE_FIX; i++) {
      const uint32_t z = k31 - w3_s_ld32c(t + 4u * (uint32_t)i);
      above += z >> 31;
      tie |= z == 0u;
    }
    if (tie) break;
    if (above) return j + (uint32_t)W3_LANE_FIX - above;
  }
  return w3_lane_rank_slow(khi, klo, a, m.g);
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

/* the coder's half of a decode step: narrow by the bit, both renormalisation loops in closed
 * form (arithmetic_coder.rs:85-102; of the E3 flips of x only the last survives the shifts) */
struct w3_lane_coder {
  uint32_t x1, x2, x; /* ArithmeticCoder — arithmetic_coder.rs:11-18 */
  uint32_t bp;        /* stream bits consumed, counted from the aligned base */
};
/* first half: the bit (all a step's consumers wait for) */
W3_HD uint32_t w3_lane_coder_bit(const w3_lane_coder &c, uint32_t P, uint32_t &m) {
  m = w3_umulhi(c.x2 - c.x1, P);                        /* lerp — arithmetic_coder.rs:81 */
  return (c.x - c.x1) <= m ? 1u : 0u;                   /* :82; x >= x1 always */
}
/* second half: narrow and renormalise */
template <bool kCareful>
W3_HD void w3_lane_coder_shift(w3_lane_coder &c, uint32_t bit, uint32_t m, const uint32_t *ab, uint32_t end_bits, uint32_t ones);
template <bool kCareful>
W3_HD uint32_t w3_lane_coder_step(w3_lane_coder &c, uint32_t P, const uint32_t *ab, uint32_t end_bits, uint32_t ones) {
  uint32_t m;
  const uint32_t bit = w3_lane_coder_bit(c, P, m);
  w3_lane_coder_shift<kCareful>(c, bit, m, ab, end_bits, ones);
  return bit;
}
template <bool kCareful>
W3_HD void w3_lane_coder_shift(w3_lane_coder &c, uint32_t bit, uint32_t m, const uint32_t *ab, uint32_t end_bits, uint32_t ones) {
  if (bit) c.x2 = c.x1 + m;                             /* :85-88 */
  else c.x1 = c.x1 + m + 1u;
  const uint32_t fresh = w3_lane_fresh<kCareful>(ab, c.bp, end_bits);
  int sh_s, sh_k;
  const int n = w3_ac_renorm_ones(c.x1, c.x2, sh_s, sh_k, ones);
  c.x = w3_fsl(fresh, c.x, (uint32_t)n) ^ ((0u - (uint32_t)sh_k) & 0x80000000u);
  c.bp += (uint32_t)n;
  if (kCareful) { /* everything from end_bits on reads as zero: keep bp from running away on a cut stream */
    const uint32_t cap = end_bits + 64u;
    c.bp = c.bp < cap ? c.bp : cap;
  }
}
W3_HD void w3_lane_coder_init(w3_lane_coder &c, const uint32_t *ab, uint32_t bp0, uint32_t end_bits) {
  c.x1 = 0u;
  c.x2 = 0xFFFFFFFFu;
  c.x = w3_lane_fresh<true>(ab, bp0, end_bits); /* new_decoder — arithmetic_coder.rs:75-78 */
  c.bp = bp0 + 32u;
}

/* Counter::p << 16 of a counter word (c0 | c1 << 16) */
W3_HD uint32_t w3_lane_P(uint32_t c) { return w3_counter_p1(c & 0xFFFFu, c >> 16) << 16; }

/* stream bits the eight steps of one byte can consume at most, plus the two words a step looks at */
#define W3_LANE_FAST_MARGIN (8u * 31u + 64u)

/* ---- lane form --------------------------------------------------------------- */
struct w3_lane_dec {
  w3_lane_coder c;
  uint32_t khi, klo; /* ACHistory.bits, newest bit at bit 63 (w3_key_of_history) */
  uint32_t ci;       /* byte offset of the current context's counter: ((rank << 3) | alignment) * 4 */
};

/* one pass of the decompress loop (main.rs:134-137) for the bit whose successor has alignment a1 */
template <bool kCareful>
W3_HD uint32_t w3_lane_dec_step(w3_lane_dec &s, uint32_t a1, const w3_lane_mem &m, const uint32_t *ab, uint32_t end_bits) {
  const uint32_t cw = w3_s_ld32(m.cnt + s.ci);
  const uint32_t bit = w3_lane_coder_step<kCareful>(s.c, w3_lane_P(cw), ab, end_bits, m.ones); /* predict + decode — main.rs:134-135 */
  w3_s_st32(m.cnt + s.ci, w3_counter_update(cw, bit)); /* adapt — main.rs:136, ordern_entropy.rs:32-34 */
  s.klo = w3_fsr(s.klo, s.khi, 1u);                    /* history — ac_history.rs:23-26 */
  s.khi = (s.khi >> 1) | (bit << 31);
  bool bad;
  uint32_t r = w3_lane_rank(s.khi, a1, m, bad);        /* hash — ordern_entropy.rs:36-45 */
  if (bad) r = w3_lane_rank_decl(s.khi, s.klo, a1, m);
  s.ci = ((r << W3_ALIGN_BITS) | a1) * 4u;
  return bit;
}

/* decompress one block: stream sp[0..slen) -> dst[0..n).  slen * 8 + 1024 must fit in 32 bits. */
W3_HD void w3_lane_decode_block(const uint8_t *sp, uint32_t slen, uint32_t n, uint8_t *dst, const w3_lane_mem &m) {
  const uint32_t sh = (uint32_t)((uintptr_t)sp & 3u);
  const uint32_t *ab = (const uint32_t *)(sp - sh);
  const uint32_t bp0 = sh * 8u, end_bits = bp0 + slen * 8u;
  w3_lane_dec s;
  w3_lane_coder_init(s.c, ab, bp0, end_bits);
  s.khi = s.klo = 0u;
  s.ci = W3_CTX0_RANKED * 4u;
  uint32_t word = 0;
  for (uint32_t i = 0; i < n; i++) {
    uint32_t byte = 0;
    if (s.c.bp + W3_LANE_FAST_MARGIN <= end_bits) {
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
      for (uint32_t a = 0; a < 8u; a++) byte = (byte << 1) | w3_lane_dec_step<false>(s, (a + 1u) & 7u, m, ab, end_bits);
    } else {
#if defined(__CUDA_ARCH__)
#pragma unroll 1
#endif
      for (uint32_t a = 0; a < 8u; a++) byte = (byte << 1) | w3_lane_dec_step<true>(s, (a + 1u) & 7u, m, ab, end_bits);
    }
    word |= byte << (8u * (i & 3u));
    if ((i & 3u) == 3u) {
      *(uint32_t *)(dst + (i - 3u)) = word;
      word = 0;
    }
  }
  for (uint32_t b = 0; b < (n & 3u); b++) dst[(n & ~3u) + b] = (uint8_t)(word >> (8u * b));
}

/* ---- pair form ----------------------------------------------------------------
 * Per bit t:
 *   F     (both lanes) decode bit t with the probability picked last time, adapt its
 *         counter, shift the bit into the key;
 *   pick  one shuffle pair takes the candidate (context, probability) of position t+1
 *         from the lane whose guess of bit t came true;
 *   B     (lane g) candidate for position t+2 for the case bit t+1 = g: table lookup on
 *         the key with g shifted in, then the probability of that context.
 * B of bit t and F of bit t+1 are independent and sit in ONE basic block, so the scheduler
 * interleaves the two chains (a lone warp issues in order: with a branch between them a step
 * costs the sum of the chains).  Hence the rare lane whose fixed lookup declined is repaired
 * late — just before its candidate is picked, behind one warp-uniform branch.
 * kCacheP: a second table per block holds Counter::p << 16 of every counter, kept current by
 * F (off the chain), so that B ends with one load instead of a division: the chain from
 * "bit known" to "candidate ready" is what a step costs when few blocks share an SM.
 * Contexts of consecutive positions differ in their alignment bits, so B never touches a
 * counter F updates in the same or the next step.  tests/hostsim runs both lanes in
 * lockstep on the CPU. */
struct w3_pair_dec {
  w3_lane_coder c;
  uint32_t khi, klo;
  uint32_t ci, P; /* counter offset and Counter::p << 16 of the bit being decoded (both lanes agree) */
};
struct w3_pair_cand {
  uint32_t ci, P;   /* this lane's candidate for the position after next: valid if the next bit equals its guess */
  uint32_t gk, glo; /* the key it was looked up with */
  bool bad;         /* the fixed lookup declined: w3_pair_repair before the pick */
};
#define W3_LANE_P_OFF (W3_NCTX * 4u) /* kCacheP: byte offset of the probability table behind a block's counters */

template <bool kCacheP>
W3_HD uint32_t w3_pair_P_of(uint32_t ci, const w3_lane_mem &m) {
  return kCacheP ? w3_s_ld32(m.cnt + W3_LANE_P_OFF + ci) : w3_lane_P(w3_s_ld32(m.cnt + ci));
}
/* B: this lane's candidate for the position of alignment a2, assuming the next bit to be decoded is g */
template <bool kCacheP>
W3_HD w3_pair_cand w3_key_guess(uint32_t khi, uint32_t klo, uint32_t g, uint32_t a2, const w3_lane_mem &m);
template <bool kCacheP>
W3_HD w3_pair_cand w3_pair_guess(const w3_pair_dec &s, uint32_t g, uint32_t a2, const w3_lane_mem &m) {
  return w3_key_guess<kCacheP>(s.khi, s.klo, g, a2, m);
}
template <bool kCacheP>
W3_HD w3_pair_cand w3_key_guess(uint32_t khi, uint32_t klo, uint32_t g, uint32_t a2, const w3_lane_mem &m) {
  w3_pair_cand c;
  c.gk = (khi >> 1) | (g << 31);
  c.glo = w3_fsr(klo, khi, 1u);
  const uint32_t r = w3_lane_rank(c.gk, a2, m, c.bad);
  c.ci = (((r & 255u) << W3_ALIGN_BITS) | a2) * 4u; /* r <= 255 unless bad: the load below stays inside the table */
#if W3_PAIR_EARLY_REPAIR
  if (c.bad) {
    const uint32_t rr = w3_lane_rank_decl(c.gk, c.glo, a2, m);
    c.ci = ((rr << W3_ALIGN_BITS) | a2) * 4u;
    c.bad = false;
  }
#endif
  c.P = w3_pair_P_of<kCacheP>(c.ci, m);             /* of a wrong counter if bad: repaired before the pick */
  return c;
}
/* the general search for a candidate the fixed lookup declined */
template <bool kCacheP>
W3_HD void w3_pair_repair(w3_pair_cand &c, uint32_t a2, const w3_lane_mem &m) {
  const uint32_t r = w3_lane_rank_decl(c.gk, c.glo, a2, m);
  c.ci = ((r << W3_ALIGN_BITS) | a2) * 4u;
  c.P = w3_pair_P_of<kCacheP>(c.ci, m);
  c.bad = false;
}
/* F: decode the current bit (main.rs:134-137 without the hash), both lanes alike */
template <bool kCareful, bool kCacheP>
W3_HD uint32_t w3_pair_front(w3_pair_dec &s, const w3_lane_mem &m, const uint32_t *ab, uint32_t end_bits) {
  const uint32_t bit = w3_lane_coder_step<kCareful>(s.c, s.P, ab, end_bits, m.ones);
  const uint32_t nv = w3_counter_update(w3_s_ld32(m.cnt + s.ci), bit);
  w3_s_st32(m.cnt + s.ci, nv); /* both lanes store the same words */
  if (kCacheP) w3_s_st32(m.cnt + W3_LANE_P_OFF + s.ci, w3_lane_P(nv));
  s.klo = w3_fsr(s.klo, s.khi, 1u);
  s.khi = (s.khi >> 1) | (bit << 31);
  return bit;
}
W3_HD void w3_pair_init(w3_pair_dec &s, const uint32_t *ab, uint32_t bp0, uint32_t end_bits) {
  w3_lane_coder_init(s.c, ab, bp0, end_bits);
  s.khi = s.klo = 0u;
  s.ci = W3_CTX0_RANKED * 4u;
  s.P = 0x80000000u; /* Counter::new: p = 32768 */
}


/* ---- split form ----------------------------------------------------------------
 * The pair form runs coder and model of a block in ONE warp: 109 instructions per bit, and a
 * lone warp issues one instruction per ~3 cycles whatever its lanes hold, so 330 cycles per bit.
 * Here the two halves run in different warps, on different schedulers:
 *
 *   coder warp    lane b = block b of the CTA's group (<= 8).  Per bit: wait for the probability
 *                 of the position, decide the bit, publish it, narrow and renormalise.  No table,
 *                 no counter: 51 instructions.
 *   model warps   eight; lanes 4b .. 4b + 3 = block b, lane & 3 = g, a guess of the two bits
 *                 before the position.  Warp W owns the positions q = W (mod 8) — alignment W:
 *                 their lookups, their probabilities and their counters, so no counter crosses a
 *                 warp.  For its position q it
 *                   adapts   the counter of its previous position q-8 as soon as bit q-8 is out
 *                            (the lane that guessed right is known by then), and the probability
 *                            kept next to it (Counter::p << 16: one load instead of a division
 *                            when it is needed);
 *                   prepares once bit q-4 is out: context and probability of position q for bits
 *                            (q-2, q-1) = g and for bit q-3 = 0 and = 1 — two lookups whose loads
 *                            interleave;
 *                   finishes once bit q-3 is out: picks one of the two and publishes it.
 *
 * What crosses warps are single words in shared memory that carry their own sequence numbers, so
 * nothing depends on the order in which different words become visible:
 *   pub[t & 7][b]      = {the block's last 32 decoded bits (newest highest), t} after t bits
 *   cand[q & 7][b][g]  = Counter::p << 16 | (q + 1) & 0xFFFF
 *
 * Why this shape (profiles/r02_split_variants.txt … r02_spin_variants.txt, in the order tried):
 * one guessed bit and one lookup after the bit arrives is no faster than the pair form — a step
 * then costs half the round trip "bit published -> seen by the model (~145 cycles,
 * r02_detect_latency.txt) -> lookup (~250: four dependent shared-memory loads) -> candidate
 * published -> seen by the coder".  Two guessed bits spread the round trip over three steps
 * (1174 ms for 1 GiB against the pair form's 1331); doing the lookups for both values of the one
 * bit still missing BEFORE it arrives leaves a load, a select and a store on the round trip
 * (1039 ms; 1013 without a stream prefetch that bought nothing; 952 with W3_SPLIT_FIX = 6): the
 * coder warp then waits in one step out of eight and its own ~200 cycles per step are what is
 * left.  Branch-free "attempts" per lane instead of a lockstep wait (the team
 * decoder's coder, w3_team.cuh) were tried too: 160-180 cycles per attempt but 1.5-1.8 attempts
 * per bit, because an attempt that just misses a candidate costs a whole attempt where a poll
 * costs 45 cycles (1193-1292 ms).
 *
 * The model's steps are `__host__ __device__`; tests/hostsim runs coder and models as independent
 * actors in a random order, which the sequence numbers must make irrelevant. */
#ifndef W3_SPLIT_FIX
/* thresholds per lookup.  The pair form does best with 4 (its cost is its instruction count); here a lane that
 * declines delays its whole warp's candidates and the coder warp polls: 6 measured 953 ms against 1013 with 4 and
 * 959 with 8 (profiles/r02_split_fix_variants.txt) */
#define W3_SPLIT_FIX 6
#endif
#define W3_SPLIT_SLOTS 8u
#define W3_SPLIT_MAXB 8u   /* blocks per CTA: 32 model lanes / 4 guesses */
#define W3_SPLIT_MODELS 8u /* model warps = alignments */
struct w3_split_ring {
  uint32_t cand[W3_SPLIT_SLOTS][W3_SPLIT_MAXB][4];
  uint32_t pub[W3_SPLIT_SLOTS][W3_SPLIT_MAXB][2];
};
static_assert(sizeof(w3_split_ring) % 16 == 0, "ring keeps the counters aligned");
W3_HD uint32_t w3_split_tag(uint32_t q) { return (q + 1u) & 0xFFFFu; } /* not 0 before q = 65535: a cleared ring matches nothing, and by then every slot has been written */

/* the coder's step for position q once its candidate word cw is there; cand/pub: the block's columns */
template <bool kCareful>
W3_HD uint32_t w3_split_coder_step(w3_lane_coder &c, uint32_t &rhist, uint32_t q, uint32_t cw, w3_sptr pub_blk,
                                   const uint32_t *ab, uint32_t end_bits, uint32_t ones) {
  uint32_t mm;
  const uint32_t bit = w3_lane_coder_bit(c, cw & 0xFFFF0000u, mm); /* predict + decode — main.rs:134-135 */
  rhist = w3_fsr(rhist, bit, 1u);
  w3_s_st64v(pub_blk + ((q + 1u) & 7u) * 64u, rhist, q + 1u);
  w3_lane_coder_shift<kCareful>(c, bit, mm, ab, end_bits, ones);
  return bit;
}
/* where the coder finds the candidate of position q: the guess that came true is the top two bits of its history */
W3_HD w3_sptr w3_split_cand_at(w3_sptr cand_blk, uint32_t q, uint32_t rhist) { return cand_blk + (q & 7u) * 128u + (rhist >> 30) * 4u; }

struct w3_split_mod {
  uint32_t khi_prev, klo;  /* the history key of the warp's previous position (the general search wants 64 bits) */
  uint32_t ci;             /* counter offset of this lane's last candidate */
  uint32_t ci2[2], pw2[2]; /* prepared: counter offset and Counter::p << 16 of the candidate, for either value of the bit still awaited */
};
W3_HD void w3_split_mod_init(w3_split_mod &s) {
  s.khi_prev = s.klo = 0u;
  s.ci = W3_CTX0_RANKED * 4u;
  s.ci2[0] = s.ci2[1] = s.pw2[0] = s.pw2[1] = 0u;
}
/* Counter::update of the context that coded `bit` (ordern_entropy.rs:32-34), and its probability */
W3_HD void w3_split_adapt(uint32_t ci, uint32_t bit, const w3_lane_mem &m) {
  const uint32_t nv = w3_counter_update(w3_s_ld32(m.cnt + ci), bit);
  w3_s_st32(m.cnt + ci, nv);
  w3_s_st32(m.cnt + W3_LANE_P_OFF + ci, w3_lane_P(nv));
}
/* prepare position q >= 1 (alignment q & 7) for this lane's guess g of bits (q-2, q-1): rh3 = the block's decoded
 * bits through q-4, newest highest.  Both lookups are one straight-line sequence; the rare lane whose fixed lookup
 * declined is repaired after. */
W3_HD void w3_split_prepare(w3_split_mod &s, uint32_t rh3, uint32_t q, uint32_t g, const w3_lane_mem &m) {
  const uint32_t a = q & 7u;
  const uint32_t gk0 = (rh3 >> 3) | (g << 30), gk1 = gk0 | (1u << 29);
  bool bad0, bad1;
  uint32_t r0 = w3_lane_rank_n<W3_SPLIT_FIX>(gk0, a, m, bad0) & 255u, r1 = w3_lane_rank_n<W3_SPLIT_FIX>(gk1, a, m, bad1) & 255u; /* hash — ordern_entropy.rs:36-45 */
  if (bad0 | bad1) {
    const uint32_t klo = w3_fsr(s.klo, s.khi_prev, W3_SPLIT_MODELS); /* the low word once bit q-3 is in (its top bits are bits of khi_prev either way) */
    const uint32_t rh0 = rh3 >> 1;
    if (bad0) r0 = w3_lane_rank_decl(gk0, w3_fsr(klo, rh0, 2u), a, m);
    if (bad1) r1 = w3_lane_rank_decl(gk1, w3_fsr(klo, rh0 | 0x80000000u, 2u), a, m);
  }
  s.ci2[0] = ((r0 << W3_ALIGN_BITS) | a) * 4u;
  s.ci2[1] = ((r1 << W3_ALIGN_BITS) | a) * 4u;
  s.pw2[0] = w3_s_ld32(m.cnt + W3_LANE_P_OFF + s.ci2[0]);
  s.pw2[1] = w3_s_ld32(m.cnt + W3_LANE_P_OFF + s.ci2[1]);
}
/* finish position q: rh = the bits through q-3.  Publishes unless !valid (a position past the block's end). */
W3_HD void w3_split_finish(w3_split_mod &s, uint32_t rh, uint32_t q, uint32_t g, uint32_t blk, bool valid, w3_sptr cand_base) {
  uint32_t P = 0x80000000u; /* position 0: OrderNEntropy starts in context 0 with a new counter */
  uint32_t ci = W3_CTX0_RANKED * 4u;
  if (q != 0u) {
    const uint32_t b = rh >> 31;
    ci = b ? s.ci2[1] : s.ci2[0];
    P = b ? s.pw2[1] : s.pw2[0];
    s.klo = w3_fsr(s.klo, s.khi_prev, W3_SPLIT_MODELS); /* ac_history.rs:23-26, eight bits on */
    s.khi_prev = rh;
  }
  s.ci = ci;
  if (valid) w3_s_st32v(cand_base + (((q & 7u) * W3_SPLIT_MAXB + blk) * 4u + g) * 4u, P | w3_split_tag(q));
}

#if defined(__CUDACC__)
/* ------------------------------------------------------------------------ */
/* kernels: blockDim.x / 32 warps; a CTA decodes warps * lanes blocks per     */
/* round (lanes = blocks per warp)                                            */

#define W3_LANE_MAX_WARPS 8u

__device__ __forceinline__ void w3_lane_stage(w3_lane_tab *lt, const w3_hashtab *tab) {
  for (uint32_t i = threadIdx.x; i < 8u * W3_LANE_ROW; i += blockDim.x) lt->thi31[i] = w3_lane_tab_thi31(tab, i);
  for (uint32_t i = threadIdx.x; i < 8u * W3_LANE_NCELL; i += blockDim.x) lt->cellb[i] = w3_lane_tab_cellb(tab, i);
  if (threadIdx.x == 0) lt->ones = 0xFFFFFFFFu;
}
__device__ __forceinline__ w3_lane_mem w3_lane_mem_of(w3_lane_tab *lt, uint32_t *cnt, const w3_hashtab *tab) {
  w3_lane_mem m;
  m.thi31 = (w3_sptr)__cvta_generic_to_shared(lt->thi31);
  m.cellb = (w3_sptr)__cvta_generic_to_shared(lt->cellb);
  m.cnt = (w3_sptr)__cvta_generic_to_shared(cnt);
  m.g = tab;
  asm volatile("ld.volatile.shared.u32 %0, [%1];" : "=r"(m.ones) : "r"((uint32_t)__cvta_generic_to_shared(&lt->ones)));
  return m;
}

__global__ void __launch_bounds__(W3_LANE_MAX_WARPS * 32u, W3_LANE_MINB)
w3_decode_blocks_lane(const uint8_t *__restrict__ comp, const uint64_t *__restrict__ offsets, uint32_t n_blocks,
                      uint32_t block_size, uint64_t n_total, uint8_t *__restrict__ out, uint32_t *__restrict__ err,
                      const w3_hashtab *__restrict__ tab, uint32_t lanes) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  w3_lane_tab *lt = reinterpret_cast<w3_lane_tab *>(smem_raw);
  uint32_t *counts = reinterpret_cast<uint32_t *>(smem_raw + sizeof(w3_lane_tab));
  const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31u;
  const uint32_t per_cta = (blockDim.x >> 5) * lanes;
  const uint32_t slot = warp * lanes + lane;

  w3_lane_stage(lt, tab);
  for (uint32_t first = blockIdx.x * per_cta; first < n_blocks; first += gridDim.x * per_cta) {
    __syncthreads(); /* tables staged / the previous round's lanes are done with their counters */
    for (uint32_t i = threadIdx.x; i < per_cta * W3_NCTX; i += blockDim.x) counts[i] = 0u; /* ordern_entropy.rs:17 */
    __syncthreads();
    const uint32_t blk = first + slot;
    if (lane < lanes && blk < n_blocks) {
      const uint64_t off = (uint64_t)blk * block_size;
      const uint32_t n = (uint32_t)(n_total - off < block_size ? n_total - off : block_size);
      const uint64_t s0 = offsets[blk], slen = offsets[blk + 1] - s0;
      if (slen > 0x1FFFFF00ull) atomicOr(err, W3_ERR_ABORT); /* stream bits must fit 32 bits (the host checks too) */
      else w3_lane_decode_block(comp + s0, (uint32_t)slen, n, out + off, w3_lane_mem_of(lt, counts + slot * W3_NCTX, tab));
    }
  }
}

/* Pair form: lanes 2p and 2p+1 of a warp decode block p of the warp together.  The pick is a
 * full-warp shuffle, so the warp's control flow is kept uniform: every lane runs block_size
 * bytes' worth of steps; pairs beyond `pairs` and beyond the last block shadow a real pair of
 * the warp (same values, same stores), a short last block keeps stepping over its exhausted
 * stream without storing. */
template <bool kCacheP>
__global__ void __launch_bounds__(W3_LANE_MAX_WARPS * 32u, W3_LANE_MINB)
w3_decode_blocks_pair(const uint8_t *__restrict__ comp, const uint64_t *__restrict__ offsets, uint32_t n_blocks,
                      uint32_t block_size, uint64_t n_total, uint8_t *__restrict__ out, uint32_t *__restrict__ err,
                      const w3_hashtab *__restrict__ tab, uint32_t pairs) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  w3_lane_tab *lt = reinterpret_cast<w3_lane_tab *>(smem_raw);
  uint32_t *counts = reinterpret_cast<uint32_t *>(smem_raw + sizeof(w3_lane_tab));
  const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31u;
  const uint32_t per_cta = (blockDim.x >> 5) * pairs;
  const uint32_t g = lane & 1u;
  const uint32_t words = kCacheP ? 2u * W3_NCTX : W3_NCTX; /* per block: counters (+ probabilities) */

  w3_lane_stage(lt, tab);
  for (uint32_t first = blockIdx.x * per_cta; first < n_blocks; first += gridDim.x * per_cta) {
    __syncthreads();
    for (uint32_t i = threadIdx.x; i < per_cta * words; i += blockDim.x)
      counts[i] = (kCacheP && (i % words) >= W3_NCTX) ? 0x80000000u : 0u; /* Counter::new: counts 0, p = 32768 */
    __syncthreads();
    const uint32_t wfirst = first + warp * pairs; /* this warp's first block */
    if (wfirst >= n_blocks) continue;              /* warp-uniform */
    const uint32_t in_warp = n_blocks - wfirst < pairs ? n_blocks - wfirst : pairs;
    const uint32_t pair = (lane >> 1) % in_warp;   /* pairs without a block of their own shadow one that has */
    const bool owner = g == 0u && (lane >> 1) < in_warp;
    const uint32_t blk = wfirst + pair;
    const uint64_t off = (uint64_t)blk * block_size;
    const uint32_t n = (uint32_t)(n_total - off < block_size ? n_total - off : block_size);
    const uint64_t s0 = offsets[blk];
    uint64_t slen = offsets[blk + 1] - s0;
    if (slen > 0x1FFFFF00ull) { atomicOr(err, W3_ERR_ABORT); slen = 0; } /* decodes zeros, flagged */
    const uint8_t *sp = comp + s0;
    uint8_t *dst = out + off;
    const uint32_t sh = (uint32_t)((uintptr_t)sp & 3u);
    const uint32_t *ab = (const uint32_t *)(sp - sh);
    const uint32_t bp0 = sh * 8u, end_bits = bp0 + (uint32_t)slen * 8u;
    const w3_lane_mem m = w3_lane_mem_of(lt, counts + (warp * pairs + pair) * words, tab);
    w3_pair_dec s;
    w3_pair_init(s, ab, bp0, end_bits);
    w3_pair_cand cand = w3_pair_guess<kCacheP>(s, g, 1u, m); /* position 1, for either value of bit 0 */
    uint32_t word = 0;
    const uint32_t n_steps = block_size < n_total ? block_size : (uint32_t)n_total; /* the same for every lane */
    auto step = [&](auto careful, uint32_t a) -> uint32_t {
      const uint32_t bit = w3_pair_front<decltype(careful)::value, kCacheP>(s, m, ab, end_bits); /* F */
#if !W3_PAIR_EARLY_REPAIR
      if (__any_sync(0xFFFFFFFFu, cand.bad)) {
        if (cand.bad) w3_pair_repair<kCacheP>(cand, (a + 1u) & 7u, m);
      }
#endif
      const uint32_t src = (lane & 30u) | bit;                                                   /* pick */
      s.ci = __shfl_sync(0xFFFFFFFFu, cand.ci, src);
      s.P = __shfl_sync(0xFFFFFFFFu, cand.P, src);
      cand = w3_pair_guess<kCacheP>(s, g, (a + 2u) & 7u, m);                                     /* B */
      return bit;
    };
    for (uint32_t i = 0; i < n_steps; i++) {
      uint32_t byte = 0;
      if (__all_sync(0xFFFFFFFFu, s.c.bp + W3_LANE_FAST_MARGIN <= end_bits)) {
#pragma unroll
        for (uint32_t a = 0; a < 8u; a++) byte = (byte << 1) | step(std::false_type(), a);
      } else {
#pragma unroll 1
        for (uint32_t a = 0; a < 8u; a++) byte = (byte << 1) | step(std::true_type(), a);
      }
      word |= byte << (8u * (i & 3u));
      if ((i & 3u) == 3u) {
        if (owner && i < n) *(uint32_t *)(dst + (i - 3u)) = word;
        word = 0;
      }
      if (owner && i + 1u == n)
        for (uint32_t b = 0; b < (n & 3u); b++) dst[(n & ~3u) + b] = (uint8_t)(word >> (8u * b));
    }
  }
}

/* Split form.  Warp 0: the coder, in lockstep over the group's blocks (lane b = block b % in_group: lanes beyond
 * the group shadow real ones — same values, same stores); warps 1, 2, 3, 5, 6, 7, 9, 10: the models of alignments
 * 0 .. 7; warps 4 and 8 idle so that the coder warp keeps its scheduler to itself.  Every warp runs block_size
 * bytes' worth of positions; a short last block keeps stepping over its exhausted stream without storing.  A wait
 * that does not end (the peer is gone) gives up after W3_SPLIT_SPIN_LIMIT polls, flags W3_ERR_ABORT and stops
 * waiting for good, so the kernel always terminates. */
#define W3_SPLIT_THREADS 352u
#define W3_SPLIT_WORDS (2u * W3_NCTX) /* per block: counters + probabilities */
#ifndef W3_SPLIT_SPIN_LIMIT
#define W3_SPLIT_SPIN_LIMIT (1u << 20)
#endif

/* the word at `a` once (word & mask) == want; budget = polls left before giving up (0 afterwards) */
__device__ __forceinline__ uint32_t w3_split_wait32(uint32_t a, uint32_t mask, uint32_t want, uint32_t &budget) {
  uint32_t w = w3_s_ld32v(a);
  if ((w & mask) != want) {
    uint32_t spins = 0;
    do w = w3_s_ld32v(a);
    while ((w & mask) != want && ++spins < budget);
    if ((w & mask) != want) budget = 0u;
  }
  return w;
}
/* the block's decoded bits once `count` of them are published */
__device__ __forceinline__ uint32_t w3_split_wait_pub(uint32_t pub_blk, uint32_t count, uint32_t &budget) {
  uint32_t rh, t;
  const uint32_t a = pub_blk + (count & 7u) * 64u;
  w3_s_ld64v(a, rh, t);
  if (t != count) {
    uint32_t spins = 0;
    do w3_s_ld64v(a, rh, t);
    while (t != count && ++spins < budget);
    if (t != count) budget = 0u;
  }
  return rh;
}

__global__ void __launch_bounds__(W3_SPLIT_THREADS, 1)
w3_decode_blocks_split(const uint8_t *__restrict__ comp, const uint64_t *__restrict__ offsets, uint32_t n_blocks,
                       uint32_t block_size, uint64_t n_total, uint8_t *__restrict__ out, uint32_t *__restrict__ err,
                       const w3_hashtab *__restrict__ tab, uint32_t group) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  w3_lane_tab *lt = reinterpret_cast<w3_lane_tab *>(smem_raw);
  w3_split_ring *ring = reinterpret_cast<w3_split_ring *>(smem_raw + sizeof(w3_lane_tab));
  uint32_t *counts = reinterpret_cast<uint32_t *>(smem_raw + sizeof(w3_lane_tab) + sizeof(w3_split_ring));
  const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31u;
  const uint32_t cand_base = (uint32_t)__cvta_generic_to_shared(&ring->cand[0][0][0]);
  const uint32_t pub_base = (uint32_t)__cvta_generic_to_shared(&ring->pub[0][0][0]);
  uint32_t budget = W3_SPLIT_SPIN_LIMIT;

  w3_lane_stage(lt, tab);
  for (uint32_t first = blockIdx.x * group; first < n_blocks; first += gridDim.x * group) {
    __syncthreads();
    for (uint32_t i = threadIdx.x; i < group * W3_SPLIT_WORDS; i += blockDim.x)
      counts[i] = (i % W3_SPLIT_WORDS) >= W3_NCTX ? 0x80000000u : 0u; /* Counter::new — ordern_entropy.rs:17: counts 0, p = 32768 */
    for (uint32_t i = threadIdx.x; i < sizeof(w3_split_ring) / 4u; i += blockDim.x) reinterpret_cast<uint32_t *>(ring)[i] = 0u; /* pub[0] = {0, 0}: no bits yet */
    __syncthreads();
    const uint32_t in_group = n_blocks - first < group ? n_blocks - first : group;
    const uint32_t n_steps = block_size < n_total ? block_size : (uint32_t)n_total; /* the same for every warp */
    if (warp != 0u && (warp & 3u) == 0u) continue; /* warps 4, 8: see above */

    if (warp != 0u) { /* ---- model warp of alignment W ---- */
      const uint32_t W = warp - 1u - (warp >> 2);
      const uint32_t b = (lane >> 2) % in_group, g = lane & 3u;
      const w3_lane_mem m = w3_lane_mem_of(lt, counts + b * W3_SPLIT_WORDS, tab);
      const uint32_t pub_blk = pub_base + 8u * b;
      w3_split_mod s;
      w3_split_mod_init(s);
      for (uint32_t i = 0; i < n_steps; i++) {
        const uint32_t q = 8u * i + W;
        if (q >= 8u) { /* adapt q-8: its bit is the newest of the first q-7, the guess that came true the two before it */
          const uint32_t rh = w3_split_wait_pub(pub_blk, q - 7u, budget);
          const uint32_t ci_old = __shfl_sync(0xFFFFFFFFu, s.ci, (lane & 28u) | ((rh >> 29) & 3u));
          w3_split_adapt(ci_old, rh >> 31, m);
        }
        if (q != 0u) w3_split_prepare(s, q >= 3u ? w3_split_wait_pub(pub_blk, q - 3u, budget) : 0u, q, g, m);
        w3_split_finish(s, q >= 2u ? w3_split_wait_pub(pub_blk, q - 2u, budget) : 0u, q, g, b, true, cand_base);
      }
      continue;
    }

    /* ---- the coder warp ---- */
    const uint32_t b = lane % in_group;
    const bool owner = lane < in_group;
    const uint32_t blk = first + b;
    const uint64_t off = (uint64_t)blk * block_size;
    const uint32_t n = (uint32_t)(n_total - off < block_size ? n_total - off : block_size);
    const uint64_t s0 = offsets[blk];
    uint64_t slen = offsets[blk + 1] - s0;
    if (slen > 0x1FFFFF00ull) { atomicOr(err, W3_ERR_ABORT); slen = 0; } /* decodes zeros, flagged */
    const uint8_t *sp = comp + s0;
    uint8_t *dst = out + off;
    const uint32_t sh = (uint32_t)((uintptr_t)sp & 3u);
    const uint32_t *ab = (const uint32_t *)(sp - sh);
    const uint32_t bp0 = sh * 8u, end_bits = bp0 + (uint32_t)slen * 8u;
    uint32_t ones;
    asm volatile("ld.volatile.shared.u32 %0, [%1];" : "=r"(ones) : "r"((uint32_t)__cvta_generic_to_shared(&lt->ones)));
    const uint32_t cand_blk = cand_base + 16u * b, pub_blk = pub_base + 8u * b;
    w3_lane_coder c;
    w3_lane_coder_init(c, ab, bp0, end_bits);
    uint32_t rhist = 0u, word = 0u;
    auto step = [&](auto careful, uint32_t q) -> uint32_t {
      const uint32_t cw = w3_split_wait32(w3_split_cand_at(cand_blk, q, rhist), 0xFFFFu, w3_split_tag(q), budget);
      return w3_split_coder_step<decltype(careful)::value>(c, rhist, q, cw, pub_blk, ab, end_bits, ones);
    };
    for (uint32_t i = 0; i < n_steps; i++) {
      uint32_t byte = 0;
      if (__all_sync(0xFFFFFFFFu, c.bp + W3_LANE_FAST_MARGIN <= end_bits)) {
#pragma unroll
        for (uint32_t a = 0; a < 8u; a++) byte = (byte << 1) | step(std::false_type(), 8u * i + a);
      } else {
#pragma unroll 1
        for (uint32_t a = 0; a < 8u; a++) byte = (byte << 1) | step(std::true_type(), 8u * i + a);
      }
      word |= byte << (8u * (i & 3u));
      if ((i & 3u) == 3u) {
        if (owner && i < n) *(uint32_t *)(dst + (i - 3u)) = word;
        word = 0;
      }
      if (owner && i + 1u == n)
        for (uint32_t bb = 0; bb < (n & 3u); bb++) dst[(n & ~3u) + bb] = (uint8_t)(word >> (8u * bb));
    }
  }
  if (budget == 0u) atomicOr(err, W3_ERR_ABORT);
}
#endif
