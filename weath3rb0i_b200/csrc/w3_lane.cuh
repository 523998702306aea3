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
 *               survivor.  Chosen while the chain of a block, not issue, sets
 *               the pace (few blocks per SM).
 *
 * The per-lane code is plain sequential code — `__host__ __device__` — and
 * tests/hostsim runs exactly these statements against the oracle on the CPU
 * (the pair form with both lanes in lockstep and the pick done by hand).
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
 * number of thresholds below the key's cell; the next W3_LANE_FIX thresholds
 * are loaded at once and compared on the top 31 bits of their high words by
 * subtraction (the sign bits add up to the rank; thresholds ascend, so those
 * past the cell compare greater by themselves).  A tie on those 31 bits, or
 * all W3_LANE_FIX thresholds below the key, sends that lane through the
 * general search (w3_rank_lookup, w3_achash_exact) over the tables in global
 * memory: about one lookup in a hundred on text.
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
/* the same, not before `after` is known: pins the load late in a straight-line sequence */
W3_HD void w3_s_ld64v_after(w3_sptr a, uint32_t &x, uint32_t &y, uint32_t after) {
  asm volatile("ld.volatile.shared.v2.u32 {%0, %1}, [%2];" : "=r"(x), "=r"(y) : "r"(a), "r"(after) : "memory");
}
W3_HD void w3_s_st64v(w3_sptr a, uint32_t x, uint32_t y) {
  asm volatile("st.volatile.shared.v2.u32 [%0], {%1, %2};" ::"r"(a), "r"(x), "r"(y) : "memory");
}
W3_HD void w3_s_st32v(w3_sptr a, uint32_t v) { asm volatile("st.volatile.shared.u32 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }
/* the 64-bit store as an exchange: returns the low word it replaced, so that something can be made to wait for it */
W3_HD uint32_t w3_s_xchg64(w3_sptr a, uint32_t x, uint32_t y) {
  uint32_t old;
  asm volatile(
      "{\n"
      ".reg .b64 v, o;\n"
      ".reg .b32 oh;\n"
      "mov.b64 v, {%1, %2};\n"
      "atom.shared.exch.b64 o, [%3], v;\n"
      "mov.b64 {%0, oh}, o;\n"
      "}\n"
      : "=r"(old)
      : "r"(x), "r"(y), "r"(a)
      : "memory");
  return old;
}
W3_HD uint32_t w3_s_ld32v(w3_sptr a) {
  uint32_t v;
  asm volatile("ld.volatile.shared.u32 %0, [%1];" : "=r"(v) : "r"(a) : "memory");
  return v;
}
#else
typedef uintptr_t w3_sptr;
W3_HD uint32_t w3_s_ld32v(w3_sptr a) { return *(const volatile uint32_t *)a; }
W3_HD uint32_t w3_s_xchg64(w3_sptr a, uint32_t x, uint32_t y) {
  const uint32_t old = ((volatile uint32_t *)a)[0];
  ((volatile uint32_t *)a)[0] = x;
  ((volatile uint32_t *)a)[1] = y;
  return old;
}
W3_HD void w3_s_ld64v(w3_sptr a, uint32_t &x, uint32_t &y) { x = ((const volatile uint32_t *)a)[0]; y = ((const volatile uint32_t *)a)[1]; }
W3_HD void w3_s_ld64v_after(w3_sptr a, uint32_t &x, uint32_t &y, uint32_t) { w3_s_ld64v(a, x, y); }
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

/* rank of a key among the thresholds of alignment a, from the top 31 bits of the key */
W3_HD uint32_t w3_lane_rank(uint32_t khi, uint32_t a, const w3_lane_mem &m, bool &bad) {
  const uint32_t base = w3_s_ld8c(m.cellb + a * W3_LANE_NCELL + (khi >> (32u - W3_LANE_PB)));
  const w3_sptr t = m.thi31 + (a * W3_LANE_ROW + base) * 4u;
  uint32_t b[W3_LANE_FIX];
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
  for (int i = 0; i < W3_LANE_FIX; i++) b[i] = w3_s_ld32c(t + 4u * (uint32_t)i);
  /* z = key31 - threshold31 fits a signed word: > 0 the threshold is below the key, < 0 above, 0 undecided */
  const uint32_t k31 = khi >> 1;
  uint32_t above = 0;
  bool tie = false;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
  for (int i = 0; i < W3_LANE_FIX; i++) {
    const uint32_t z = k31 - b[i];
    above += z >> 31;
    tie |= z == 0u;
  }
  bad = tie || above == 0u; /* none above: there may be more thresholds below the key than were looked at */
  return base + (uint32_t)W3_LANE_FIX - above;
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
 * lone warp issues one instruction per ~3 cycles whatever its lanes hold (profiles/r02_*), so
 * 330 cycles per bit.  Here the two halves run in different warps, on different schedulers:
 *   coder warp   lane = block.  Per bit: take the probability of the position from the ring,
 *                decide the bit, publish it, narrow and renormalise.  No table, no counter.
 *   model warps  two, one per parity of the bit position; lane pair = block, lane g guesses g.
 *                A model warp follows every published bit (history, and which of its candidates
 *                came true); at its own positions t it looks up the context of position t + 2
 *                for "bit t + 1 = g", publishes that context's probability, and then adapts the
 *                counter of position t.  Counters of even alignments are only ever touched by
 *                the even warp, odd ones by the odd warp: no counter crosses a warp.
 * What crosses warps are single words in shared memory that carry their own sequence tags
 * (w3_split_ring): no ordering between different words is relied on.
 * The functions below are the model warp's steps; tests/hostsim runs them in the kernel's order. */
struct w3_split_mod {
  uint32_t khi, klo; /* history key through the last published bit taken */
  uint32_t cand_ci;  /* counter offset of this lane's candidate for the warp's next position */
};
/* a published bit enters the history — ac_history.rs:23-26 */
W3_HD void w3_split_take(w3_split_mod &s, uint32_t bit) {
  s.klo = w3_fsr(s.klo, s.khi, 1u);
  s.khi = (s.khi >> 1) | (bit << 31);
}
/* candidate for the position of alignment a2 if the next bit is g: returns Counter::p << 16 of its context */
#ifndef W3_SPLIT_CACHE_P
#define W3_SPLIT_CACHE_P 0 /* 1: Counter::p << 16 of every counter in a second table per block (see kCacheP) */
#endif
W3_HD uint32_t w3_split_guess(w3_split_mod &s, uint32_t g, uint32_t a2, const w3_lane_mem &m) {
  const w3_pair_cand c = w3_key_guess<W3_SPLIT_CACHE_P != 0>(s.khi, s.klo, g, a2, m);
  s.cand_ci = c.ci;
  return c.P;
}
/* Counter::update of the context that coded `bit` — ordern_entropy.rs:32-34 */
W3_HD void w3_split_adapt(uint32_t ci, uint32_t bit, const w3_lane_mem &m) {
  const uint32_t nv = w3_counter_update(w3_s_ld32(m.cnt + ci), bit);
  w3_s_st32(m.cnt + ci, nv);
  if (W3_SPLIT_CACHE_P) w3_s_st32(m.cnt + W3_LANE_P_OFF + ci, w3_lane_P(nv));
}
/* ring words.  A candidate: Counter::p in the top 16 bits, tag of its position q in the low 8;
 * a bit word: bit of block b at bit b, tag of its position in the top 16.  Tags have their top
 * bit set (a cleared ring matches nothing) and differ between consecutive users of a slot. */
#define W3_SPLIT_CAND_SLOTS 4u
#define W3_SPLIT_BIT_SLOTS 8u
W3_HD uint32_t w3_split_cand_tag(uint32_t q) { return 0x80u | ((q >> 2) & 0x7Fu); }
W3_HD uint32_t w3_split_bit_tag(uint32_t t) { return (0x8000u | ((t >> 3) & 0x7FFFu)) << 16; }


/* ---- split form, two bits ahead ---------------------------------------------------
 * Measured (profiles/r02_split_variants.txt): with ONE guessed bit the split form is no faster
 * than the pair form — a step costs half of the round trip "bit known -> lookup -> probability
 * -> next-but-one bit", and that round trip is a chain of dependent shared-memory loads and
 * conversions (~700 cycles), not an issue-rate matter.  So the models look further ahead: four
 * lanes per block guess the next TWO bits (lane g: bit q-2 = g & 1, bit q-1 = g >> 1), the
 * candidate for position q is started as soon as bit q-3 is out, and the round trip is spread
 * over three steps.  kNB model warps (2 or 4); warp W owns the positions q = W (mod kNB) —
 * their lookups, their probabilities and their counters (alignments = W mod kNB: no counter
 * crosses a warp).  Its activation for q, once bit q-3 is published:
 *   take the kNB bits published since its last activation into the history key;
 *   look up the context of position q for its guess, publish Counter::p of it (from a table of
 *   probabilities kept next to the counters: one load instead of a division on the chain);
 *   then adapt the counter (and the probability) of its position q-4, whose context was one of
 *   the warp's earlier candidates — the lane that guessed right is known by now.
 * Bit words carry two bits per block (positions t-1 and t, interleaved: block b at bits 2b and
 * 2b+1), so one word per two steps is all a model reads. */
template <uint32_t kNB>
struct w3_ahead_mod {
  uint32_t khi, klo;       /* history key through the last bit taken */
  uint32_t ci[4u / kNB];   /* this lane's candidate counter offsets of the warp's last positions, newest first */
  uint32_t vprev;          /* kNB == 2: the two bits taken last time */
};
template <uint32_t kNB>
W3_HD void w3_ahead_init(w3_ahead_mod<kNB> &s) {
  s.khi = s.klo = s.vprev = 0u;
  for (uint32_t i = 0; i < 4u / kNB; i++) s.ci[i] = W3_CTX0_RANKED * 4u;
}
/* Activation for position q >= 1 of alignment a: v = the kNB bits q-2-kNB .. q-3, oldest lowest (zeros
 * before the start of the block).  Returns Counter::p << 16 of this lane's candidate and leaves its counter
 * offset in ci_new; the caller then picks and adapts position q-4 and calls w3_ahead_rotate. */
template <uint32_t kNB>
W3_HD uint32_t w3_ahead_guess(w3_ahead_mod<kNB> &s, uint32_t v, uint32_t g, uint32_t a, const w3_lane_mem &m, uint32_t &ci_new) {
  s.klo = w3_fsr(s.klo, s.khi, kNB); /* ac_history.rs:23-26, kNB times */
  s.khi = (s.khi >> kNB) | (v << (32u - kNB));
  const uint32_t gk = (s.khi >> 2) | (g << 30), glo = w3_fsr(s.klo, s.khi, 2u);
  bool bad;
  uint32_t r = w3_lane_rank(gk, a, m, bad); /* hash — ordern_entropy.rs:36-45 */
  if (bad) r = w3_lane_rank_decl(gk, glo, a, m);
  ci_new = ((r << W3_ALIGN_BITS) | a) * 4u;
  return w3_s_ld32(m.cnt + W3_LANE_P_OFF + ci_new);
}
/* which earlier candidate, and which bit, the adapt of position q-4 needs: pick = bits (q-6, q-5) as a lane
 * index within the block's four lanes, bit = bit q-4 */
template <uint32_t kNB>
W3_HD void w3_ahead_adapt_args(const w3_ahead_mod<kNB> &s, uint32_t v, uint32_t &pick, uint32_t &bit) {
  if (kNB == 4u) { pick = v & 3u; bit = (v >> 2) & 1u; }
  else { pick = s.vprev; bit = v & 1u; }
}
/* Counter::update of position q-4's context (ordern_entropy.rs:32-34) and its probability */
W3_HD void w3_ahead_adapt(uint32_t ci, uint32_t bit, const w3_lane_mem &m) {
  const uint32_t nv = w3_counter_update(w3_s_ld32(m.cnt + ci), bit);
  w3_s_st32(m.cnt + ci, nv);
  w3_s_st32(m.cnt + W3_LANE_P_OFF + ci, w3_lane_P(nv));
}
template <uint32_t kNB>
W3_HD void w3_ahead_rotate(w3_ahead_mod<kNB> &s, uint32_t v, uint32_t ci_new) {
  if (kNB == 2u) s.ci[1] = s.ci[0];
  s.ci[0] = ci_new;
  s.vprev = v;
}
#define W3_AHEAD_SLOTS 8u
W3_HD uint32_t w3_ahead_cand_tag(uint32_t q) { return 0x80u | ((q >> 3) & 0x7Fu); }


/* ---- elastic split form ("lane teams") ----------------------------------------------
 * Measured on the forms above (profiles/r02_ahead_prof.txt): the coder warp of the two-bits-ahead
 * form needs ~215 cycles for the 56 instructions of a step even when it never waits — the vote,
 * the wait's branch and the convergence barriers around them are what a lone warp pays dearly
 * (the team decoder's finding, w3_team.cuh: ~40 cycles per conditional branch).  So here the coder
 * warp runs the team decoder's branch-free ATTEMPTS, one lane per block, every lane at its own
 * position: an attempt decodes the lane's next bit iff the candidate it fetched carried the right
 * tag, else it changes nothing and fetches again.  Nothing keeps the blocks of a group
 * in step any more; what couples them is only that a model warp serves its blocks' positions in
 * one instruction stream.
 *   pub[t & 7][b]        = {the block's last 32 decoded bits (newest highest), t} after t bits
 *   cand[q & 7][b][g]    = the candidate for position q given bits (q-2, q-1) = (g & 1, g >> 1):
 *                          Counter::p << 16 | tag (q + 1) & 0xFFFF
 * Model warp W of kNB (4 or 8) owns the positions q = W (mod kNB): guess(q) once bit q-3 is
 * published, adapt(q) (counter and probability of the context that coded bit q) once bit q is. */
#define W3_LT_SLOTS 8u
#define W3_LT_MAXB 8u
#define W3_LT_ATTEMPTS 16u /* per trip of the coder warp: straight-line code between two looks at the loop conditions */
struct w3_lt_ring {
  uint32_t cand[W3_LT_SLOTS][W3_LT_MAXB][4];
  uint32_t pub[W3_LT_SLOTS][W3_LT_MAXB][2];
  uint32_t dummy_cand[W3_LT_SLOTS][W3_LT_MAXB][4]; /* what coder lanes without a block look at: stays zero */
  uint32_t dummy_pub[W3_LT_SLOTS][W3_LT_MAXB][2];
};
static_assert(sizeof(w3_lt_ring) % 16 == 0, "ring keeps the counters aligned");
W3_HD uint32_t w3_lt_tag(uint32_t q) { return (q + 1u) & 0xFFFFu; }

struct w3_lt_coder {
  uint32_t x1, x2, x, bp;  /* ArithmeticCoder + stream position — arithmetic_coder.rs:11-18 */
  uint32_t rhist, rlo, t;  /* decoded bits, newest at bit 31 of rhist; their number */
  uint32_t cw;             /* the candidate word fetched for position t */
};
W3_HD void w3_lt_coder_init(w3_lt_coder &c, const uint32_t *ab, uint32_t bp0, uint32_t end_bits) {
  c.x1 = 0u;
  c.x2 = 0xFFFFFFFFu;
  c.x = w3_lane_fresh<true>(ab, bp0, end_bits); /* new_decoder — arithmetic_coder.rs:75-78 */
  c.bp = bp0 + 32u;
  c.rhist = c.rlo = c.t = 0u;
  c.cw = 0u; /* no tag matches 0 at t = 0 */
}
/* One attempt.  cand_blk / pub_blk: the block's columns of the rings (base + 16 b, base + 8 b).
 * Timing (profiles/r02_ahead_tl.txt): what a step costs is the loop "bit q-3 decided -> published -> a model sees
 * it -> looks up -> candidate of q published -> fetched -> bit q decided", over three steps.  The candidate is
 * therefore fetched by the attempt that uses it (one word: by then both guessed bits are known), right after the
 * previous decision — its latency runs beside the renormalisation — and the decision is published at once. */
#ifndef W3_LT_XCHG
#define W3_LT_XCHG 1
#endif
template <bool kCareful>
W3_HD void w3_lt_attempt(w3_lt_coder &c, w3_sptr cand_blk, w3_sptr pub_blk, const uint32_t *ab, uint32_t end_bits, uint32_t ones, uint32_t zero) {
  const bool ok = ((c.cw ^ (c.t + 1u)) & 0xFFFFu) == 0u; /* the published candidate of position t */
  const uint32_t xmid = w3_umulhi(c.x2 - c.x1, c.cw & 0xFFFF0000u); /* lerp — arithmetic_coder.rs:81 */
  const uint32_t bit = (c.x - c.x1) <= xmid ? 1u : 0u;   /* :82 */
  if (ok && bit) c.x2 = c.x1 + xmid;                     /* :85-88 */
  if (ok && !bit) c.x1 = c.x1 + xmid + 1u;
  if (ok) {
    c.rlo = w3_fsr(c.rlo, c.rhist, 1u);
    c.rhist = w3_fsr(c.rhist, bit, 1u);
    c.t += 1u;
  }
#if W3_LT_XCHG
  /* Left to itself the scheduler puts store and fetch where the fetch's latency is just covered — behind the
   * renormalisation it hoists, ~85 cycles after the decision — and the models see the bit that much later.  So
   * the store is an exchange and the fetch's address waits for what it returns (`zero` & it): both now sit right
   * behind the decision, and the fetch's latency runs beside the renormalisation instead of after it. */
  const uint32_t was = w3_s_xchg64(pub_blk + (c.t & 7u) * 64u, c.rhist, c.t);
  c.cw = w3_s_ld32v(cand_blk + (c.t & 7u) * 128u + (c.rhist >> 30) * 4u + (was & zero)); /* position t, given bits (t-2, t-1) */
#else
  w3_s_st64v(pub_blk + (c.t & 7u) * 64u, c.rhist, c.t);
  c.cw = w3_s_ld32v(cand_blk + (c.t & 7u) * 128u + (c.rhist >> 30) * 4u); /* position t, given bits (t-2, t-1) */
#endif
  const uint32_t fresh = w3_lane_fresh<kCareful>(ab, c.bp, end_bits);
  int sh_s, sh_k;
  const int n = w3_ac_renorm_ones(c.x1, c.x2, sh_s, sh_k, ones); /* :89-102; a range that did not narrow is normalised already: n = 0 */
  c.x = w3_fsl(fresh, c.x, (uint32_t)n) ^ ((0u - (uint32_t)sh_k) & 0x80000000u);
  c.bp += (uint32_t)n;
  if (kCareful) {
    const uint32_t cap = end_bits + 64u;
    c.bp = c.bp < cap ? c.bp : cap;
  }
}

struct w3_lt_mod {
  uint32_t khi_prev, klo; /* the history key of the warp's previous position (the general search wants 64 bits) */
  uint32_t ci;            /* counter offset of this lane's last candidate */
  uint32_t ci2[2], pw2[2]; /* prepared: counter offset and Counter::p << 16 of the candidate, for either value of the bit still awaited */
};
/* Position q >= 1 of alignment q & 7, for this lane's guess g of bits (q-2, q-1), in two halves so that the table
 * lookups are over before the last bit they depend on arrives:
 *   prepare  rh3 = the block's decoded bits through q-4 (newest highest): look up context and probability of
 *            position q for bit q-3 = 0 and = 1;
 *   finish   rh = the bits through q-3: pick, publish (unless !valid: a position past the block's end).
 * What a step of the coder waits for is then one load, one select and one store behind the bit's arrival. */
template <uint32_t kNB>
W3_HD void w3_lt_prepare(w3_lt_mod &s, uint32_t rh3, uint32_t q, uint32_t g, const w3_lane_mem &m) {
  const uint32_t a = q & 7u;
  const uint32_t klo = w3_fsr(s.klo, s.khi_prev, kNB); /* the low word once bit q-3 is in; its top bit is a bit of khi_prev either way */
  /* both lookups as one straight-line sequence (their loads interleave); the rare lane whose fixed lookup declined is repaired after */
  const uint32_t gk0 = (rh3 >> 3) | (g << 30), gk1 = gk0 | (1u << 29);
  bool bad0, bad1;
  uint32_t r0 = w3_lane_rank(gk0, a, m, bad0) & 255u, r1 = w3_lane_rank(gk1, a, m, bad1) & 255u; /* hash — ordern_entropy.rs:36-45 */
  if (bad0 | bad1) {
    const uint32_t rh0 = rh3 >> 1;
    if (bad0) r0 = w3_lane_rank_decl(gk0, w3_fsr(klo, rh0, 2u), a, m);
    if (bad1) r1 = w3_lane_rank_decl(gk1, w3_fsr(klo, rh0 | 0x80000000u, 2u), a, m);
  }
  s.ci2[0] = ((r0 << W3_ALIGN_BITS) | a) * 4u;
  s.ci2[1] = ((r1 << W3_ALIGN_BITS) | a) * 4u;
  s.pw2[0] = w3_s_ld32(m.cnt + W3_LANE_P_OFF + s.ci2[0]);
  s.pw2[1] = w3_s_ld32(m.cnt + W3_LANE_P_OFF + s.ci2[1]);
}
template <uint32_t kNB>
W3_HD void w3_lt_finish(w3_lt_mod &s, uint32_t rh, uint32_t q, uint32_t g, uint32_t blk, bool valid, w3_sptr cand_base) {
  uint32_t P = 0x80000000u; /* position 0: OrderNEntropy starts in context 0 with a new counter */
  uint32_t ci = W3_CTX0_RANKED * 4u;
  if (q != 0u) {
    const uint32_t b = rh >> 31;
    ci = b ? s.ci2[1] : s.ci2[0];
    P = b ? s.pw2[1] : s.pw2[0];
    s.klo = w3_fsr(s.klo, s.khi_prev, kNB);
    s.khi_prev = rh;
  }
  s.ci = ci;
  if (valid) w3_s_st32v(cand_base + (((q & 7u) * W3_LT_MAXB + blk) * 4u + g) * 4u, P | w3_lt_tag(q));
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

/* Split form: a CTA of three warps decodes `group` (<= 16) blocks at a time — warp 0 the coder
 * (lane b = block b % group), warps 1 and 2 the models of even and odd positions (lanes 2p, 2p + 1
 * = block p % group; lanes beyond the group shadow real ones and store the same values).  Every
 * warp runs the same number of steps; a wait that does not end (the peer is gone) gives up after
 * W3_SPLIT_SPIN_LIMIT polls, flags W3_ERR_ABORT and stops waiting for good. */
#define W3_SPLIT_THREADS 96u
#define W3_SPLIT_WORDS (W3_SPLIT_CACHE_P ? 2u * W3_NCTX : W3_NCTX) /* per block: counters (+ probabilities) */
#define W3_SPLIT_MAX_GROUP (W3_SPLIT_CACHE_P ? 8u : 16u)
#ifndef W3_SPLIT_SPIN_LIMIT
#define W3_SPLIT_SPIN_LIMIT (1u << 20)
#endif
#ifndef W3_SPLIT_PREFETCH
#define W3_SPLIT_PREFETCH 1 /* the coder warp pulls its streams' next lines into L1 ahead of use */
#endif
struct w3_split_ring {
  uint32_t cand[W3_SPLIT_CAND_SLOTS][32];
  uint32_t bits[W3_SPLIT_BIT_SLOTS];
  uint32_t pad[24];
};
static_assert(sizeof(w3_split_ring) % 16 == 0, "ring keeps the counters aligned");

__device__ __forceinline__ uint32_t w3_ring_ld(uint32_t a) {
  uint32_t v;
  asm volatile("ld.volatile.shared.u32 %0, [%1];" : "=r"(v) : "r"(a) : "memory");
  return v;
}
__device__ __forceinline__ void w3_ring_st(uint32_t a, uint32_t v) {
  asm volatile("st.volatile.shared.u32 [%0], %1;" ::"r"(a), "r"(v) : "memory");
}
/* the word at `a` once (word & mask) == want; budget = polls left before giving up (0 afterwards) */
__device__ __forceinline__ uint32_t w3_ring_wait(uint32_t a, uint32_t mask, uint32_t want, uint32_t &budget) {
  uint32_t w = w3_ring_ld(a);
  if ((w & mask) != want) {
    uint32_t spins = 0;
    do w = w3_ring_ld(a);
    while ((w & mask) != want && ++spins < budget);
    if ((w & mask) != want) budget = 0u;
  }
  return w;
}

#ifndef W3_AHEAD_PROFILE
#define W3_AHEAD_PROFILE 0 /* 1: CTA 0 counts waits and chain lengths into w3_ahead_prof (experiments only) */
#endif
#if W3_AHEAD_PROFILE
__device__ unsigned long long w3_ahead_prof[16];
#define W3_TL_Q0 40000u
#define W3_TL_N 256u
__device__ unsigned int w3_dbg_pubtime[8]; /* CTA 0, block 0: clock when count t was published, slot t & 7 (global memory: only CTA 0 writes) */
__device__ unsigned int w3_ahead_tl[4][W3_TL_N]; /* CTA 0, block 0: [0] coder saw bit q decoded, [1] model saw count q - 2, [2] model published q */
__device__ __forceinline__ uint32_t w3_ring_wait_prof(uint32_t a, uint32_t mask, uint32_t want, uint32_t &budget, uint32_t &n_spun, uint32_t &cyc_spun) {
  uint32_t w = w3_ring_ld(a);
  if ((w & mask) != want) {
    const uint32_t t0 = (uint32_t)clock();
    uint32_t spins = 0;
    do w = w3_ring_ld(a);
    while ((w & mask) != want && ++spins < budget);
    if ((w & mask) != want) budget = 0u;
    n_spun++;
    cyc_spun += (uint32_t)clock() - t0;
  }
  return w;
}
#endif

/* model warp of parity kPar: positions t = kPar, kPar + 2, ... of every block of the group */
template <uint32_t kPar>
__device__ __forceinline__ void w3_split_model_warp(uint32_t n_steps, uint32_t lane, uint32_t blk_in_group, const w3_lane_mem &m,
                                                    uint32_t ring_cand, uint32_t ring_bits, uint32_t &budget) {
  const uint32_t g = lane & 1u;
  w3_split_mod s;
  s.khi = s.klo = 0u;
  s.cand_ci = W3_CTX0_RANKED * 4u; /* position 0: the empty history's context, whatever came "before" */
  if (kPar == 0u) {                /* position 0: Counter::new, p = 32768 */
    w3_ring_st(ring_cand + lane * 4u, 0x80000000u | w3_split_cand_tag(0u));
  } else {                         /* position 1, for either value of bit 0 */
    const uint32_t P = w3_split_guess(s, g, 1u, m);
    w3_ring_st(ring_cand + (1u * 32u + lane) * 4u, P | w3_split_cand_tag(1u));
  }
  for (uint32_t i = 0; i < n_steps; i++) {
    const uint32_t btag = w3_split_bit_tag(8u * i);
#pragma unroll
    for (uint32_t a = kPar; a < 8u; a += 2u) {
      const uint32_t t = 8u * i + a;
      uint32_t b1 = 0u;
      if (t > 0u) { /* bit t - 1: which of the pair's candidates for position t came true */
        const uint32_t slot = (a + 7u) & 7u;
        const uint32_t w = w3_ring_wait(ring_bits + slot * 4u, 0xFFFF0000u, a == 0u ? w3_split_bit_tag(t - 1u) : btag, budget);
        b1 = (w >> blk_in_group) & 1u;
      }
      const uint32_t ci_t = __shfl_sync(0xFFFFFFFFu, s.cand_ci, (lane & 30u) | b1);
      if (t > 0u) w3_split_take(s, b1);
      const uint32_t w0 = w3_ring_wait(ring_bits + a * 4u, 0xFFFF0000u, btag, budget);
      const uint32_t b0 = (w0 >> blk_in_group) & 1u;
      w3_split_take(s, b0);
      const uint32_t P = w3_split_guess(s, g, (a + 2u) & 7u, m);
      w3_ring_st(ring_cand + ((((a + 2u) & 3u) * 32u) + lane) * 4u, P | w3_split_cand_tag(t + 2u));
      w3_split_adapt(ci_t, b0, m);
    }
  }
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
  const uint32_t ring_cand = (uint32_t)__cvta_generic_to_shared(&ring->cand[0][0]);
  const uint32_t ring_bits = (uint32_t)__cvta_generic_to_shared(&ring->bits[0]);
  uint32_t budget = W3_SPLIT_SPIN_LIMIT;

  w3_lane_stage(lt, tab);
  for (uint32_t first = blockIdx.x * group; first < n_blocks; first += gridDim.x * group) {
    __syncthreads();
    for (uint32_t i = threadIdx.x; i < group * W3_SPLIT_WORDS; i += blockDim.x)
      counts[i] = (W3_SPLIT_CACHE_P && (i % W3_SPLIT_WORDS) >= W3_NCTX) ? 0x80000000u : 0u; /* Counter::new — ordern_entropy.rs:17 */
    for (uint32_t i = threadIdx.x; i < sizeof(w3_split_ring) / 4u; i += blockDim.x) reinterpret_cast<uint32_t *>(ring)[i] = 0u;
    __syncthreads();
    const uint32_t in_group = n_blocks - first < group ? n_blocks - first : group;
    const uint32_t n_steps = block_size < n_total ? block_size : (uint32_t)n_total; /* the same for every warp */
    if (warp != 0u) {
      const uint32_t b = (lane >> 1) % in_group;
      const w3_lane_mem m = w3_lane_mem_of(lt, counts + b * W3_SPLIT_WORDS, tab);
      if (warp == 1u) w3_split_model_warp<0u>(n_steps, lane, b, m, ring_cand, ring_bits, budget);
      else w3_split_model_warp<1u>(n_steps, lane, b, m, ring_cand, ring_bits, budget);
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
    w3_lane_coder c;
    w3_lane_coder_init(c, ab, bp0, end_bits);
    const uint32_t my_cand = ring_cand + 8u * b; /* + slot * 128 + bit * 4: the pair's lane that guessed right */
    uint32_t prev = 0u, word = 0u;
    auto step = [&](auto careful, uint32_t i, uint32_t a) -> uint32_t {
      const uint32_t P = w3_ring_wait(my_cand + (a & 3u) * 128u + prev * 4u, 0xFFu, w3_split_cand_tag(8u * i + a), budget) & 0xFFFF0000u;
      uint32_t mm;
      const uint32_t bit = w3_lane_coder_bit(c, P, mm);
      const uint32_t all = __ballot_sync(0xFFFFFFFFu, bit);
      if (lane == 0u) w3_ring_st(ring_bits + a * 4u, (all & 0xFFFFu) | w3_split_bit_tag(8u * i));
      w3_lane_coder_shift<decltype(careful)::value>(c, bit, mm, ab, end_bits, ones);
      prev = bit;
      return bit;
    };
    for (uint32_t i = 0; i < n_steps; i++) {
      uint32_t byte = 0;
#if W3_SPLIT_PREFETCH
      if ((i & 15u) == 0u) { /* a byte takes ~3.5 stream bits: 128 bytes ahead covers the next 16 several times over */
        const uint32_t ahead = (c.bp >> 5) + 32u;
        if (ahead * 32u < end_bits) asm volatile("prefetch.global.L1 [%0];" ::"l"(ab + ahead));
      }
#endif
      if (__all_sync(0xFFFFFFFFu, c.bp + W3_LANE_FAST_MARGIN <= end_bits)) {
#pragma unroll
        for (uint32_t a = 0; a < 8u; a++) byte = (byte << 1) | step(std::false_type(), i, a);
      } else {
#pragma unroll 1
        for (uint32_t a = 0; a < 8u; a++) byte = (byte << 1) | step(std::true_type(), i, a);
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

/* Split form, two bits ahead: warp 0 the coder (lanes 2b, 2b + 1 = block b % group: both run the
 * block's coder, the odd lane votes the bit just decoded and the even lane the one before, so that one
 * ballot is the bit word), kNB model warps (lanes 4b .. 4b + 3 = block b % group, lane & 3 = guess).
 * Model warp W = 0 .. kNB - 1 is CTA warp W + 1, except that with four model warps the fourth is
 * warp 5 and warp 4 idles: the coder warp keeps its scheduler to itself. */
#define W3_AHEAD_MAX_GROUP 8u
#define W3_AHEAD_WORDS (2u * W3_NCTX) /* per block: counters + probabilities */
#define W3_AHEAD_THREADS(nb) ((nb) == 4u ? 192u : 96u)
struct w3_ahead_ring {
  uint32_t cand[W3_AHEAD_SLOTS][32];
  uint32_t bits[W3_AHEAD_SLOTS];
  uint32_t pad[24];
};
static_assert(sizeof(w3_ahead_ring) % 16 == 0, "ring keeps the counters aligned");

template <uint32_t kNB>
__device__ __forceinline__ void w3_ahead_model_warp(uint32_t W, uint32_t n_steps, uint32_t lane, uint32_t blk_in_group, const w3_lane_mem &m,
                                                    uint32_t ring_cand, uint32_t ring_bits, uint32_t &budget) {
  const uint32_t g = lane & 3u, sh = 2u * blk_in_group;
  w3_ahead_mod<kNB> s;
  w3_ahead_init(s);
#if W3_AHEAD_PROFILE
  uint32_t n_spun = 0, cyc_spun = 0, cyc_chain = 0, n_act = 0;
  const uint32_t t_begin = (uint32_t)clock();
#endif
  for (uint32_t i = 0; i < n_steps; i++) {
#pragma unroll
    for (uint32_t k = 0; k < 8u / kNB; k++) {
      const uint32_t a = W + k * kNB, q = 8u * i + a;
      uint32_t v = 0u;
      if (kNB == 4u && q >= 5u) v = (w3_ring_wait(ring_bits + ((q - 5u) & 7u) * 4u, 0xFFFF0000u, w3_split_bit_tag(q - 5u), budget) >> sh) & 3u;
      if (q >= 3u) {
#if W3_AHEAD_PROFILE
        const uint32_t va = (w3_ring_wait_prof(ring_bits + ((q - 3u) & 7u) * 4u, 0xFFFF0000u, w3_split_bit_tag(q - 3u), budget, n_spun, cyc_spun) >> sh) & 3u;
#else
        const uint32_t va = (w3_ring_wait(ring_bits + ((q - 3u) & 7u) * 4u, 0xFFFF0000u, w3_split_bit_tag(q - 3u), budget) >> sh) & 3u;
#endif
        v = kNB == 4u ? v | (va << 2) : va;
      }
#if W3_AHEAD_PROFILE
      const uint32_t t_arrived = (uint32_t)clock();
#endif
      uint32_t ci_new = W3_CTX0_RANKED * 4u, P = 0x80000000u; /* position 0: OrderNEntropy starts in context 0 with a new counter */
      if (q != 0u) P = w3_ahead_guess<kNB>(s, v, g, a, m, ci_new);
      w3_ring_st(ring_cand + ((q & 7u) * 32u + lane) * 4u, P | w3_ahead_cand_tag(q));
#if W3_AHEAD_PROFILE
      cyc_chain += (uint32_t)clock() - t_arrived;
      n_act++;
#endif
      uint32_t pick, bit;
      w3_ahead_adapt_args<kNB>(s, v, pick, bit);
      const uint32_t ci_old = __shfl_sync(0xFFFFFFFFu, s.ci[4u / kNB - 1u], (lane & 28u) | pick);
      if (q >= 4u) w3_ahead_adapt(ci_old, bit, m);
      w3_ahead_rotate<kNB>(s, v, ci_new);
    }
  }
#if W3_AHEAD_PROFILE
  if (blockIdx.x == 0 && lane == 0 && W == 1u) {
    w3_ahead_prof[4] = (uint32_t)clock() - t_begin; /* wraps: only good for short runs */
    w3_ahead_prof[5] = n_act;
    w3_ahead_prof[6] = n_spun;
    w3_ahead_prof[7] = cyc_spun;
    w3_ahead_prof[8] = cyc_chain;
  }
#endif
}

template <uint32_t kNB>
__global__ void __launch_bounds__(W3_AHEAD_THREADS(kNB), 1)
w3_decode_blocks_ahead(const uint8_t *__restrict__ comp, const uint64_t *__restrict__ offsets, uint32_t n_blocks,
                       uint32_t block_size, uint64_t n_total, uint8_t *__restrict__ out, uint32_t *__restrict__ err,
                       const w3_hashtab *__restrict__ tab, uint32_t group) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  w3_lane_tab *lt = reinterpret_cast<w3_lane_tab *>(smem_raw);
  w3_ahead_ring *ring = reinterpret_cast<w3_ahead_ring *>(smem_raw + sizeof(w3_lane_tab));
  uint32_t *counts = reinterpret_cast<uint32_t *>(smem_raw + sizeof(w3_lane_tab) + sizeof(w3_ahead_ring));
  const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31u;
  const uint32_t ring_cand = (uint32_t)__cvta_generic_to_shared(&ring->cand[0][0]);
  const uint32_t ring_bits = (uint32_t)__cvta_generic_to_shared(&ring->bits[0]);
  uint32_t budget = W3_SPLIT_SPIN_LIMIT;

  w3_lane_stage(lt, tab);
  for (uint32_t first = blockIdx.x * group; first < n_blocks; first += gridDim.x * group) {
    __syncthreads();
    for (uint32_t i = threadIdx.x; i < group * W3_AHEAD_WORDS; i += blockDim.x)
      counts[i] = (i % W3_AHEAD_WORDS) >= W3_NCTX ? 0x80000000u : 0u; /* Counter::new — ordern_entropy.rs:17: counts 0, p = 32768 */
    for (uint32_t i = threadIdx.x; i < sizeof(w3_ahead_ring) / 4u; i += blockDim.x) reinterpret_cast<uint32_t *>(ring)[i] = 0u;
    __syncthreads();
    const uint32_t in_group = n_blocks - first < group ? n_blocks - first : group;
    const uint32_t n_steps = block_size < n_total ? block_size : (uint32_t)n_total; /* the same for every warp */
    if (warp == 4u) continue; /* only there so that the fourth model warp does not land on the coder's scheduler */
    if (warp != 0u) {
      const uint32_t W = warp == 5u ? 0u : warp == kNB ? 0u : warp; /* kNB = 2: warps 1, 2 -> W 1, 0; kNB = 4: warps 1, 2, 3, 5 -> W 1, 2, 3, 0 */
      const uint32_t b = (lane >> 2) % in_group;
      w3_ahead_model_warp<kNB>(W, n_steps, lane, b, w3_lane_mem_of(lt, counts + b * W3_AHEAD_WORDS, tab), ring_cand, ring_bits, budget);
      continue;
    }
    /* ---- the coder warp ---- */
    const uint32_t b = (lane >> 1) % in_group;
    const bool owner = (lane & 1u) && (lane >> 1) < in_group;
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
    w3_lane_coder c;
    w3_lane_coder_init(c, ab, bp0, end_bits);
    const uint32_t my_cand = ring_cand + 16u * b; /* + slot * 128 + (the last two bits) * 4: the lane that guessed them */
    uint32_t prev2 = 0u, word = 0u; /* prev2: bit q-2 | bit q-1 << 1 */
#if W3_AHEAD_PROFILE
    uint32_t n_spun = 0, cyc_spun = 0;
    const uint32_t t_begin = (uint32_t)clock();
#endif
    auto step = [&](auto careful, uint32_t i, uint32_t a) -> uint32_t {
#if W3_AHEAD_PROFILE
      const uint32_t P = w3_ring_wait_prof(my_cand + a * 128u + prev2 * 4u, 0xFFu, w3_ahead_cand_tag(8u * i + a), budget, n_spun, cyc_spun) & 0xFFFF0000u;
#else
      const uint32_t P = w3_ring_wait(my_cand + a * 128u + prev2 * 4u, 0xFFu, w3_ahead_cand_tag(8u * i + a), budget) & 0xFFFF0000u;
#endif
      uint32_t mm;
      const uint32_t bit = w3_lane_coder_bit(c, P, mm);
      const uint32_t all = __ballot_sync(0xFFFFFFFFu, (lane & 1u) ? bit : (prev2 >> 1));
      if (lane == 0u) w3_ring_st(ring_bits + a * 4u, (all & 0xFFFFu) | w3_split_bit_tag(8u * i + a));
      w3_lane_coder_shift<decltype(careful)::value>(c, bit, mm, ab, end_bits, ones);
      prev2 = (prev2 >> 1) | (bit << 1);
      return bit;
    };
    for (uint32_t i = 0; i < n_steps; i++) {
      uint32_t byte = 0;
#if W3_SPLIT_PREFETCH
      if ((i & 15u) == 0u) {
        const uint32_t ahead = (c.bp >> 5) + 32u;
        if (ahead * 32u < end_bits) asm volatile("prefetch.global.L1 [%0];" ::"l"(ab + ahead));
      }
#endif
      if (__all_sync(0xFFFFFFFFu, c.bp + W3_LANE_FAST_MARGIN <= end_bits)) {
#pragma unroll
        for (uint32_t a = 0; a < 8u; a++) byte = (byte << 1) | step(std::false_type(), i, a);
      } else {
#pragma unroll 1
        for (uint32_t a = 0; a < 8u; a++) byte = (byte << 1) | step(std::true_type(), i, a);
      }
      word |= byte << (8u * (i & 3u));
      if ((i & 3u) == 3u) {
        if (owner && i < n) *(uint32_t *)(dst + (i - 3u)) = word;
        word = 0;
      }
      if (owner && i + 1u == n)
        for (uint32_t bb = 0; bb < (n & 3u); bb++) dst[(n & ~3u) + bb] = (uint8_t)(word >> (8u * bb));
    }
#if W3_AHEAD_PROFILE
    if (blockIdx.x == 0 && lane == 1u) {
      w3_ahead_prof[0] = (uint32_t)clock() - t_begin;
      w3_ahead_prof[1] = 8ull * n_steps;
      w3_ahead_prof[2] = n_spun;
      w3_ahead_prof[3] = cyc_spun;
    }
#endif
  }
  if (budget == 0u) atomicOr(err, W3_ERR_ABORT);
}

/* Elastic split form.  Warp 0: the coder (lane b = block b of the group); warps 1.. : the model warps
 * (lanes 4b .. 4b + 3 = block b, lane & 3 = guess), except that warps 4 and 8 idle so that the coder warp
 * keeps its scheduler to itself. */
#define W3_LT_THREADS(nb) ((nb) == 4u ? 192u : 352u)
#define W3_LT_WORDS (2u * W3_NCTX) /* per block: counters + probabilities */
#ifndef W3_LT_STALL_LIMIT
#define W3_LT_STALL_LIMIT (1u << 16) /* trips without any lane decoding a bit */
#endif

/* the block's pub word for bit count `count`, once it is there (budget: see w3_ring_wait) */
__device__ __forceinline__ uint32_t w3_lt_wait(uint32_t pub_blk, uint32_t count, uint32_t &budget) {
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

template <uint32_t kNB>
__device__ __forceinline__ void w3_lt_model_warp(uint32_t W, uint32_t n_steps, uint32_t nbits, uint32_t lane, uint32_t blk, const w3_lane_mem &m,
                                                 uint32_t cand_base, uint32_t pub_blk, uint32_t &budget) {
  const uint32_t g = lane & 3u;
  w3_lt_mod s;
  s.khi_prev = s.klo = 0u;
  s.ci = W3_CTX0_RANKED * 4u;
#if W3_AHEAD_PROFILE
  uint32_t prof_wait = 0, prof_chain = 0, prof_n = 0;
  const uint32_t prof_begin = (uint32_t)clock();
#endif
  for (uint32_t i = 0; i < n_steps; i++) {
#pragma unroll
    for (uint32_t k = 0; k < 8u / kNB; k++) {
      const uint32_t q = 8u * i + W + k * kNB;
      uint32_t rh = 0u;
      if (kNB == 8u) { /* adapt(q - 8) as soon as its bit is out, long before the guess for q needs the probability */
        const bool on = q >= 8u && q - 8u < nbits;
        if (on) rh = w3_lt_wait(pub_blk, q - 7u, budget);
        const uint32_t ci_old = __shfl_sync(0xFFFFFFFFu, s.ci, (lane & 28u) | ((rh >> 29) & 3u));
        if (on) w3_ahead_adapt(ci_old, rh >> 31, m);
      }
      const bool valid = q < nbits;
      uint32_t rh3 = 0u;
#if W3_AHEAD_PROFILE
      const uint32_t t_arr = (uint32_t)clock();
#endif
      if (valid && q >= 3u) rh3 = w3_lt_wait(pub_blk, q - 3u, budget);
#if W3_AHEAD_PROFILE
      const uint32_t t_p0 = (uint32_t)clock();
#endif
      if (q != 0u) w3_lt_prepare<kNB>(s, rh3, q, g, m);
#if W3_AHEAD_PROFILE
      if (blockIdx.x == 0 && lane == 0 && q >= W3_TL_Q0 && q < W3_TL_Q0 + W3_TL_N) w3_ahead_tl[3][q - W3_TL_Q0] = ((t_p0 - t_arr) & 0xFFFFu) | (((t_p0 - *(volatile unsigned int *)&w3_dbg_pubtime[(q - 3u) & 7u]) & 0xFFFFu) << 16);
#endif
      if (kNB == 4u) { /* adapt(q - 4): bit q - 4 is the newest of rh3, its context came from guesses of bits (q - 6, q - 5) */
        const uint32_t ci_old = __shfl_sync(0xFFFFFFFFu, s.ci, (lane & 28u) | ((rh3 >> 29) & 3u));
        if (valid && q >= 4u) w3_ahead_adapt(ci_old, rh3 >> 31, m);
      }
#if W3_AHEAD_PROFILE
      const uint32_t t_w0 = (uint32_t)clock();
#endif
      rh = 0u;
      if (valid && q >= 2u) rh = w3_lt_wait(pub_blk, q - 2u, budget);
#if W3_AHEAD_PROFILE
      const uint32_t t_w1 = (uint32_t)clock();
#endif
      w3_lt_finish<kNB>(s, rh, q, g, blk, valid, cand_base);
#if W3_AHEAD_PROFILE
      if (blockIdx.x == 0 && lane == 0 && q >= W3_TL_Q0 && q < W3_TL_Q0 + W3_TL_N) {
        w3_ahead_tl[1][q - W3_TL_Q0] = t_w1;
        w3_ahead_tl[2][q - W3_TL_Q0] = (uint32_t)clock();
      }
      prof_wait += t_w1 - t_w0;
      prof_chain += (uint32_t)clock() - t_w1;
      prof_n++;
#endif
    }
  }
#if W3_AHEAD_PROFILE
  if (blockIdx.x == 0 && lane == 0 && W == 1u) {
    w3_ahead_prof[4] = (uint32_t)clock() - prof_begin;
    w3_ahead_prof[5] = prof_n;
    w3_ahead_prof[6] = 0;
    w3_ahead_prof[7] = prof_wait;
    w3_ahead_prof[8] = prof_chain;
  }
#endif
}

template <uint32_t kNB>
__global__ void __launch_bounds__(W3_LT_THREADS(kNB), 1)
w3_decode_blocks_lt(const uint8_t *__restrict__ comp, const uint64_t *__restrict__ offsets, uint32_t n_blocks,
                    uint32_t block_size, uint64_t n_total, uint8_t *__restrict__ out, uint32_t *__restrict__ err,
                    const w3_hashtab *__restrict__ tab, uint32_t group) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  w3_lane_tab *lt = reinterpret_cast<w3_lane_tab *>(smem_raw);
  w3_lt_ring *ring = reinterpret_cast<w3_lt_ring *>(smem_raw + sizeof(w3_lane_tab));
  uint32_t *counts = reinterpret_cast<uint32_t *>(smem_raw + sizeof(w3_lane_tab) + sizeof(w3_lt_ring));
  const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31u;
  const uint32_t cand_base = (uint32_t)__cvta_generic_to_shared(&ring->cand[0][0][0]);
  const uint32_t pub_base = (uint32_t)__cvta_generic_to_shared(&ring->pub[0][0][0]);
  uint32_t budget = W3_SPLIT_SPIN_LIMIT;
  bool dead = false;

  w3_lane_stage(lt, tab);
  for (uint32_t first = blockIdx.x * group; first < n_blocks; first += gridDim.x * group) {
    __syncthreads();
    for (uint32_t i = threadIdx.x; i < group * W3_LT_WORDS; i += blockDim.x)
      counts[i] = (i % W3_LT_WORDS) >= W3_NCTX ? 0x80000000u : 0u; /* Counter::new — ordern_entropy.rs:17: counts 0, p = 32768 */
    for (uint32_t i = threadIdx.x; i < sizeof(w3_lt_ring) / 4u; i += blockDim.x) reinterpret_cast<uint32_t *>(ring)[i] = 0u;
    __syncthreads();
    const uint32_t in_group = n_blocks - first < group ? n_blocks - first : group;
    const uint32_t n_steps = block_size < n_total ? block_size : (uint32_t)n_total; /* the longest block of any group */
    if (warp != 0u && (warp & 3u) == 0u) continue;
    if (warp != 0u) {
      const uint32_t W = warp - 1u - (warp >> 2);
      const uint32_t b = lane >> 2;
      uint32_t nbits = 0u;
      if (b < in_group) {
        const uint64_t off = (uint64_t)(first + b) * block_size;
        nbits = 8u * (uint32_t)(n_total - off < block_size ? n_total - off : block_size);
      }
      const uint32_t bb = b < in_group ? b : 0u; /* lanes without a block: nbits = 0, they neither wait nor publish anything valid nor adapt */
      w3_lt_model_warp<kNB>(W, n_steps, nbits, lane, b, w3_lane_mem_of(lt, counts + bb * W3_LT_WORDS, tab),
                            b < in_group ? cand_base : (uint32_t)__cvta_generic_to_shared(&ring->dummy_cand[0][0][0]), pub_base + 8u * bb, budget);
      continue;
    }
    /* ---- the coder warp ---- */
    const bool mine = lane < in_group;
    const uint32_t blk = first + (mine ? lane : 0u);
    const uint64_t off = (uint64_t)blk * block_size;
    const uint32_t n = mine ? (uint32_t)(n_total - off < block_size ? n_total - off : block_size) : 0u;
    const uint32_t nbits = 8u * n;
    const uint64_t s0 = offsets[blk];
    uint64_t slen = offsets[blk + 1] - s0;
    if (slen > 0x1FFFFF00ull) { if (mine) atomicOr(err, W3_ERR_ABORT); slen = 0; } /* decodes zeros, flagged */
    const uint8_t *sp = comp + s0;
    uint8_t *dst = out + off;
    const uint32_t sh = (uint32_t)((uintptr_t)sp & 3u);
    const uint32_t *idle_words = reinterpret_cast<const uint32_t *>(tab); /* readable words for lanes that have nothing to read */
    const uint32_t *ab = mine ? (const uint32_t *)(sp - sh) : idle_words;
    const uint32_t bp0 = sh * 8u;
    uint32_t end_bits = mine ? bp0 + (uint32_t)slen * 8u : 0u;
    uint32_t ones;
    asm volatile("ld.volatile.shared.u32 %0, [%1];" : "=r"(ones) : "r"((uint32_t)__cvta_generic_to_shared(&lt->ones)));
    const uint32_t col = lane & 7u;
    const uint32_t cand_blk = (mine ? cand_base : (uint32_t)__cvta_generic_to_shared(&ring->dummy_cand[0][0][0])) + 16u * col;
    const uint32_t pub_blk = (mine ? pub_base : (uint32_t)__cvta_generic_to_shared(&ring->dummy_pub[0][0][0])) + 8u * col;
    const uint32_t zero = ones + 1u; /* a zero the compiler cannot see through */
    w3_lt_coder c;
    w3_lt_coder_init(c, ab, mine ? bp0 : 0u, end_bits);
    bool busy = nbits != 0u; /* this lane still reads its stream */
    uint32_t stalled = 0;
#if W3_AHEAD_PROFILE
    uint32_t prof_trips = 0;
    const uint32_t prof_begin = (uint32_t)clock();
#endif
    while (__any_sync(0xFFFFFFFFu, c.t < nbits) && !dead) {
      if (busy && c.t >= nbits) { /* done: stop looking at the stream (its last words may be the buffer's last) */
        busy = false;
        ab = idle_words;
        c.bp = 0u;
        end_bits = 0u;
      }
      const uint32_t t0 = c.t;
      if (__all_sync(0xFFFFFFFFu, !busy || c.bp + (W3_LT_ATTEMPTS * 31u + 64u) <= end_bits)) {
#pragma unroll
        for (uint32_t k = 0; k < W3_LT_ATTEMPTS; k++) {
#if W3_AHEAD_PROFILE
          const uint32_t tb = c.t;
#endif
          w3_lt_attempt<false>(c, cand_blk, pub_blk, ab, end_bits, ones, zero);
#if W3_AHEAD_PROFILE
          if (blockIdx.x == 0 && lane == 0 && c.t != tb) {
            const uint32_t now = (uint32_t)clock();
            *(volatile unsigned int *)&w3_dbg_pubtime[c.t & 7u] = now;
            if (tb >= W3_TL_Q0 && tb < W3_TL_Q0 + W3_TL_N) w3_ahead_tl[0][tb - W3_TL_Q0] = now;
          }
#endif
        }
      } else {
#pragma unroll 1
        for (uint32_t k = 0; k < W3_LT_ATTEMPTS; k++) w3_lt_attempt<true>(c, cand_blk, pub_blk, ab, end_bits, ones, zero); /* end_bits = 0: reads nothing */
      }
      if ((c.t ^ t0) & 32u) { /* a word completed c.t - (c.t & ~31) < 16 bits ago */
        const uint32_t w = w3_fsl(c.rlo, c.rhist, c.t & 31u);
        *reinterpret_cast<uint32_t *>(dst + (c.t >> 5) * 4u - 4u) = w3_bswap32(__brev(w));
      }
      stalled = __any_sync(0xFFFFFFFFu, c.t != t0) ? 0u : stalled + 1u;
      if (stalled > W3_LT_STALL_LIMIT) dead = true;
#if W3_AHEAD_PROFILE
      prof_trips++;
#endif
    }
#if W3_AHEAD_PROFILE
    if (blockIdx.x == 0 && lane == 0u) {
      w3_ahead_prof[0] = (uint32_t)clock() - prof_begin;
      w3_ahead_prof[1] = nbits;
      w3_ahead_prof[2] = prof_trips * W3_LT_ATTEMPTS;
      w3_ahead_prof[3] = 0;
    }
#endif
    if (mine) {
      const uint32_t rem = n & 3u; /* bytes after the last whole word */
      const uint32_t hist = __brev(c.rhist); /* newest bit lowest */
      for (uint32_t b = 0; b < rem; b++) dst[(n & ~3u) + b] = (uint8_t)(hist >> (8u * (rem - 1u - b)));
    }
  }
  if (budget == 0u || dead) atomicOr(err, W3_ERR_ABORT);
}
#endif
