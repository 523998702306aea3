/*
 * w3_core.cuh — the integer arithmetic of weath3rb0i's predict→update→code
 * path, restated for one-warp-per-block execution on sm_100a.
 *
 * Everything here is `__host__ __device__` so that tests/hostsim can compile
 * the very same statements with g++ and compare them with the oracle on the
 * CPU before a kernel ever runs; the product only ever calls them from
 * kernels (w3_kernels.cu).  Citations are relative to /root/reference/.
 *
 * What differs from the reference's formulation (and why it is still exact):
 *  - the two `while` renormalisation loops of ArithmeticCoder::encode/decode
 *    (src/entropy_coding/arithmetic_coder.rs:51-62, 91-102) are replaced by a
 *    closed form: s = clz(x1^x2) matching-prefix shifts followed by k E3
 *    shifts, k = length of the run of (1,0) bit pairs after the first
 *    differing pair.  s + k <= 31 whenever x1 != x2.
 *  - ACHistory::hash (src/history/ac_history.rs:28-46) is a monotone step
 *    function of the bit-reversed history: an arithmetic coder maps
 *    lexicographic order of its input strings to numeric order of its code,
 *    and the hash is the first 8 code bits.  It is therefore fully described,
 *    per alignment, by <= 255 sorted 64-bit thresholds; a threshold key whose
 *    own 64 bits never complete 8 code bits ("singular" key) falls back to the
 *    literal loop (w3_achash_exact).  See w3_hash_lookup.
 *  - Counter::p's u64 division (src/models/counter.rs:14-17) is a float
 *    estimate with an exact integer remainder fix-up.
 */
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define W3_HD __host__ __device__ __forceinline__
#else
#define W3_HD static inline
#endif

/* ---- StationaryModel::for_book1 — src/models/ac_hash/stationary.rs:40-42 -- */
#define W3_BOOK1_TABLE {1, 50188, 62497, 15819, 22545, 31499, 22988, 29616}

/* geometry fixed by init_model — src/main.rs:151 */
#define W3_CTX_BITS 11
#define W3_ALIGN_BITS 3
#define W3_HASH_BITS 8
#define W3_NCTX (1 << W3_CTX_BITS)

/* first-level cell index of the threshold table: top W3_PB key bits */
#define W3_PB 11
#define W3_NCELL (1 << W3_PB)
/* the decode speculators use a finer first level (fewer thresholds per cell,
 * so that a fixed number of loads resolves a lookup): top W3_PB2 key bits */
#define W3_PB2 13
#define W3_NCELL2 (1 << W3_PB2)
#define W3_FIX_THR 6 /* thresholds a fixed-latency lookup compares */

/* bits of the device error word the codec kernels OR into */
#define W3_ERR_OVERFLOW 1u /* a stream did not fit its scratch region */
#define W3_ERR_ABORT 2u    /* a kernel gave up (bounded inter-warp wait, or a length it cannot handle) */

W3_HD int w3_clz32(uint32_t v) {
#if defined(__CUDA_ARCH__)
  return __clz((int)v);
#else
  return v ? __builtin_clz(v) : 32;
#endif
}
/* clz for a non-zero word.  On the device FLO costs 23 cycles of dependent
 * latency, an int->float conversion 4.4 (tools/microbench*.cu, round 1): with
 * round-toward-zero the float's exponent is exactly floor(log2 v). */
W3_HD int w3_clz32_nz(uint32_t v) {
#if defined(__CUDA_ARCH__)
  float f;
  asm("cvt.rz.f32.u32 %0, %1;" : "=f"(f) : "r"(v));
  return 158 - (int)(__float_as_uint(f) >> 23);
#else
  return __builtin_clz(v);
#endif
}
W3_HD uint32_t w3_umulhi(uint32_t a, uint32_t b) {
#if defined(__CUDA_ARCH__)
  return __umulhi(a, b);
#else
  return (uint32_t)(((uint64_t)a * b) >> 32);
#endif
}
W3_HD uint32_t w3_brev32(uint32_t v) {
#if defined(__CUDA_ARCH__)
  return __brev(v);
#else
  v = ((v >> 1) & 0x55555555u) | ((v & 0x55555555u) << 1);
  v = ((v >> 2) & 0x33333333u) | ((v & 0x33333333u) << 2);
  v = ((v >> 4) & 0x0F0F0F0Fu) | ((v & 0x0F0F0F0Fu) << 4);
  v = ((v >> 8) & 0x00FF00FFu) | ((v & 0x00FF00FFu) << 8);
  return (v >> 16) | (v << 16);
#endif
}
/* history (newest bit = bit 0) -> key (newest bit = bit 63) */
W3_HD uint64_t w3_key_of_history(uint64_t bits) {
  return ((uint64_t)w3_brev32((uint32_t)bits) << 32) | w3_brev32((uint32_t)(bits >> 32));
}

/* ---- lerp + range update — arithmetic_coder.rs:42-48, 109-119 ------------ */
/* P is the 32-bit probability lerp() builds: prob == 0 ? 1 : prob << 16 */
W3_HD uint32_t w3_prob32(uint32_t prob) { return prob ? prob << 16 : 1u; }

W3_HD void w3_ac_narrow(uint32_t &x1, uint32_t &x2, uint32_t bit, uint32_t P) {
  uint32_t xmid = x1 + w3_umulhi(x2 - x1, P);
  if (bit) x2 = xmid;      /* 1 takes the lower sub-range (:47) */
  else x1 = xmid + 1;      /* :46 */
}

/* Both renormalisation loops (arithmetic_coder.rs:51-62 / 91-102) at once.
 * Requires x1 != x2 (always true for prob in [1,65535] on a normalised range,
 * see DESIGN.md).  On return x1/x2 are renormalised, s = matching-prefix
 * shifts (bits to emit / read plainly), k = E3 shifts; returns s + k <= 31. */
W3_HD int w3_ac_renorm(uint32_t &x1, uint32_t &x2, int &s, int &k) {
  s = w3_clz32_nz(x1 ^ x2);
  /* 0-bits of (~x1 | x2) are (1,0) pairs; after dropping the s equal pairs and
   * the first differing one, the leading zero run is the E3 count.  The `| 1`
   * lands on bit s after the shift and caps the run at the word's end. */
  uint32_t h = (((~x1 | x2) << 1) | 1u) << s;
  k = w3_clz32_nz(h);
  int n = s + k;
  x1 = (x1 << n) & 0x7FFFFFFFu;
  x2 = (x2 << n) | ~(0xFFFFFFFFu << n) | 0x80000000u;
  return n;
}

/* (hi:lo) >> min(sh, 32), low word — __funnelshift_rc */
W3_HD uint32_t w3_fsr_clamp(uint32_t lo, uint32_t hi, uint32_t sh) {
#if defined(__CUDA_ARCH__)
  return __funnelshift_rc(lo, hi, sh);
#else
  if (sh > 32u) sh = 32u;
  return (uint32_t)(((((uint64_t)hi) << 32) | lo) >> sh);
#endif
}

/* high word of (hi:lo) << (sh & 31) — __funnelshift_l */
W3_HD uint32_t w3_fsl(uint32_t lo, uint32_t hi, uint32_t sh) {
#if defined(__CUDA_ARCH__)
  return __funnelshift_l(lo, hi, sh);
#else
  sh &= 31u;
  return sh ? (hi << sh) | (lo >> (32u - sh)) : hi;
#endif
}
/* low word of (hi:lo) >> (sh & 31) — __funnelshift_r */
W3_HD uint32_t w3_fsr(uint32_t lo, uint32_t hi, uint32_t sh) {
#if defined(__CUDA_ARCH__)
  return __funnelshift_r(lo, hi, sh);
#else
  sh &= 31u;
  return sh ? (lo >> sh) | (hi << (32u - sh)) : lo;
#endif
}
W3_HD uint32_t w3_bswap32(uint32_t v) {
#if defined(__CUDA_ARCH__)
  return __byte_perm(v, 0, 0x0123);
#else
  return __builtin_bswap32(v);
#endif
}
/* read-only global word (the streams and inputs are never written by the kernel that reads them) */
W3_HD uint32_t w3_ldg32(const uint32_t *p) {
#if defined(__CUDA_ARCH__)
  return __ldg(p);
#else
  return *p;
#endif
}

/* The coder's state as (x1, w = x2 - x1), and both renormalisation loops by
 * their TOTAL shift, found with one count-leading-zeros of the width instead of
 * two chained ones over the bounds (the closed form above): 54 instead of 80
 * cycles of dependent latency per coded bit.
 *
 * Every shift of either loop (arithmetic_coder.rs:51-62) maps the width to
 * 2w + 1, and either loop runs only while w < 2^31; a state both loops have
 * left has w > 2^30 (its bounds straddle the midpoint and are not both inside
 * the middle half).  So with j0 = clz(w) the state after j0 shifts is surely
 * final, the one after j0 - 2 surely not, and the total is j0 - 1 or j0.  The
 * state after j = j0 - 1 shifts is final iff its bounds' top bits differ
 * (the plain shifts are over) and their next bits are not (1, 0) (no further
 * E3 shift) — read off the RAW shifted bounds, whose top bits are the pair an
 * E3 shift would have dropped: (0,1) if none happened, (1,0) otherwise, which
 * is also the decoder's "flip the top bit of x" (arithmetic_coder.rs:98-102).
 * Checked against the literal loops on 5.9e8 states (tests/hostsim, and
 * tools/renorm_check.cpp).  Requires w >= 1.
 *
 * narrow (:43-49) + renormalise; returns the total shift, *flip = 0x80000000
 * iff an E3 shift happened; x1p/x2p = the bounds after narrowing, before the
 * shifts (what the encoder's writer needs). */
struct w3_coder {
  uint32_t x1, w;
};
W3_HD w3_coder w3_coder_new() { return w3_coder{0u, 0xFFFFFFFFu}; }
/* `ones` = 0xFFFFFFFF: a caller in a tight loop keeps it in a register the
 * compiler cannot re-materialise (it otherwise spends a move per use).
 * kNeedShift = false: the caller wants the new state only (the encoder's coder
 * warp: every instruction of a lone warp costs 2.3-2.9 cycles). */
template <bool kNeedShift = true>
W3_HD int w3_coder_renorm(w3_coder &c, uint32_t nx1, uint32_t nw, uint32_t &flip, uint32_t ones = 0xFFFFFFFFu) {
  const uint32_t nx2 = nx1 + nw;
  const uint32_t sh = 33u - (uint32_t)w3_clz32_nz(nw); /* 64-bit right shift that leaves j = max(j0-1, 0) */
  uint32_t a = w3_fsr_clamp(0u, nx1, sh), wa = w3_fsr_clamp(ones, nw, sh); /* after j shifts */
  const uint32_t b = w3_fsr_clamp(ones, nx2, sh);
  /* final after j shifts: top bits differ, next bits are not (1, 0); else one more */
  uint32_t fin;
#if defined(__CUDA_ARCH__)
  /* spelled out: the compiler turns the C form into six instructions, this is
   * LOP3 (with predicate), LOP3, ISETP and two predicated updates */
  if (kNeedShift) {
    asm("{\n"
        ".reg .pred p, q;\n"
        ".reg .b32 y, z;\n"
        "lop3.b32 z, %0, %3, 0x40000000, 0x20;\n" /* a & ~b & bit 30 */
        "xor.b32 y, %0, %3;\n"
        "setp.ne.u32 q, z, 0;\n"
        "setp.lt.and.s32 p, y, 0, !q;\n"
        "@!p add.u32 %0, %0, %0;\n"
        "@!p mad.lo.u32 %1, %1, 2, 1;\n"
        "selp.u32 %2, 1, 0, p;\n"
        "}\n"
        : "+r"(a), "+r"(wa), "=r"(fin)
        : "r"(b));
  } else {
    asm("{\n"
        ".reg .pred p, q;\n"
        ".reg .b32 y, z;\n"
        "lop3.b32 z, %0, %2, 0x40000000, 0x20;\n"
        "xor.b32 y, %0, %2;\n"
        "setp.ne.u32 q, z, 0;\n"
        "setp.lt.and.s32 p, y, 0, !q;\n"
        "@!p add.u32 %0, %0, %0;\n"
        "@!p mad.lo.u32 %1, %1, 2, 1;\n"
        "}\n"
        : "+r"(a), "+r"(wa)
        : "r"(b));
    fin = 1u;
  }
#else
  fin = ((int32_t)(a ^ b) < 0 && ((a & ~b) & 0x40000000u) == 0u) ? 1u : 0u;
  if (!fin) {
    a = a + a;
    wa = wa + wa + 1u;
  }
#endif
  flip = a & 0x80000000u;
  c.x1 = a & 0x7FFFFFFFu;
  c.w = wa;
  const int j0 = 33 - (int)sh;
  return fin ? j0 - (j0 != 0 ? 1 : 0) : j0; /* j0 = 0: the state is final as it stands */
}
/* one coded bit: P = Counter::p << 16 of a 1 */
template <bool kNeedShift = true>
W3_HD int w3_coder_step(w3_coder &c, uint32_t bit, uint32_t P, uint32_t &flip, uint32_t &x1p, uint32_t &x2p,
                        uint32_t ones = 0xFFFFFFFFu) {
  const uint32_t m = w3_umulhi(c.w, P);
  x1p = bit ? c.x1 : c.x1 + m + 1u;   /* 1 takes the lower sub-range (:46-47) */
  const uint32_t nw = bit ? m : c.w - m - 1u;
  x2p = x1p + nw;
  return w3_coder_renorm<kNeedShift>(c, x1p, nw, flip, ones);
}

/* w3_ac_renorm with the caller's all-ones register (see w3_coder_renorm): x2's
 * new low bits come in through one funnel shift */
W3_HD int w3_ac_renorm_ones(uint32_t &x1, uint32_t &x2, int &s, int &k, uint32_t ones) {
  s = w3_clz32_nz(x1 ^ x2);
  uint32_t h = (((~x1 | x2) << 1) | 1u) << s;
  k = w3_clz32_nz(h);
  int n = s + k;
  x1 = (x1 << n) & 0x7FFFFFFFu;
#if defined(__CUDA_ARCH__)
  x2 = __funnelshift_l(ones, x2, (uint32_t)n) | 0x80000000u;
#else
  x2 = (x2 << n) | ~(ones << n) | 0x80000000u;
#endif
  return n;
}

/* ---- Counter — src/models/counter.rs ------------------------------------- */
/* Counter::p (:13-18) with c0,c1 <= 65534 */
W3_HD uint32_t w3_counter_p(uint32_t c0, uint32_t c1) {
  uint32_t a = c1 + 1u, d = c0 + c1 + 2u;
#if defined(__CUDA_ARCH__)
  float rcp; /* d >= 2: no denormal or overflow case to guard */
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(rcp) : "f"((float)d));
  uint32_t q = (uint32_t)((float)a * 131072.0f * rcp);
#else
  uint32_t q = (uint32_t)((float)a * 131072.0f / (float)d);
#endif
  /* exact remainder of a*2^17 - q*d; |rem| < 2^19 so mod-2^32 maths is exact.
   * (By the error bound the estimate is within one of the quotient and one step
   * each way would do; that shorter form measured slower in the encoder's
   * chain, 81.5 vs 80.9 cycles per bit, so the two-step form stays.) */
  int32_t rem = (int32_t)((a << 17) - q * d);
  if (rem < 0) { q -= 1u; rem += (int32_t)d; }
  if (rem < 0) { q -= 1u; rem += (int32_t)d; }
  if (rem >= (int32_t)d) { q += 1u; rem -= (int32_t)d; }
  if (rem >= (int32_t)d) { q += 1u; }
  return (q >> 1) + (q & 1u); /* rounding (:17) */
}
/* the literal statement, used by the device self-test to check the above */
W3_HD uint32_t w3_counter_p_u64(uint32_t c0, uint32_t c1) {
  uint64_t q = (((uint64_t)1 << 17) * ((uint64_t)c1 + 1)) / ((uint64_t)c0 + c1 + 2);
  return (uint32_t)((q >> 1) + (q & 1));
}
/* Counter::update (:20-26); slot = c0 | c1 << 16 */
W3_HD uint32_t w3_counter_update(uint32_t slot, uint32_t bit) {
  /* both 16-bit halves at once: c <= 65534 before, so the +1 cannot carry, and
   * (c >> 1) + (c & 1) <= 0x8000 keeps each half in its lane */
  uint32_t v = slot + (bit ? 0x10000u : 1u);
  if ((bit ? v >> 16 : v & 0xFFFFu) == 0xFFFFu) v = ((v >> 1) & 0x7FFF7FFFu) + (v & 0x00010001u);
  return v;
}

/* 1/d for the float estimates of Counter::p; d >= 2: no denormal or overflow case */
W3_HD float w3_rcp_of(uint32_t d) {
#if defined(__CUDA_ARCH__)
  float rcp;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(rcp) : "f"((float)d));
  return rcp;
#else
  return 1.0f / (float)d;
#endif
}
/* Counter::p with ONE correction each way: the float estimate of the quotient q <= 131068 carries
 * a relative error of a few 2^-23 (two roundings and rcp.approx), i.e. at most 0.05 absolute, so
 * its floor is within one of the true floor.  Ten instructions fewer than w3_counter_p; the device
 * self-test (w3_selftest_counter_p) checks all 65535^2 count pairs against the u64 statement. */
W3_HD uint32_t w3_counter_p1(uint32_t c0, uint32_t c1) {
  const uint32_t a = c1 + 1u, d = c0 + c1 + 2u;
  uint32_t q = (uint32_t)((float)a * (w3_rcp_of(d) * 131072.0f));
  const int32_t rem = (int32_t)((a << 17) - q * d);
  q = q - ((uint32_t)rem >> 31) + (rem >= (int32_t)d ? 1u : 0u);
  return (q >> 1) + (q & 1u);
}

/* Counter::p after Counter::update with a 0 and with a 1, from the counter as it
 * stands (slot = c0 | c1 << 16): what the decode speculators prepare for every
 * guessed context.  Unless the update halves (counter.rs:22-25; then the general
 * functions run) both new states share the total d = c0 + c1 + 3, hence one
 * reciprocal; and since the float estimate is within 0.03 of the quotient, its
 * floor is within one of the true one: a single correction each way.  The
 * device self-test checks every counter state against the u64 statement. */
W3_HD void w3_counter_p_both(uint32_t slot, uint32_t &p0, uint32_t &p1) {
  const uint32_t c0 = slot & 0xFFFFu, c1 = slot >> 16;
  if (c0 >= 0xFFFEu || c1 >= 0xFFFEu) { /* an update would reach 65535 and halve */
    const uint32_t n0 = w3_counter_update(slot, 0u), n1 = w3_counter_update(slot, 1u);
    p0 = w3_counter_p(n0 & 0xFFFFu, n0 >> 16);
    p1 = w3_counter_p(n1 & 0xFFFFu, n1 >> 16);
    return;
  }
  const uint32_t d = c0 + c1 + 3u;
  const float r17 = w3_rcp_of(d) * 131072.0f; /* exact scaling */
  const uint32_t a0 = c1 + 1u, a1 = c1 + 2u;
  uint32_t q0 = (uint32_t)((float)a0 * r17), q1 = (uint32_t)((float)a1 * r17);
  const int32_t rem0 = (int32_t)((a0 << 17) - q0 * d), rem1 = (int32_t)((a1 << 17) - q1 * d);
  q0 = q0 - ((uint32_t)rem0 >> 31) + (rem0 >= (int32_t)d ? 1u : 0u);
  q1 = q1 - ((uint32_t)rem1 >> 31) + (rem1 >= (int32_t)d ? 1u : 0u);
  p0 = (q0 >> 1) + (q0 & 1u);
  p1 = (q1 >> 1) + (q1 & 1u);
}

/* ---- ACHistory::hash, literal — src/history/ac_history.rs:28-88 ----------- */
struct w3_hash_result {
  uint32_t hash; /* state >> (32 - idx) (:45), 0 when idx == 0 */
  int idx;       /* code bits captured, == max_bits unless the 64 bits ran out */
  int rev;       /* pending E3 bits when the loop ended (only if idx < max_bits) */
};

W3_HD w3_hash_result w3_achash_exact(uint64_t bits, uint32_t pos, const uint16_t *table,
                                     int max_bits) {
  uint32_t y1 = 0, y2 = 0xFFFFFFFFu, state = 0; /* :29-35 */
  int idx = 0, rev = 0;
  uint32_t a = pos & 7u;                        /* :36, stationary.rs:50-52 */
  bool full = false;
  for (int i = 0; i < 64 && !full; i++) {       /* :37 */
    uint32_t bit = (uint32_t)(bits >> i) & 1u;  /* :38 */
    a = (a + 7u) & 7u;                          /* stationary.rs:55 */
    w3_ac_narrow(y1, y2, bit, w3_prob32(table[a]));
    while (((y1 ^ y2) >> 31) == 0) {            /* arithmetic_coder.rs:51-55 */
      uint32_t b = y1 >> 31;
      /* EntropyWriter::write_bit — ac_history.rs:58-80 */
      if (idx == max_bits) { full = true; break; }
      state = (state >> 1) | (b << 31); idx++;
      while (rev > 0) {
        rev--;
        if (idx == max_bits) { full = true; break; }
        state = (state >> 1) | ((b ^ 1u) << 31); idx++;
      }
      if (full) break;
      y1 <<= 1;
      y2 = (y2 << 1) | 1u;
    }
    if (full) break;                            /* :40-42 */
    while (y1 >= 0x40000000u && y2 < 0xC0000000u) { /* arithmetic_coder.rs:58-62 */
      rev++;
      y1 = (y1 << 1) & 0x7FFFFFFFu;
      y2 = (y2 << 1) | 0x80000001u;
    }
    if (idx == max_bits) break; /* the sink can no longer change */
  }
  w3_hash_result r;
  r.hash = idx == 0 ? 0u : state >> (32 - idx);
  r.idx = idx;
  r.rev = rev;
  return r;
}

/* ---- the threshold table ------------------------------------------------- */
struct w3_hashtab {
  uint64_t thr[8][256];       /* thr[a][j], j < 255 ascending; [255] = ~0 pad */
  uint16_t cell[8][W3_NCELL]; /* (#thr below the cell) | (#thr inside) << 8 */
  uint8_t rval[8][256];       /* hash of regular region r = #{j : thr[j] <= K} */
  uint8_t sing[8][256];       /* 1: thr[a][j] is a singular key */
  uint16_t cell2[8][W3_NCELL2]; /* same as cell[][] for W3_PB2 key bits */
  uint16_t table[8];          /* the StationaryModel probabilities */
  uint16_t max_bits;
  uint16_t bad;               /* set by the builder if monotonicity fails */
};

/* Rank of a key in code order, doubled: 2r for a key whose 8 code bits
 * completed (regular region r), 2r+1 for a singular key sitting on the border
 * between regions r and r+1.  Code bins descend as the key grows because a
 * 1-bit takes the lower sub-range. */
W3_HD uint32_t w3_key_rank2(uint64_t key, uint32_t a, const uint16_t *table, int max_bits) {
  uint64_t bits = w3_key_of_history(key); /* the bit reversal is an involution */
  w3_hash_result h = w3_achash_exact(bits, a, table, max_bits);
  uint32_t nb = (uint32_t)max_bits;
  uint32_t top = (1u << nb) - 1u;
  /* hash bit 0 is the first code bit: code prefix = reverse of the idx bits */
  uint32_t code = h.idx ? (w3_brev32(h.hash) >> (32 - h.idx)) : 0u;
  if (h.idx == (int)nb) return 2u * (top - code);
  /* unfinished: the interval straddles code prefix . 1 0 0 0… */
  uint32_t upper = (code << (nb - h.idx)) | (1u << (nb - 1 - h.idx));
  return 2u * (top - upper) + 1u;
}

/* smallest key whose rank2 >= 2j+1 (binary search on the monotone rank) */
W3_HD uint64_t w3_threshold_search(uint32_t a, uint32_t j, const uint16_t *table, int max_bits) {
  uint32_t target = 2u * j + 1u;
  if (w3_key_rank2(~(uint64_t)0, a, table, max_bits) < target) return ~(uint64_t)0;
  uint64_t lo = 0, hi = ~(uint64_t)0; /* invariant: rank2(hi) >= target */
  while (lo < hi) {
    uint64_t mid = lo + ((hi - lo) >> 1);
    if (w3_key_rank2(mid, a, table, max_bits) >= target) hi = mid;
    else lo = mid + 1;
  }
  return lo;
}

/* Table form of ACHistory::hash.  `thr`, `cell` may point to shared-memory
 * copies.  Returns the RANK r = #{thresholds <= key} (0..255): region r holds
 * code bin 255 - r and the hash is that code with its first bit lowest, i.e.
 * hash = brev8(255 - r).  Returns 0xFFFFFFFF when the key equals a threshold
 * (possible singular key): the caller must then use w3_achash_exact.
 * rval[][] of the global table is the same mapping evaluated with the literal
 * loop; the builder cross-checks the two (w3_build_rval). */
W3_HD uint32_t w3_rank_lookup(uint64_t key, uint32_t a, const uint64_t *thr /*[8][256]*/,
                              const uint16_t *cell /*[8][NCELL]*/) {
  uint32_t e = cell[a * W3_NCELL + (uint32_t)(key >> (64 - W3_PB))];
  uint32_t r = e & 0xFFu, cnt = e >> 8;
  const uint64_t *t = thr + a * 256u + r;
  bool eq = false;
  for (uint32_t i = 0; i < cnt; i++) { /* independent loads: no early exit on purpose */
    uint64_t b = t[i];
    r += key >= b ? 1u : 0u;
    eq |= key == b;
  }
  return eq ? 0xFFFFFFFFu : r;
}
/* Fixed-latency form for the decode speculators: W3_FIX_THR independent
 * threshold loads, no loop.  `thr` must be readable W3_FIX_THR entries past
 * its 8 x 256.  Returns 0xFFFFFFFF when the cell holds more thresholds than
 * that or the key equals one of them: the caller then uses w3_rank_lookup /
 * w3_achash_exact. */
/* In the speculators the 32 lanes of a warp look up keys that differ only in
 * their top (guessed) bits: with the natural layout every lane would hit the
 * same shared-memory bank.  The swizzled layout puts the guessed bits lowest. */
W3_HD uint32_t w3_cell2_swizzle(uint32_t c) {
  return ((c & ((1u << (W3_PB2 - 6)) - 1u)) << 6) | (c >> (W3_PB2 - 6));
}
template <bool kSwizzled>
W3_HD uint32_t w3_rank_lookup_fixed(uint64_t key, uint32_t a, const uint64_t *thr,
                                    const uint16_t *cell2 /*[8][NCELL2]*/) {
  uint32_t c = (uint32_t)(key >> (64 - W3_PB2));
  uint32_t e = cell2[a * W3_NCELL2 + (kSwizzled ? w3_cell2_swizzle(c) : c)];
  uint32_t base = e & 0xFFu, cnt = e >> 8;
  const uint64_t *t = thr + a * 256u + base;
  uint64_t b[W3_FIX_THR];
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
  for (int i = 0; i < W3_FIX_THR; i++) b[i] = t[i];
  uint32_t r = base;
  bool bad = cnt > (uint32_t)W3_FIX_THR;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
  for (int i = 0; i < W3_FIX_THR; i++) {
    bool in = (uint32_t)i < cnt;
    r += (in && key >= b[i]) ? 1u : 0u;
    bad |= in && key == b[i];
  }
  return bad ? 0xFFFFFFFFu : r;
}
W3_HD uint32_t w3_hash_of_rank(uint32_t r) { return w3_brev32(255u - r) >> 24; }
W3_HD uint32_t w3_rank_of_hash(uint32_t h) { return 255u - (w3_brev32(h) >> 24); }
W3_HD uint32_t w3_hash_lookup(uint64_t key, uint32_t a, const uint64_t *thr, const uint16_t *cell) {
  uint32_t r = w3_rank_lookup(key, a, thr, cell);
  return r == 0xFFFFFFFFu ? r : w3_hash_of_rank(r);
}

/* The kernels index the counter table by (rank << 3) | alignment instead of
 * (hash << 3) | alignment (ordern_entropy.rs:43-44).  rank <-> hash is a
 * bijection on 0..255 and the table is private to a block and starts at zero,
 * so this only relabels the 2048 counters: every predicted probability, hence
 * every output byte, is unchanged (tests/test_hostsim.py checks it against the
 * oracle).  ctx_0 = 0 (ordern_entropy.rs:20) is hash 0, alignment 0. */
#define W3_CTX0_RANKED (255u << W3_ALIGN_BITS)

/* ---- bit writer (ACWriter — src/entropy_coding/io.rs:52-101) -------------- */
/* Words are assembled MSB-first in a 64-bit accumulator and stored as
 * big-endian u32s; `store_word(index, value)` is supplied by the caller so the
 * same code runs in the kernel (lane-0 global store) and in hostsim. */
struct w3_writer {
  uint64_t acc;   /* low `nb` bits are pending output, oldest highest */
  uint32_t nb;    /* < 32 between calls */
  uint32_t words; /* complete 32-bit words produced so far */
  uint32_t rev;   /* rev_bits (io.rs:55) */
};
W3_HD void w3_writer_init(w3_writer &w) { w.acc = 0; w.nb = 0; w.words = 0; w.rev = 0; }

template <class Store>
W3_HD void w3_writer_put(w3_writer &w, uint32_t v, uint32_t c, Store &&store_word) {
  /* c in [1,32]; v holds c bits, first-written bit highest */
  w.acc = (w.acc << c) | v;
  w.nb += c;
  if (w.nb >= 32u) {
    w.nb -= 32u;
    store_word(w.words, (uint32_t)(w.acc >> w.nb));
    w.words++;
  }
}

/* one coder step's output: s plain bits = top s bits of x1_pre (the first one
 * followed by the pending inverse run, io.rs:83-88), then k new pending bits */
template <class Store>
W3_HD void w3_writer_step(w3_writer &w, uint32_t x1_pre, int s, int k, Store &&store_word) {
  if (s > 0) {
    uint32_t b = x1_pre >> 31;
    w3_writer_put(w, b, 1u, store_word);
    while (w.rev) {
      uint32_t c = w.rev < 32u ? w.rev : 32u;
      uint32_t ones = c == 32u ? 0xFFFFFFFFu : ((1u << c) - 1u);
      w3_writer_put(w, b ? 0u : ones, c, store_word);
      w.rev -= c;
    }
    if (s > 1) w3_writer_put(w, (x1_pre << 1) >> (33 - s), (uint32_t)(s - 1), store_word);
  }
  w.rev += (uint32_t)k;
}

/* ArithmeticCoder::flush → ACWriter::flush (arithmetic_coder.rs:67-71,
 * io.rs:91-100): top bit of x2 (+ pending run), then successive bits of x2 up
 * to the byte boundary.  Returns the stream length in bytes; the caller stores
 * the last partial word (`tail`, left-aligned, `tail_bytes` long). */
template <class Store>
W3_HD uint64_t w3_writer_flush(w3_writer &w, uint32_t x2, uint32_t &tail, uint32_t &tail_bytes,
                               Store &&store_word) {
  w3_writer_step(w, x2, 1, 0, store_word);
  uint32_t pad = (8u - (w.nb & 7u)) & 7u;
  if (pad) w3_writer_put(w, (x2 << 1) >> (32 - pad), pad, store_word);
  tail_bytes = w.nb >> 3;
  tail = w.nb ? (uint32_t)(w.acc << (32u - w.nb)) : 0u;
  return (uint64_t)w.words * 4u + tail_bytes;
}

/* ---- table construction: one call per (a, j) / (a, cell) / (a, r) --------
 * Run by w3_build_* kernels at context creation (one thread per call) and by
 * tests/hostsim on the CPU.  Phase 1 must complete before phases 2 and 3. */
W3_HD void w3_build_thr(w3_hashtab *tab, uint32_t a, uint32_t j) {
  if (j >= 255u) { tab->thr[a][255] = ~(uint64_t)0; tab->sing[a][255] = 0; return; }
  uint64_t b = w3_threshold_search(a, j, tab->table, tab->max_bits);
  tab->thr[a][j] = b;
  tab->sing[a][j] = (uint8_t)(w3_key_rank2(b, a, tab->table, tab->max_bits) & 1u);
}
W3_HD void w3_build_cell(w3_hashtab *tab, uint32_t a, uint32_t c) {
  uint64_t lo = (uint64_t)c << (64 - W3_PB);
  uint64_t hi = lo | (((uint64_t)1 << (64 - W3_PB)) - 1u);
  uint32_t below = 0, inside = 0;
  for (uint32_t j = 0; j < 255u; j++) {
    uint64_t b = tab->thr[a][j];
    below += b < lo ? 1u : 0u;
    inside += (b >= lo && b <= hi) ? 1u : 0u;
  }
  tab->cell[a][c] = (uint16_t)(below | (inside << 8));
}
W3_HD void w3_build_cell2(w3_hashtab *tab, uint32_t a, uint32_t c) {
  uint64_t lo = (uint64_t)c << (64 - W3_PB2);
  uint64_t hi = lo | (((uint64_t)1 << (64 - W3_PB2)) - 1u);
  uint32_t below = 0, inside = 0;
  for (uint32_t j = 0; j < 255u; j++) {
    uint64_t b = tab->thr[a][j];
    below += b < lo ? 1u : 0u;
    inside += (b >= lo && b <= hi) ? 1u : 0u;
  }
  tab->cell2[a][c] = (uint16_t)(below | (inside << 8));
}
W3_HD void w3_build_rval(w3_hashtab *tab, uint32_t a, uint32_t r) {
  /* first regular key of region r = [thr[r-1] (+1 if singular), thr[r]) */
  uint64_t lo = 0;
  bool empty = false;
  if (r > 0) {
    lo = tab->thr[a][r - 1];
    if (tab->sing[a][r - 1]) { if (lo == ~(uint64_t)0) empty = true; else lo += 1; }
  }
  if (r < 255u && !empty) {
    uint64_t b = tab->thr[a][r];
    if (lo >= b) empty = true;
  }
  if (r > 1 && tab->thr[a][r - 1] < tab->thr[a][r - 2]) tab->bad = 1; /* must ascend */
  if (empty) { tab->rval[a][r] = 0; return; }
  w3_hash_result h = w3_achash_exact(w3_key_of_history(lo), a, tab->table, tab->max_bits);
  if (h.idx != (int)tab->max_bits ||
      w3_key_rank2(lo, a, tab->table, tab->max_bits) != 2u * r)
    tab->bad = 1;
  tab->rval[a][r] = (uint8_t)h.hash;
  if (h.hash != (w3_brev32(255u - r) >> 24)) tab->bad = 1; /* analytic form used by lookups */
}
