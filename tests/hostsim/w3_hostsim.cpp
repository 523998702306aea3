/*
 * w3_hostsim.cpp — TEST SUPPORT ONLY.  Compiles weath3rb0i_b200/csrc/w3_core.cuh
 * (the statements the kernels execute) with g++ so that the closed-form
 * renormalisation, the threshold-table hash, the float-estimate Counter::p and
 * the word writer can be compared with the oracle on the CPU.  Nothing in the
 * product loads this library.
 */
#include <stdlib.h>
#include <string.h>
#include <vector>

#include "../../weath3rb0i_b200/csrc/w3_core.cuh"
#include "../../weath3rb0i_b200/csrc/w3_lane.cuh"
#include "../../weath3rb0i_b200/csrc/w3_tierb.cuh"
#include "../../weath3rb0i_b200/csrc/w3_cm.cuh"

extern "C" {

int hs_tab_size(void) { return (int)sizeof(w3_hashtab); }

int hs_build_table(const uint16_t *table, int max_bits, w3_hashtab *tab) {
  memset(tab, 0, sizeof *tab);
  for (int i = 0; i < 8; i++) tab->table[i] = table[i];
  tab->max_bits = (uint16_t)max_bits;
  for (uint32_t a = 0; a < 8; a++)
    for (uint32_t j = 0; j < 256; j++) w3_build_thr(tab, a, j);
  for (uint32_t a = 0; a < 8; a++) {
    for (uint32_t c = 0; c < W3_NCELL; c++) w3_build_cell(tab, a, c);
    for (uint32_t c = 0; c < W3_NCELL2; c++) w3_build_cell2(tab, a, c);
    for (uint32_t r = 0; r < 256; r++) w3_build_rval(tab, a, r);
  }
  return tab->bad;
}

uint32_t hs_hash_exact(uint64_t bits, uint32_t pos, const uint16_t *table, int max_bits,
                       int *idx) {
  w3_hash_result r = w3_achash_exact(bits, pos, table, max_bits);
  if (idx) *idx = r.idx;
  return r.hash;
}

/* table lookup with the exact fallback, as the kernels do it */
uint32_t hs_hash(const w3_hashtab *tab, uint64_t bits, uint32_t pos, int *slow) {
  uint32_t h = w3_hash_lookup(w3_key_of_history(bits), pos & 7u, &tab->thr[0][0],
                              &tab->cell[0][0]);
  if (slow) *slow = h == 0xFFFFFFFFu;
  if (h == 0xFFFFFFFFu) h = w3_achash_exact(bits, pos, tab->table, tab->max_bits).hash;
  return h;
}

/* the fixed-latency lookup of the decode speculators; -1 = needs the general path */
int hs_rank_fixed(const w3_hashtab *tab, uint64_t bits, uint32_t pos) {
  std::vector<uint64_t> thr(8 * 256 + W3_FIX_THR, ~(uint64_t)0);
  memcpy(thr.data(), &tab->thr[0][0], sizeof tab->thr);
  uint32_t r = w3_rank_lookup_fixed<false>(w3_key_of_history(bits), pos & 7u, thr.data(), &tab->cell2[0][0]);
  return r == 0xFFFFFFFFu ? -1 : (int)r;
}
int hs_rank(const w3_hashtab *tab, uint64_t bits, uint32_t pos) {
  uint32_t r = w3_rank_lookup(w3_key_of_history(bits), pos & 7u, &tab->thr[0][0], &tab->cell[0][0]);
  return r == 0xFFFFFFFFu ? -1 : (int)r;
}

/* the relabelled context the kernels use: (rank << 3) | alignment */
static uint32_t ranked_ctx(const w3_hashtab *tab, uint64_t bits, uint32_t pos) {
  uint32_t r = w3_rank_lookup(w3_key_of_history(bits), pos & 7u, &tab->thr[0][0], &tab->cell[0][0]);
  if (r == 0xFFFFFFFFu) r = w3_rank_of_hash(w3_achash_exact(bits, pos, tab->table, tab->max_bits).hash);
  return (r << 3) | (pos & 7u);
}

/* number of thresholds compared for this key (cell occupancy statistics) */
uint32_t hs_cell_count(const w3_hashtab *tab, uint64_t bits, uint32_t pos) {
  uint64_t key = w3_key_of_history(bits);
  return tab->cell[pos & 7u][key >> (64 - W3_PB)] >> 8;
}

uint32_t hs_counter_p(uint32_t c0, uint32_t c1) { return w3_counter_p(c0, c1); }
/* returns number of mismatches of the float-estimate p against the u64 form */
uint64_t hs_counter_p_sweep(uint32_t c0_lo, uint32_t c0_hi, uint32_t step) {
  uint64_t bad = 0;
  for (uint32_t c0 = c0_lo; c0 < c0_hi; c0 += step)
    for (uint32_t c1 = 0; c1 <= 65534; c1++) {
      bad += w3_counter_p(c0, c1) != w3_counter_p_u64(c0, c1);
      bad += w3_counter_p1(c0, c1) != w3_counter_p_u64(c0, c1);
      const uint32_t slot = c0 | (c1 << 16);
      uint32_t p0, p1;
      w3_counter_p_both(slot, p0, p1);
      const uint32_t n0 = w3_counter_update(slot, 0u), n1 = w3_counter_update(slot, 1u);
      bad += p0 != w3_counter_p_u64(n0 & 0xFFFFu, n0 >> 16);
      bad += p1 != w3_counter_p_u64(n1 & 0xFFFFu, n1 >> 16);
    }
  return bad;
}
uint32_t hs_counter_update(uint32_t slot, uint32_t bit) { return w3_counter_update(slot, bit); }

/* scalar encode of one block, built only from w3_core.cuh statements */
static int g_ranked = 0; /* 1: index counters by rank, as the kernels do */
void hs_set_ranked(int on) { g_ranked = on; }
static int g_coder = 0; /* 1: the width-carrying coder step (w3_coder_step), as the team kernels do */
void hs_set_coder(int on) { g_coder = on; }

int hs_compress_block(const w3_hashtab *tab, const uint8_t *in, size_t n, uint8_t *out, size_t cap,
                      size_t *out_len) {
  std::vector<uint32_t> stats(W3_NCTX, 0u);
  uint32_t x1 = 0, x2 = 0xFFFFFFFFu, ctx = g_ranked ? W3_CTX0_RANKED : 0;
  w3_coder cw = w3_coder_new();
  uint64_t hist = 0;
  w3_writer w;
  w3_writer_init(w);
  auto store = [&](uint32_t idx, uint32_t v) {
    for (int b = 0; b < 4; b++)
      if ((size_t)idx * 4 + b < cap) out[(size_t)idx * 4 + b] = (uint8_t)(v >> (24 - 8 * b));
  };
  uint64_t t = 0;
  for (size_t i = 0; i < n; i++)
    for (int kbit = 7; kbit >= 0; kbit--) {
      uint32_t bit = (in[i] >> kbit) & 1u;
      uint32_t slot = stats[ctx];
      uint32_t p = w3_counter_p(slot & 0xFFFFu, slot >> 16);
      stats[ctx] = w3_counter_update(slot, bit);
      hist = (hist << 1) | bit;
      t++;
      ctx = g_ranked ? ranked_ctx(tab, hist, (uint32_t)t)
                     : (hs_hash(tab, hist, (uint32_t)t, nullptr) << 3) | (uint32_t)(t & 7u);
      if (g_coder) {
        uint32_t flip, x1p, x2p;
        const int nsh = w3_coder_step(cw, bit, w3_prob32(p), flip, x1p, x2p);
        const int s = w3_clz32_nz(x1p ^ x2p); /* the bounds' common prefix: what is emitted plainly */
        w3_writer_step(w, x1p, s, nsh - s, store);
        x2 = cw.x1 + cw.w;
        continue;
      }
      w3_ac_narrow(x1, x2, bit, w3_prob32(p));
      uint32_t pre = x1;
      int s, k;
      w3_ac_renorm(x1, x2, s, k);
      w3_writer_step(w, pre, s, k, store);
    }
  uint32_t tail, tb;
  uint64_t len = w3_writer_flush(w, x2, tail, tb, store);
  for (uint32_t b = 0; b < tb; b++)
    if ((size_t)w.words * 4 + b < cap) out[(size_t)w.words * 4 + b] = (uint8_t)(tail >> (24 - 8 * b));
  *out_len = (size_t)len;
  return len > cap ? -1 : 0;
}

/* (p << 16) | bit for every coded bit of a block: the coder's input sequence
 * (used by tools/microbench to feed realistic statistics) */
void hs_trace_block(const w3_hashtab *tab, const uint8_t *in, size_t n, uint32_t *entries) {
  std::vector<uint32_t> stats(W3_NCTX, 0u);
  uint32_t ctx = 0;
  uint64_t hist = 0, t = 0;
  for (size_t i = 0; i < n; i++)
    for (int kbit = 7; kbit >= 0; kbit--) {
      uint32_t bit = (in[i] >> kbit) & 1u;
      uint32_t slot = stats[ctx];
      uint32_t p = w3_counter_p(slot & 0xFFFFu, slot >> 16);
      stats[ctx] = w3_counter_update(slot, bit);
      hist = (hist << 1) | bit;
      entries[t++] = (p << 16) | bit;
      ctx = (hs_hash(tab, hist, (uint32_t)t, nullptr) << 3) | (uint32_t)(t & 7u);
    }
}

/* scalar decode of one block stream */
int hs_decompress_block(const w3_hashtab *tab, const uint8_t *in, size_t n_in, uint8_t *out,
                        size_t n_out) {
  std::vector<uint32_t> stats(W3_NCTX, 0u);
  uint32_t x1 = 0, x2 = 0xFFFFFFFFu, ctx = g_ranked ? W3_CTX0_RANKED : 0;
  w3_coder cw = w3_coder_new();
  uint64_t hist = 0, t = 0;
  size_t bitpos = 0;
  auto take = [&](int c) -> uint32_t {
    uint32_t v = 0;
    for (int i = 0; i < c; i++, bitpos++) {
      size_t by = bitpos >> 3;
      uint32_t b = by < n_in ? (in[by] >> (7 - (bitpos & 7))) & 1u : 0u;
      v = (v << 1) | b;
    }
    return v;
  };
  uint32_t x = take(16) << 16;
  x |= take(16);
  for (size_t i = 0; i < n_out; i++) {
    uint32_t byte = 0;
    for (int kbit = 0; kbit < 8; kbit++) {
      uint32_t slot = stats[ctx];
      uint32_t P = w3_prob32(w3_counter_p(slot & 0xFFFFu, slot >> 16));
      uint32_t bit;
      if (g_coder) { /* as the team kernel's decoder warp */
        const uint32_t m = w3_umulhi(cw.w, P);
        bit = (x - cw.x1) <= m ? 1u : 0u;
        uint32_t flip;
        const int nsh = w3_coder_renorm(cw, bit ? cw.x1 : cw.x1 + m + 1u, bit ? m : cw.w - m - 1u, flip);
        if (nsh) x = ((x << nsh) | take(nsh)) ^ flip;
      } else {
        uint32_t xmid = x1 + w3_umulhi(x2 - x1, P);
        bit = x <= xmid ? 1u : 0u;
        if (bit) x2 = xmid; else x1 = xmid + 1;
        int s, k;
        int nsh = w3_ac_renorm(x1, x2, s, k);
        if (nsh) x = ((x << nsh) | take(nsh)) ^ (k ? 0x80000000u : 0u);
      }
      stats[ctx] = w3_counter_update(slot, bit);
      hist = (hist << 1) | bit;
      t++;
      ctx = g_ranked ? ranked_ctx(tab, hist, (uint32_t)t)
                     : (hs_hash(tab, hist, (uint32_t)t, nullptr) << 3) | (uint32_t)(t & 7u);
      byte = (byte << 1) | bit;
    }
    out[i] = (uint8_t)byte;
  }
  return 0;
}

/* the lane decoder (w3_lane.cuh: what every lane of w3_decode_blocks_lane executes), on a copy of
 * the stream placed `misalign` bytes past a word boundary and followed by `junk` (what the next
 * block's stream would be in a packed buffer) */
struct hs_lane_env {
  std::vector<w3_lane_tab> lt;
  std::vector<uint32_t> buf;
  uint8_t *sp;
  hs_lane_env(const w3_hashtab *tab, const uint8_t *in, size_t n_in, unsigned misalign, uint8_t junk)
      : lt(1), buf((n_in + 64) / 4 + 8, 0x01010101u * junk) {
    for (uint32_t i = 0; i < 8u * W3_LANE_ROW; i++) lt[0].thi31[i] = w3_lane_tab_thi31(tab, i);
    for (uint32_t i = 0; i < 8u * W3_LANE_NCELL; i++) lt[0].cellb[i] = w3_lane_tab_cellb(tab, i);
    sp = reinterpret_cast<uint8_t *>(buf.data()) + 4 + (misalign & 3u);
    memcpy(sp, in, n_in);
  }
  w3_lane_mem mem(uint32_t *cnt, const w3_hashtab *tab) {
    w3_lane_mem m;
    m.thi31 = (w3_sptr)lt[0].thi31;
    m.cellb = (w3_sptr)lt[0].cellb;
    m.cnt = (w3_sptr)cnt;
    m.g = tab;
    m.ones = 0xFFFFFFFFu;
    return m;
  }
};
int hs_lane_decode_block(const w3_hashtab *tab, const uint8_t *in, size_t n_in, unsigned misalign,
                         uint8_t junk, uint8_t *out, size_t n_out, int /*unused*/) {
  hs_lane_env env(tab, in, n_in, misalign, junk);
  std::vector<uint32_t> cnt(W3_NCTX, 0u);
  std::vector<uint32_t> dst((n_out + 3) / 4 + 1, 0u);
  w3_lane_decode_block(env.sp, (uint32_t)n_in, (uint32_t)n_out, reinterpret_cast<uint8_t *>(dst.data()),
                       env.mem(cnt.data(), tab));
  memcpy(out, dst.data(), n_out);
  return 0;
}

/* the pair form of the lane decoder: both lanes of a pair in lockstep, the pick (a shuffle on the
 * device) done by hand.  Each lane applies the (identical) counter update, as on the device; so
 * that lane 1's does not see lane 0's, each has its own copy, and they must stay equal. */
} /* extern "C" */
template <bool kCacheP>
static int hs_pair_decode(const w3_hashtab *tab, const uint8_t *in, size_t n_in, unsigned misalign,
                          uint8_t junk, uint8_t *out, size_t n_out) {
  hs_lane_env env(tab, in, n_in, misalign, junk);
  const uint32_t sh = (uint32_t)((uintptr_t)env.sp & 3u);
  const uint32_t *ab = (const uint32_t *)(env.sp - sh);
  const uint32_t bp0 = sh * 8u, end_bits = bp0 + (uint32_t)n_in * 8u;
  const size_t words = kCacheP ? 2 * W3_NCTX : W3_NCTX;
  std::vector<uint32_t> cnt0(words, 0u), cnt1(words, 0u);
  if (kCacheP)
    for (size_t i = W3_NCTX; i < words; i++) cnt0[i] = cnt1[i] = 0x80000000u;
  const w3_lane_mem m[2] = {env.mem(cnt0.data(), tab), env.mem(cnt1.data(), tab)};
  w3_pair_dec s[2];
  w3_pair_cand cand[2];
  for (uint32_t g = 0; g < 2; g++) {
    w3_pair_init(s[g], ab, bp0, end_bits);
    cand[g] = w3_pair_guess<kCacheP>(s[g], g, 1u, m[g]);
  }
  for (size_t i = 0; i < n_out; i++) {
    uint32_t byte = 0;
    const bool fast = s[0].c.bp + W3_LANE_FAST_MARGIN <= end_bits;
    for (uint32_t a = 0; a < 8u; a++) {
      uint32_t bit[2];
      for (uint32_t g = 0; g < 2; g++) { /* F, then the late repair, per lane, in the kernel's order */
        bit[g] = fast ? w3_pair_front<false, kCacheP>(s[g], m[g], ab, end_bits)
                      : w3_pair_front<true, kCacheP>(s[g], m[g], ab, end_bits);
        if (cand[g].bad) w3_pair_repair<kCacheP>(cand[g], (a + 1u) & 7u, m[g]);
      }
      if (bit[0] != bit[1]) return -1;
      const uint32_t ci = cand[bit[0]].ci, P = cand[bit[0]].P; /* the pick */
      for (uint32_t g = 0; g < 2; g++) {
        s[g].ci = ci;
        s[g].P = P;
        cand[g] = w3_pair_guess<kCacheP>(s[g], g, (a + 2u) & 7u, m[g]); /* B */
      }
      byte = (byte << 1) | bit[0];
    }
    out[i] = (uint8_t)byte;
  }
  return cnt0 == cnt1 ? 0 : -2;
}
/* the split form (w3_decode_blocks_split): one block's coder lane and its four lanes in each of the eight model
 * warps, as independent actors scheduled at random (seeded): whatever the interleaving, the sequence numbers in
 * the ring words must keep the result right.  An actor whose word is not there yet does nothing (on the device it
 * polls).  Returns -6 if nothing can move any more (a deadlock). */
static int hs_split_decode(const w3_hashtab *tab, const uint8_t *in, size_t n_in, unsigned misalign,
                           uint8_t junk, uint8_t *out, size_t n_out, uint32_t seed) {
  hs_lane_env env(tab, in, n_in, misalign, junk);
  const uint32_t sh = (uint32_t)((uintptr_t)env.sp & 3u);
  const uint32_t *ab = (const uint32_t *)(env.sp - sh);
  const uint32_t bp0 = sh * 8u, end_bits = bp0 + (uint32_t)n_in * 8u;
  const uint32_t nbits = 8u * (uint32_t)n_out, blk = 3u; /* any column of the rings */
  std::vector<uint32_t> cnt[4]; /* per lane of the block's four */
  w3_lane_mem m[4];
  for (uint32_t g = 0; g < 4; g++) {
    cnt[g].assign(2 * W3_NCTX, 0u);
    for (size_t i = W3_NCTX; i < 2 * W3_NCTX; i++) cnt[g][i] = 0x80000000u;
    m[g] = env.mem(cnt[g].data(), tab);
  }
  std::vector<w3_split_ring> ringv(1);
  memset(&ringv[0], 0, sizeof(w3_split_ring));
  const w3_sptr cand_base = (w3_sptr)&ringv[0].cand[0][0][0], pub_base = (w3_sptr)&ringv[0].pub[0][0][0];
  const w3_sptr cand_blk = cand_base + 16u * blk, pub_blk = pub_base + 8u * blk;
  auto pub_at = [&](uint32_t count, uint32_t &rh) -> bool {
    uint32_t t;
    w3_s_ld64v(pub_blk + (count & 7u) * 64u, rh, t);
    return t == count;
  };
  w3_lane_coder c;
  w3_lane_coder_init(c, ab, bp0, end_bits);
  uint32_t rhist = 0, t = 0, byte = 0;
  w3_split_mod mod[W3_SPLIT_MODELS][4];
  uint32_t next_q[W3_SPLIT_MODELS], phase[W3_SPLIT_MODELS]; /* 0: adapt q - 8; 1: prepare q; 2: finish q */
  for (uint32_t w = 0; w < W3_SPLIT_MODELS; w++) {
    next_q[w] = w;
    phase[w] = 0u;
    for (uint32_t g = 0; g < 4; g++) w3_split_mod_init(mod[w][g]);
  }
  uint64_t rng = seed * 0x9E3779B97F4A7C15ull + 12345u;
  uint32_t idle = 0;
  while (t < nbits) {
    rng = rng * 6364136223846793005ull + 1442695040888963407ull;
    const uint32_t actor = (uint32_t)(rng >> 33) % (W3_SPLIT_MODELS + (seed & 1u ? 1u : W3_SPLIT_MODELS)); /* odd seeds: a slow coder; even: slow models */
    bool moved = false;
    if (actor >= W3_SPLIT_MODELS) { /* the coder: a step if its candidate is there */
      const uint32_t cw = w3_s_ld32v(w3_split_cand_at(cand_blk, t, rhist));
      if ((cw & 0xFFFFu) == w3_split_tag(t)) {
        const uint32_t a = t & 7u;
        uint32_t bit;
        /* the kernel decides fast / careful per byte, for the whole warp */
        if (a == 0u) byte = 0u;
        if (c.bp + W3_LANE_FAST_MARGIN <= end_bits) bit = w3_split_coder_step<false>(c, rhist, t, cw, pub_blk, ab, end_bits, 0xFFFFFFFFu);
        else bit = w3_split_coder_step<true>(c, rhist, t, cw, pub_blk, ab, end_bits, 0xFFFFFFFFu);
        byte = (byte << 1) | bit;
        if (a == 7u) out[t >> 3] = (uint8_t)byte;
        t++;
        moved = true;
      }
    } else {
      const uint32_t W = actor, q = next_q[W];
      uint32_t rh = 0u;
      if (phase[W] == 0u) {
        if (q < 8u || pub_at(q - 7u, rh)) {
          if (q >= 8u) {
            const uint32_t ci_old = mod[W][(rh >> 29) & 3u].ci; /* the shuffle */
            for (uint32_t g = 0; g < 4; g++) w3_split_adapt(ci_old, rh >> 31, m[g]);
          }
          phase[W] = 1u;
          moved = true;
        }
      } else if (phase[W] == 1u) {
        if (q < 3u || pub_at(q - 3u, rh)) {
          if (q < 3u) rh = 0u;
          if (q != 0u)
            for (uint32_t g = 0; g < 4; g++) w3_split_prepare(mod[W][g], rh, q, g, m[g]);
          phase[W] = 2u;
          moved = true;
        }
      } else {
        if (q < 2u || pub_at(q - 2u, rh)) {
          if (q < 2u) rh = 0u;
          for (uint32_t g = 0; g < 4; g++) w3_split_finish(mod[W][g], rh, q, g, blk, true, cand_base);
          next_q[W] = q + W3_SPLIT_MODELS;
          phase[W] = 0u;
          moved = true;
        }
      }
    }
    idle = moved ? 0u : idle + 1u;
    if (idle > 100000u) return -6;
  }
  for (uint32_t g = 1; g < 4; g++)
    if (cnt[g] != cnt[0]) return -2;
  return 0;
}
extern "C" {
int hs_split_decode_block(const w3_hashtab *tab, const uint8_t *in, size_t n_in, unsigned misalign,
                          uint8_t junk, uint8_t *out, size_t n_out, int seed) {
  return hs_split_decode(tab, in, n_in, misalign, junk, out, n_out, (uint32_t)seed);
}
/* the pair form of the lane decoder: both lanes of a pair in lockstep, the pick (a shuffle on the
 * device) done by hand.  Each lane applies the (identical) counter update, as on the device; so
 * that lane 1's does not see lane 0's, each has its own copy, and they must stay equal.
 * cache_p: the variant with the probability table (w3_decode_blocks_pair<true>). */
int hs_pair_decode_block(const w3_hashtab *tab, const uint8_t *in, size_t n_in, unsigned misalign,
                         uint8_t junk, uint8_t *out, size_t n_out, int cache_p) {
  return cache_p ? hs_pair_decode<true>(tab, in, n_in, misalign, junk, out, n_out)
                 : hs_pair_decode<false>(tab, in, n_in, misalign, junk, out, n_out);
}

/* ---- w3_tierb.cuh on the CPU -------------------------------------------------------------- */
void hs_state_entry(uint32_t which, uint32_t i, uint16_t out[3]) {
  const w3_state_entry e = which ? w3_aux_state_entry(i) : w3_state_entry_of(i);
  out[0] = e.next[0]; out[1] = e.next[1]; out[2] = e.prob;
}
uint32_t hs_slot_idx(uint32_t id, uint32_t bit_id, uint32_t nib_ctx) { return w3_slot_idx(id, bit_id, nib_ctx); }
uint32_t hs_slot_get_state(const uint8_t *cell, uint32_t id, uint32_t bit_id, uint32_t nib_ctx) { return w3_slot_get_state(cell, id, bit_id, nib_ctx); }
void hs_slot_set_state(uint8_t *cell, uint32_t id, uint32_t bit_id, uint32_t nib_ctx, uint32_t st) { w3_slot_set_state(cell, id, bit_id, nib_ctx, st); }
uint32_t hs_cell_get_slot(const uint8_t *cell, uint64_t hash) { return w3_cell_get_slot(cell, hash); }
void hs_slot_set_nib(uint8_t *cell, uint32_t id, uint32_t nib, const uint16_t *s) { w3_slot_set_nib(cell, id, nib, s); }
void hs_slot_get_nib(const uint8_t *cell, uint32_t id, uint32_t nib, uint16_t *s) { w3_slot_get_nib(cell, id, nib, s); }
uint64_t hs_cell_index(uint64_t hash, uint32_t log_cell_count) { return w3_cell_index(hash, log_cell_count); }
uint32_t hs_opinion_mix2(uint32_t p1, uint32_t p2) { return w3_opinion_mix2(p1, p2); }

/* ---- w3_cm.cuh on the CPU ----------------------------------------------------------------- */
static const int16_t HS_SQ33[33] = W3_CM_SQ33;
int hs_cm_squash(int d) { return w3_cm_squash(d, HS_SQ33); }
int hs_cm_stretch(int p) { return w3_cm_stretch_entry(p, HS_SQ33); }
uint32_t hs_cm_state_bits(uint32_t s) { return w3_cm_state_bits(s); }
uint64_t hs_cm_hash(uint64_t last8, uint32_t model, uint32_t low, uint32_t hi) { return w3_cm_hash(last8, model, low, hi); }
uint32_t hs_cm_find_slot(uint8_t *cell, uint64_t hash) { return w3_cm_find_slot(cell, hash, nullptr); }
int32_t hs_cm_train(int32_t w, int st, int err) { return w3_cm_train(w, st, err); }

} /* extern "C" */
