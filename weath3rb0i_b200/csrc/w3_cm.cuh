/*
 * w3_cm.cuh — the hashed context-mixing stack "w3cm" (SURVEY.md §8 rows a19 /
 * f4): M hashed contexts per nibble in the reference's cell / slot layout
 * (src/hashmap.rs, docs/hashslots.md), bit histories through its state table
 * (src/state_table/naive.rs), a logistic mixer, an APM, and the reference's
 * arithmetic coder (src/entropy_coding/arithmetic_coder.rs) behind them.
 *
 * PARITY UNPINNED BY THE REFERENCE: it has no logistic mixer and no APM
 * (README.md:6-10 lists them as goals), and nothing there calls its hash map or
 * its state table.  The specification is oracle/w3_cm_oracle.c (hash, slot
 * replacement, stretch / squash, weight and APM arithmetic are all spelled out
 * in its header); this file must agree with it bit for bit, and tests compare
 * the two per block.
 *
 * One WARP per block, lanes = context models, as BASELINE.json's north star
 * lays it out: per nibble every lane hashes its context and fetches its 96-byte
 * cell from the block's table in HBM (3 sectors in, 3 out: the gather roofline
 * of SURVEY.md §8(d) binds here — 2 M cells per input byte); per bit the lanes
 * read their state's probability, stretch it, and a shuffle reduction forms the
 * mixer's dot product; lane 0 squashes, refines through the APM and runs the
 * coder; every lane then trains its weight and advances its state.
 *
 * Helpers below are `__host__ __device__`: tests/hostsim checks them against
 * the oracle on the CPU.
 */
#pragma once
#include "w3_tierb.cuh"

#define W3_CM_MODELS 6u
#define W3_CM_K1 0x9E3779B97F4A7C15ull
#define W3_CM_K2 0xD6E8FEB86659FD93ull
#define W3_CM_K3 0xA24BAED4963EE407ull
#define W3_CM_APM_BINS 33u
#define W3_CM_ERR_SHIFT 16
#define W3_CM_LR 6
#define W3_CM_WMAX (1 << 18)

/* round(4096 * sigma((k - 16) / 2)), k = 0..32 */
#define W3_CM_SQ33 {1, 2, 3, 6, 10, 16, 27, 45, 73, 120, 194, 310, 488, 747, 1101, 1546, 2047, 2549, 2994, 3348, 3607, 3785, 3901, 3975, 4024, 4050, 4068, 4079, 4085, 4089, 4092, 4093, 4094}

/* squash: 12-bit logistic of d / 256, linear between the 33 anchors */
W3_HD int w3_cm_squash(int d, const int16_t *sq33) {
  if (d > 2047) return 4095;
  if (d < -2047) return 0;
  const int w = d & 127, k = (d >> 7) + 16;
  return (sq33[k] * (128 - w) + sq33[k + 1] * w + 64) >> 7;
}
/* entry p of the stretch table: the smallest d with squash(d) >= p (2047 past the end) — table build only */
W3_HD int w3_cm_stretch_entry(int p, const int16_t *sq33) {
  int lo = -2047, hi = 2047; /* squash is monotone: binary search for the first d with squash(d) >= p */
  if (w3_cm_squash(hi, sq33) < p) return 2047;
  while (lo < hi) {
    const int mid = lo + ((hi - lo) >> 1);
    if (w3_cm_squash(mid, sq33) >= p) hi = mid;
    else lo = mid + 1;
  }
  return lo;
}
/* bits the first state of a slot has seen (replacement policy): 0, 1, or level + 1 */
W3_HD uint32_t w3_cm_state_bits(uint32_t s) {
  if (s == 0u) return 0u;
  if (s < 3u) return 1u;
  const uint32_t k = (s - 3u) % W3_ST_SUB;
  uint32_t level = 1u, filled = 0u;
  while (filled + level <= k) { filled += level; level++; }
  return level + 1u;
}
/* context hash of model i (the last i + 1 bytes; the low nibble also sees the high one) */
W3_HD uint64_t w3_cm_hash(uint64_t last8, uint32_t model, uint32_t low_nibble, uint32_t hi) {
  const uint32_t order = model + 1u;
  const uint64_t ctx = last8 & (order >= 8u ? ~0ull : (((uint64_t)1 << (8u * order)) - 1u));
  uint64_t x = ctx * W3_CM_K1 + (uint64_t)(model + 1u) * W3_CM_K2 + (low_nibble ? (uint64_t)(0x100u | hi) * W3_CM_K3 : 0ull);
  x ^= x >> 29;
  x *= 0xBF58476D1CE4E5B9ull;
  x ^= x >> 32;
  return x;
}
/* Cell::get_slot (src/hashmap.rs:43-71) on a hit; on a miss the slot whose first state has seen the
 * fewest bits (ties: lowest id) is cleared and takes the check hash.  bits_of = table of
 * w3_cm_state_bits per state, or NULL to compute. */
W3_HD uint32_t w3_cm_find_slot(uint8_t *cell, uint64_t hash, const uint8_t *bits_of) {
  uint64_t hc = 0;
  for (uint32_t i = 0; i < W3_CELL_HASH_BYTES; i++) hc = (hc << 8) | cell[i];
  const uint64_t chk = hash & 0xFFFu;
  if (chk == (hc & 0xFFFu)) return 3u;
  if (chk == ((hc >> 12) & 0xFFFu)) return 2u;
  if (chk == ((hc >> 24) & 0xFFFu)) return 1u;
  if (chk == ((hc >> 36) & 0xFFFu)) return 0u;
  uint32_t victim = 0u, best = 0xFFFFFFFFu;
  for (uint32_t id = 0; id < 4u; id++) {
    const uint32_t s = w3_slot_get_state(cell, id, 0u, 0u);
    const uint32_t n = bits_of ? bits_of[s] : w3_cm_state_bits(s);
    if (n < best) { best = n; victim = id; }
  }
  const uint32_t sh = 12u * (3u - victim);
  hc = (hc & ~((uint64_t)0xFFFu << sh)) | (chk << sh);
  for (uint32_t i = 0; i < W3_CELL_HASH_BYTES; i++) cell[i] = (uint8_t)(hc >> (8u * (5u - i)));
  w3_slot_clear(cell, victim); /* = Slot::set_state(bit_id, nib_ctx, 0) for all 15 states, as the oracle spells it */
  return victim;
}
/* one weight's training step */
W3_HD int32_t w3_cm_train(int32_t w, int st, int err) {
  const int32_t nw = w + ((st * err + (1 << 15)) >> W3_CM_ERR_SHIFT);
  return nw > W3_CM_WMAX ? W3_CM_WMAX : nw < -W3_CM_WMAX ? -W3_CM_WMAX : nw;
}
