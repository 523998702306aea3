/*
 * w3_tierb.cuh — the reference's bit-history state table and hash-slot cells,
 * as device functions (groundwork for the hashed
 * context-mixing stack, w3_cm.cuh).
 *
 *   state table   NaiveStateTable (src/state_table/naive.rs:9-115) builds 3963
 *                 12-bit states at compile time: 3 entry nodes and 4 sub-tables
 *                 (selected by the last two bits seen) of 990 triangular count
 *                 states over 44 levels; StateTable::next / p / next4 / p4
 *                 (src/state_table/mod.rs:3-52) index it.  Here every entry is
 *                 a closed form of its index (w3_state_entry), evaluated once
 *                 per context into a table in device memory.
 *   hash slots    Cell = 6 bytes of four 12-bit check hashes + 90 bytes = 4
 *                 slots x 15 states x 12 bits (src/hashmap.rs:32-129,
 *                 docs/hashslots.md).  Same byte layout, same index formula.
 *
 * `__host__ __device__`: tests/hostsim runs the same statements on the CPU;
 * the device versions are checked against the reference's own dumps
 * (docs/state_table/{state_table,aux_state_table}.csv, docs/hashslots.md:69-133) through
 * w3_selftest_state_table / w3_selftest_hashslots.
 */
#pragma once
#include "w3_core.cuh"

#define W3_ST_MAX_LEVEL 44u                                        /* naive.rs:9 */
#define W3_ST_SUB (W3_ST_MAX_LEVEL * (W3_ST_MAX_LEVEL + 1u) / 2u)  /* :10 — 990 */
#define W3_ST_SIZE (3u + 4u * W3_ST_SUB)                           /* :11 — 3963 <= 4096 */

struct w3_state_entry {
  uint16_t prob;    /* StateEntry.prob — src/state_table/mod.rs:25-29 */
  uint16_t next[2]; /* StateEntry.next */
  uint16_t pad;     /* 8 bytes: one load per entry */
};

/* entry k of the auxiliary table (gen_auxiliary_table — naive.rs:50-75): level L holds L
 * nodes, node j = "j ones out of L - 1 bits"; indices run level by level */
W3_HD w3_state_entry w3_aux_state_entry(uint32_t k) {
  uint32_t level = 1u, filled = 0u;
  while (filled + level <= k) { filled += level; level++; } /* <= 44 steps; table build only */
  const uint32_t node = k - filled;
  w3_state_entry e;
  e.pad = 0;
  e.prob = (uint16_t)((65536u * (node + 1u)) / (level + 1u)); /* calc_prob(node, level - 1) — :107-113, floor */
  if (level == W3_ST_MAX_LEVEL) { /* get_next_nodes — :79-98: restart from the level with half the counts */
    const uint32_t next_level = (level + 2u) / 2u - 1u;
    const uint32_t nn = ((node + 2u) / 2u - 1u) + (next_level - 1u) * next_level / 2u;
    e.next[0] = (uint16_t)(node == 0u ? k : nn);
    e.next[1] = (uint16_t)(node == level - 1u ? k : nn);
  } else { /* :100-103 */
    e.next[0] = (uint16_t)(filled + level + node);
    e.next[1] = (uint16_t)(filled + level + node + 1u);
  }
  return e;
}
/* entry i of NaiveStateTable's TABLE (gen_table — naive.rs:17-48) */
W3_HD w3_state_entry w3_state_entry_of(uint32_t i) {
  const uint32_t a = 3u, b = a + W3_ST_SUB, c = a + 2u * W3_ST_SUB, d = a + 3u * W3_ST_SUB;
  w3_state_entry e;
  e.pad = 0;
  if (i < 3u) { /* entry nodes — :24-27 */
    e.prob = 0x8000u;
    e.next[0] = (uint16_t)(i == 0u ? 1u : i == 1u ? a : c);
    e.next[1] = (uint16_t)(i == 0u ? 2u : i == 1u ? b : d);
    return e;
  }
  const uint32_t sub = (i - 3u) / W3_ST_SUB; /* 0..3 = sub-tables a, b, c, d */
  const w3_state_entry x = w3_aux_state_entry((i - 3u) % W3_ST_SUB);
  e.prob = x.prob;
  /* :41-44 — a and c continue into (a, b), b and d into (c, d) */
  e.next[0] = (uint16_t)(((sub & 1u) ? c : a) + x.next[0]);
  e.next[1] = (uint16_t)(((sub & 1u) ? d : b) + x.next[1]);
  return e;
}

/* StateTable::next / ::p (mod.rs:41-49) over a built table */
W3_HD uint32_t w3_state_next(const w3_state_entry *t, uint32_t state, uint32_t bit) { return t[state].next[bit & 1u]; }
W3_HD uint32_t w3_state_p(const w3_state_entry *t, uint32_t state) { return t[state].prob; }
/* StateTable::next4 / ::p4 (mod.rs:5-22) */
W3_HD void w3_state_next4(const w3_state_entry *t, const uint16_t s[4], uint32_t nib, uint16_t out[4]) {
  out[0] = (uint16_t)w3_state_next(t, s[0], nib >> 3);
  out[1] = (uint16_t)w3_state_next(t, s[1], (nib >> 2) & 1u);
  out[2] = (uint16_t)w3_state_next(t, s[2], (nib >> 1) & 1u);
  out[3] = (uint16_t)w3_state_next(t, s[3], nib & 1u);
}
W3_HD void w3_state_p4(const w3_state_entry *t, const uint16_t s[4], uint16_t out[4]) {
  for (int i = 0; i < 4; i++) out[i] = (uint16_t)w3_state_p(t, s[i]);
}

/* ---- hash slots — src/hashmap.rs ----------------------------------------- */
#define W3_CELL_BYTES 96u  /* Cell { hashes: [u8; 6], slots: [u8; 90] } — hashmap.rs:32-35 */
#define W3_CELL_HASH_BYTES 6u

/* HashMap::get_slot (hashmap.rs:25-28): the HIGH log_cell_count bits of the hash pick the cell */
W3_HD uint64_t w3_cell_index(uint64_t hash, uint32_t log_cell_count) { return hash >> (64u - log_cell_count); }

/* Cell::get_slot (hashmap.rs:43-71): slot 3 sits in the low 12 bits of the 48-bit big-endian
 * hash field; a miss returns slot 1 and leaves the check hashes alone (":64-68 TODO") */
W3_HD uint32_t w3_cell_get_slot(const uint8_t *cell, uint64_t hash) {
  uint64_t hc = 0;
  for (uint32_t i = 0; i < W3_CELL_HASH_BYTES; i++) hc = (hc << 8) | cell[i];
  const uint64_t h = hash & 0xFFFu;
  if (h == (hc & 0xFFFu)) return 3u;
  if (h == ((hc >> 12) & 0xFFFu)) return 2u;
  if (h == ((hc >> 24) & 0xFFFu)) return 1u;
  if (h == ((hc >> 36) & 0xFFFu)) return 0u;
  return 1u;
}
/* Slot::get_idx (hashmap.rs:80-84): 12-bit state number -> byte index into slots[] and parity
 * (0: the state is the high 12 bits of the big-endian u16 there, 1: the low 12 bits) */
W3_HD uint32_t w3_slot_idx(uint32_t id, uint32_t bit_id, uint32_t nib_ctx) {
  return (3u << bit_id) + 3u * nib_ctx + 45u * id - 3u; /* >> 1 = abs_idx, & 1 = parity */
}
/* Slot::get_state (hashmap.rs:86-97) */
W3_HD uint32_t w3_slot_get_state(const uint8_t *cell, uint32_t id, uint32_t bit_id, uint32_t nib_ctx) {
  const uint32_t idx = w3_slot_idx(id, bit_id, nib_ctx);
  const uint8_t *p = cell + W3_CELL_HASH_BYTES + (idx >> 1);
  const uint32_t state = ((uint32_t)p[0] << 8) | p[1];
  return (idx & 1u) ? (state & 0xFFFu) : (state >> 4);
}
/* Slot::set_state (hashmap.rs:99-112) */
W3_HD void w3_slot_set_state(uint8_t *cell, uint32_t id, uint32_t bit_id, uint32_t nib_ctx, uint32_t new_state) {
  const uint32_t idx = w3_slot_idx(id, bit_id, nib_ctx);
  uint8_t *p = cell + W3_CELL_HASH_BYTES + (idx >> 1);
  if (!(idx & 1u)) {
    p[0] = (uint8_t)(new_state >> 4);
    p[1] = (uint8_t)(((new_state << 4) & 0xF0u) | (p[1] & 0x0Fu));
  } else {
    p[0] = (uint8_t)((new_state >> 8) | (p[0] & 0xF0u)); /* as the reference: no mask on the state's top bits */
    p[1] = (uint8_t)new_state;
  }
}
/* Slot::set_nib / get_nib (hashmap.rs:114-128): the four states a nibble walks through */
W3_HD void w3_slot_set_nib(uint8_t *cell, uint32_t id, uint32_t nib, const uint16_t s[4]) {
  w3_slot_set_state(cell, id, 0u, 0u, s[0]);
  w3_slot_set_state(cell, id, 1u, nib >> 3, s[1]);
  w3_slot_set_state(cell, id, 2u, nib >> 2, s[2]);
  w3_slot_set_state(cell, id, 3u, nib >> 1, s[3]);
}
W3_HD void w3_slot_get_nib(const uint8_t *cell, uint32_t id, uint32_t nib, uint16_t s[4]) {
  s[0] = (uint16_t)w3_slot_get_state(cell, id, 0u, 0u);
  s[1] = (uint16_t)w3_slot_get_state(cell, id, 1u, nib >> 3);
  s[2] = (uint16_t)w3_slot_get_state(cell, id, 2u, nib >> 2);
  s[3] = (uint16_t)w3_slot_get_state(cell, id, 3u, nib >> 1);
}

/* all 15 states of a slot to 0 — what 15 Slot::set_state(.., 0) calls leave behind, in 24 byte
 * stores: slot id holds bits [180 id, 180 id + 180) of slots[]; odd slots start, even slots end,
 * on a half byte shared with their neighbour (docs/hashslots.md:136-139) */
W3_HD void w3_slot_clear(uint8_t *cell, uint32_t id) {
  uint8_t *p = cell + W3_CELL_HASH_BYTES + (id * 45u >> 1); /* byte 0, 22, 45, 67 */
  if (id & 1u) {
    p[0] &= 0xF0u;
    for (uint32_t i = 1; i <= 22u; i++) p[i] = 0;
  } else {
    for (uint32_t i = 0; i < 22u; i++) p[i] = 0;
    p[22] &= 0x0Fu;
  }
}

/* OpinionMixer2::mix (src/mixers/opinion_mixer2.rs:5-10): the more confident of two probabilities */
W3_HD uint32_t w3_opinion_mix2(uint32_t p1, uint32_t p2) {
  const uint32_t d1 = p1 >= 0x8000u ? p1 - 0x8000u : 0x8000u - p1;
  const uint32_t d2 = p2 >= 0x8000u ? p2 - 0x8000u : 0x8000u - p2;
  return d1 >= d2 ? p1 : p2;
}
