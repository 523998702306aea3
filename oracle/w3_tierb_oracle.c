/*
 * w3_tierb_oracle.c — CPU restatement of the reference's state table and hash
 * slots.  TEST INFRASTRUCTURE ONLY (see
 * w3_oracle.h): the checker for weath3rb0i_b200/csrc/w3_tierb.cuh.
 *
 * Both pieces are dead code in the reference (private modules without a
 * caller, src/lib.rs:11-13), so what is pinned is their unit behaviour:
 *   - the table NaiveStateTable builds at compile time (src/state_table/
 *     naive.rs:9-115), dumped by the reference as docs/state_table/
 *     state_table.csv (3963 rows) and aux_state_table.csv (990 rows);
 *   - Slot::get_idx and the cell layout (src/hashmap.rs:32-129), tabulated in
 *     docs/hashslots.md:69-133 (60 rows).
 * tests/test_tierb.py checks this file against those tables (committed as
 * tests/golden/tier_b_golden.json by tests/golden/make_tierb_golden.py).
 *
 * All citations are relative to /root/reference/.
 */
#include <math.h>
#include <string.h>

#include "w3_oracle.h"

/* ------------------------------------------------------------------------ */
/* NaiveStateTable — src/state_table/naive.rs                                */

#define MAX_LEVEL 44                                  /* naive.rs:9 */
#define SUBTABLE_SIZE (MAX_LEVEL * (MAX_LEVEL + 1) / 2) /* :10 — 990 */
#define ST_SIZE (3 + 4 * SUBTABLE_SIZE)               /* :11 — 3963 */
#define ST_HALF 0x8000u                               /* :14 */
#define ST_OFFSET SUBTABLE_SIZE                       /* :15 */

typedef struct {
  uint16_t prob;
  uint16_t next[2];
} st_entry; /* StateEntry — src/state_table/mod.rs:25-35 */

/* calc_prob — naive.rs:107-113: floor, not rounded */
static uint16_t calc_prob(uint16_t count, uint16_t total) {
  uint64_t c1 = count, t = total;
  uint64_t p = ((uint64_t)1 << 16) * (c1 + 1) / (t + 2);
  return (uint16_t)p;
}

/* get_next_nodes — naive.rs:77-104 */
static void get_next_nodes(unsigned level, unsigned filled, unsigned node, uint16_t next[2]) {
  if (level == MAX_LEVEL) { /* :79-98: fall back to the level with half the counts */
    unsigned next_level = ((level + 2) / 2) - 1; /* 22 */
    unsigned next_node_idx = ((node + 2) / 2) - 1;
    uint16_t curr_node = (uint16_t)(filled + node);
    uint16_t next_node = (uint16_t)(next_node_idx + (next_level - 1) * next_level / 2);
    next[0] = next[1] = next_node;
    if (node == 0) next[0] = curr_node;
    if (node == level - 1) next[1] = curr_node;
    return;
  }
  uint16_t nn = (uint16_t)(filled + level + node); /* :100-103 */
  next[0] = nn;
  next[1] = (uint16_t)(nn + 1);
}

/* gen_auxiliary_table — naive.rs:50-75 */
static void gen_auxiliary_table(st_entry at[SUBTABLE_SIZE]) {
  unsigned level = 1, filled = 0;
  while (level <= MAX_LEVEL) {
    for (unsigned node = 0; node < level; node++) {
      at[filled + node].prob = calc_prob((uint16_t)node, (uint16_t)(level - 1));
      get_next_nodes(level, filled, node, at[filled + node].next);
    }
    filled += level;
    level += 1;
  }
}

/* gen_table — naive.rs:17-48 */
static void gen_table(st_entry t[ST_SIZE]) {
  const uint16_t a = 3, b = a + ST_OFFSET, c = a + 2 * ST_OFFSET, d = a + 3 * ST_OFFSET;
  memset(t, 0, sizeof(st_entry) * ST_SIZE);
  t[0] = (st_entry){ST_HALF, {1, 2}}; /* entry nodes :24-27 */
  t[1] = (st_entry){ST_HALF, {a, b}};
  t[2] = (st_entry){ST_HALF, {c, d}};
  static st_entry at[SUBTABLE_SIZE];
  gen_auxiliary_table(at);
  for (unsigned i = 0; i < SUBTABLE_SIZE; i++) { /* :37-45 */
    const uint16_t *nx = at[i].next;
    uint16_t p = at[i].prob;
    t[a + i] = (st_entry){p, {(uint16_t)(a + nx[0]), (uint16_t)(b + nx[1])}};
    t[b + i] = (st_entry){p, {(uint16_t)(c + nx[0]), (uint16_t)(d + nx[1])}};
    t[c + i] = (st_entry){p, {(uint16_t)(a + nx[0]), (uint16_t)(b + nx[1])}};
    t[d + i] = (st_entry){p, {(uint16_t)(c + nx[0]), (uint16_t)(d + nx[1])}};
  }
}

static st_entry g_table[ST_SIZE];
static st_entry g_aux[SUBTABLE_SIZE];
static int g_table_ready = 0;
static void table_init(void) {
  if (g_table_ready) return;
  gen_table(g_table);
  gen_auxiliary_table(g_aux);
  g_table_ready = 1;
}

int w3o_state_table_size(void) { return ST_SIZE; }
int w3o_aux_state_table_size(void) { return SUBTABLE_SIZE; }

/* StateTable::next / ::p — src/state_table/mod.rs:41-49 */
uint16_t w3o_state_next(uint16_t state, int bit) {
  table_init();
  return g_table[state].next[bit & 1];
}
uint16_t w3o_state_p(uint16_t state) {
  table_init();
  return g_table[state].prob;
}
/* the auxiliary table the reference also dumps (docs/state_table/aux_state_table.csv) */
uint16_t w3o_aux_state_next(uint16_t state, int bit) {
  table_init();
  return g_aux[state].next[bit & 1];
}
uint16_t w3o_aux_state_p(uint16_t state) {
  table_init();
  return g_aux[state].prob;
}
/* StateTable::next4 / ::p4 — src/state_table/mod.rs:5-22 */
void w3o_state_next4(const uint16_t states[4], uint8_t nib, uint16_t out[4]) {
  out[0] = w3o_state_next(states[0], nib >> 3);
  out[1] = w3o_state_next(states[1], (nib >> 2) & 1);
  out[2] = w3o_state_next(states[2], (nib >> 1) & 1);
  out[3] = w3o_state_next(states[3], nib & 1);
}
void w3o_state_p4(const uint16_t states[4], uint16_t out[4]) {
  for (int i = 0; i < 4; i++) out[i] = w3o_state_p(states[i]);
}

/* ------------------------------------------------------------------------ */
/* HashMap / Cell / Slot — src/hashmap.rs                                    */

/* HashMap::new — hashmap.rs:7-22: log_cell_count from f64 logarithms */
uint32_t w3o_hashmap_log_cell_count(uint64_t size) {
  double cell_size = 96.0; /* size_of::<Cell>() = 6 + 90 */
  return (uint32_t)(log2((double)size) - log2(cell_size));
}
/* HashMap::get_slot — hashmap.rs:25-28: the HIGH bits of the hash pick the cell */
uint64_t w3o_hashmap_cell_index(uint64_t hash, uint32_t log_cell_count) {
  return hash >> (64 - log_cell_count);
}

/* Cell::get_slot — hashmap.rs:43-71: cell = hashes[6] | slots[90]; returns the slot id */
int w3o_cell_get_slot(const uint8_t cell[96], uint64_t hash) {
  uint64_t hc = 0;
  for (int i = 0; i < 6; i++) hc = (hc << 8) | cell[i]; /* from_be_bytes([0,0,h0..h5]) */
  const uint64_t mask = (1u << 12) - 1;
  const uint64_t h = hash & mask;
  if (h == (hc & mask)) return 3;
  if (h == ((hc >> 12) & mask)) return 2;
  if (h == ((hc >> 24) & mask)) return 1;
  if (h == ((hc >> 36) & mask)) return 0;
  return 1; /* :64-68 "TODO: Select min and set new hash" — the check hash is NOT stored */
}

/* Slot::get_idx — hashmap.rs:80-84 */
void w3o_slot_get_idx(int id, int bit_id, int nib_ctx, uint32_t *abs_idx, int *parity) {
  uint32_t idx = (3u << bit_id) + 3u * (uint32_t)nib_ctx + 45u * (uint32_t)id - 3u;
  *parity = (int)(idx & 1u);
  *abs_idx = idx >> 1;
}
/* Slot::get_state — hashmap.rs:86-97 */
uint16_t w3o_slot_get_state(const uint8_t cell[96], int id, int bit_id, int nib_ctx) {
  uint32_t abs_idx;
  int parity;
  w3o_slot_get_idx(id, bit_id, nib_ctx, &abs_idx, &parity);
  const uint8_t *slots = cell + 6;
  uint16_t state = (uint16_t)((slots[abs_idx] << 8) | slots[abs_idx + 1]); /* from_be_bytes */
  return parity ? (uint16_t)(state & 0xFFF) : (uint16_t)(state >> 4);
}
/* Slot::set_state — hashmap.rs:99-112 */
void w3o_slot_set_state(uint8_t cell[96], int id, int bit_id, int nib_ctx, uint16_t new_state) {
  uint32_t abs_idx;
  int parity;
  w3o_slot_get_idx(id, bit_id, nib_ctx, &abs_idx, &parity);
  uint8_t *slots = cell + 6;
  if (!parity) {
    uint16_t v = (uint16_t)(new_state << 4);
    slots[abs_idx] = (uint8_t)(v >> 8);
    slots[abs_idx + 1] = (uint8_t)((v & 0xFF) | (slots[abs_idx + 1] & 15));
  } else {
    slots[abs_idx] = (uint8_t)((new_state >> 8) | (slots[abs_idx] & (15 << 4)));
    slots[abs_idx + 1] = (uint8_t)(new_state & 0xFF);
  }
}
/* Slot::set_nib / get_nib — hashmap.rs:114-128 */
void w3o_slot_set_nib(uint8_t cell[96], int id, uint8_t nib, const uint16_t s[4]) {
  w3o_slot_set_state(cell, id, 0, 0, s[0]);
  w3o_slot_set_state(cell, id, 1, nib >> 3, s[1]);
  w3o_slot_set_state(cell, id, 2, nib >> 2, s[2]);
  w3o_slot_set_state(cell, id, 3, nib >> 1, s[3]);
}
void w3o_slot_get_nib(const uint8_t cell[96], int id, uint8_t nib, uint16_t s[4]) {
  s[0] = w3o_slot_get_state(cell, id, 0, 0);
  s[1] = w3o_slot_get_state(cell, id, 1, nib >> 3);
  s[2] = w3o_slot_get_state(cell, id, 2, nib >> 2);
  s[3] = w3o_slot_get_state(cell, id, 3, nib >> 1);
}

/* OpinionMixer2::mix — src/mixers/opinion_mixer2.rs:5-10: the only mixer the reference has */
uint16_t w3o_opinion_mix2(uint16_t p1, uint16_t p2) {
  const uint16_t half = 1u << 15;
  uint16_t d1 = p1 >= half ? (uint16_t)(p1 - half) : (uint16_t)(half - p1);
  uint16_t d2 = p2 >= half ? (uint16_t)(p2 - half) : (uint16_t)(half - p2);
  return d1 >= d2 ? p1 : p2;
}
