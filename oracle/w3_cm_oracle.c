/*
 * w3_cm_oracle.c — CPU oracle of the hashed context-mixing stack "w3cm"
 * (SURVEY.md §8 rows a19 / f4).  TEST INFRASTRUCTURE ONLY (see w3_oracle.h).
 *
 * PARITY UNPINNED BY THE REFERENCE, by construction: the reference has no
 * logistic mixer and no APM ("APM mixers" is a goal in README.md:10), its hash
 * map and state table have no caller (src/lib.rs:11-13), and its miss path
 * never stores a check hash (src/hashmap.rs:64-68, "TODO").  This file is the
 * SPECIFICATION of the stack the north star describes, built from the
 * reference's own pieces where they exist:
 *
 *   cell / slot layout   src/hashmap.rs:32-129 (w3_tierb_oracle.c), cell index =
 *                        the high bits of a 64-bit hash (:25-28), check = low 12
 *   bit histories        NaiveStateTable (src/state_table/naive.rs): next, p
 *   coder + bit I/O      ArithmeticCoder / ACWriter / ACReader
 *                        (src/entropy_coding/, w3_oracle.c) — unchanged
 * and from this document where they do not:
 *
 *   models      M = 6 hashed contexts per nibble (docs/hashslots.md:4: one cell per nibble
 *               per context): the last 1..6 bytes; for the low nibble also the high nibble.
 *   hash        x = ctx * K1 + (i + 1) * K2 + (low nibble ? (0x100 | hi) * K3 : 0), then a
 *               64-bit finaliser (xor-shift 29, multiply, xor-shift 32).
 *   replacement on a miss the slot whose FIRST state has seen the fewest bits is
 *               cleared and takes the check hash (ties: lowest id).
 *   stretch / squash   12-bit logistic pair: squash interpolates the 33 values
 *               round(4096 * sigma((k - 16) / 2)) linearly; stretch is its inverse table.
 *   mixer       8 weight sets (bit position in the byte) of M + 1 weights (the last
 *               multiplies a constant 256), 16.16 fixed point, initial 0x4000 / 0:
 *               dot = sum((w * st) >> 8), pm = squash(dot >> 8);
 *               w += (st * err + 2^15) >> 16 with err = ((bit << 12) - pm) * 6, |w| <= 2^18.
 *   APM         256 contexts (the partial byte) x 33 bins of 16 bits, initial
 *               squash((j - 16) * 128) * 16; interpolation and update at rate 7;
 *               p = (pm + 3 * apm + 2) >> 2, clamped to [1, 4095]; coder prob = p << 4.
 *   update      every model's state follows StateTable::next; weights, APM bins.
 * Every block starts from zeroed cells, initial weights and bins (blocks are independent).
 */
#include <stdlib.h>
#include <string.h>

#include "w3_oracle.h"

#define CM_MODELS 6
#define K1 0x9E3779B97F4A7C15ull
#define K2 0xD6E8FEB86659FD93ull
#define K3 0xA24BAED4963EE407ull

/* ---- stretch / squash ------------------------------------------------------ */
static const int16_t SQ33[33] = {1,    2,    3,    6,    10,   16,   27,   45,   73,   120,  194,
                                 310,  488,  747,  1101, 1546, 2047, 2549, 2994, 3348, 3607, 3785,
                                 3901, 3975, 4024, 4050, 4068, 4079, 4085, 4089, 4092, 4093, 4094};
int w3o_cm_squash(int d) {
  if (d > 2047) return 4095;
  if (d < -2047) return 0;
  int w = d & 127;
  int k = (d >> 7) + 16; /* arithmetic shift: floor */
  return (SQ33[k] * (128 - w) + SQ33[k + 1] * w + 64) >> 7;
}
static int16_t g_stretch[4096];
static int g_stretch_ready = 0;
static void stretch_init(void) {
  if (g_stretch_ready) return;
  int pi = 0;
  for (int x = -2047; x <= 2047; x++) {
    int v = w3o_cm_squash(x);
    for (int i = pi; i <= v; i++) g_stretch[i] = (int16_t)x;
    pi = v + 1;
  }
  for (int i = pi; i < 4096; i++) g_stretch[i] = 2047;
  g_stretch_ready = 1;
}
int w3o_cm_stretch(int p) {
  stretch_init();
  return g_stretch[p & 4095];
}

/* bits the first state of a slot has seen: 0 for the empty state, 1 for the entry nodes'
 * successors, level + 1 inside the four sub-tables (src/state_table/naive.rs:17-75) */
int w3o_cm_state_bits(uint16_t s) {
  if (s == 0) return 0;
  if (s < 3) return 1;
  unsigned k = (s - 3u) % 990u, level = 1, filled = 0;
  while (filled + level <= k) { filled += level; level++; }
  return (int)level + 1;
}

uint64_t w3o_cm_hash(uint64_t last8, int model, int low_nibble, unsigned hi) {
  const int order = model + 1;
  const uint64_t ctx = last8 & ((order >= 8) ? ~0ull : (((uint64_t)1 << (8 * order)) - 1));
  uint64_t x = ctx * K1 + (uint64_t)(model + 1) * K2 + (low_nibble ? (uint64_t)(0x100u | hi) * K3 : 0);
  x ^= x >> 29;
  x *= 0xBF58476D1CE4E5B9ull;
  x ^= x >> 32;
  return x;
}

/* Cell::get_slot (src/hashmap.rs:43-71) with the replacement policy of this specification */
int w3o_cm_find_slot(uint8_t cell[96], uint64_t hash) {
  uint64_t hc = 0;
  for (int i = 0; i < 6; i++) hc = (hc << 8) | cell[i];
  const uint64_t chk = hash & 0xFFF;
  if (chk == (hc & 0xFFF)) return 3;
  if (chk == ((hc >> 12) & 0xFFF)) return 2;
  if (chk == ((hc >> 24) & 0xFFF)) return 1;
  if (chk == ((hc >> 36) & 0xFFF)) return 0;
  int victim = 0, best = 1 << 30;
  for (int id = 0; id < 4; id++) {
    int n = w3o_cm_state_bits(w3o_slot_get_state(cell, id, 0, 0));
    if (n < best) { best = n; victim = id; }
  }
  const int sh = 12 * (3 - victim);
  hc = (hc & ~((uint64_t)0xFFF << sh)) | (chk << sh);
  for (int i = 0; i < 6; i++) cell[i] = (uint8_t)(hc >> (8 * (5 - i)));
  for (int bit_id = 0; bit_id < 4; bit_id++)
    for (int c = 0; c < (1 << bit_id); c++) w3o_slot_set_state(cell, victim, bit_id, c, 0);
  return victim;
}

typedef struct {
  uint8_t *cells; /* 96 << log_cells */
  int log_cells;
  int32_t w[8][CM_MODELS + 1];
  uint16_t apm[256 * 33];
  uint64_t last8;
  uint32_t c0;
  /* per nibble */
  uint8_t *cell[CM_MODELS];
  int slot[CM_MODELS];
  /* per bit (filled by predict) */
  uint16_t st_state[CM_MODELS];
  int st[CM_MODELS], pm, apm_idx, apm_w;
} cm_model;

static cm_model *cm_new(int log_cells) {
  stretch_init();
  cm_model *m = (cm_model *)calloc(1, sizeof *m);
  if (!m) return NULL;
  m->cells = (uint8_t *)calloc((size_t)96 << log_cells, 1);
  if (!m->cells) { free(m); return NULL; }
  m->log_cells = log_cells;
  for (int s = 0; s < 8; s++) {
    for (int i = 0; i < CM_MODELS; i++) m->w[s][i] = 0x4000;
    m->w[s][CM_MODELS] = 0;
  }
  for (int c = 0; c < 256; c++)
    for (int j = 0; j < 33; j++) m->apm[c * 33 + j] = (uint16_t)(w3o_cm_squash((j - 16) * 128) * 16);
  m->c0 = 1;
  return m;
}
static void cm_free(cm_model *m) {
  if (m) free(m->cells);
  free(m);
}
static void cm_nibble_start(cm_model *m, int low, unsigned hi) {
  for (int i = 0; i < CM_MODELS; i++) {
    uint64_t h = w3o_cm_hash(m->last8, i, low, hi);
    m->cell[i] = m->cells + (size_t)96 * (size_t)(h >> (64 - m->log_cells));
    m->slot[i] = w3o_cm_find_slot(m->cell[i], h);
  }
}
/* 12-bit probability of a 1 for bit j of the current nibble (nib_ctx = its bits so far) */
static int cm_predict(cm_model *m, int bitpos, int j, unsigned nib_ctx) {
  int64_t dot = 0;
  const int32_t *w = m->w[bitpos];
  for (int i = 0; i < CM_MODELS; i++) {
    m->st_state[i] = w3o_slot_get_state(m->cell[i], m->slot[i], j, (int)nib_ctx);
    m->st[i] = w3o_cm_stretch(w3o_state_p(m->st_state[i]) >> 4);
    dot += ((int64_t)w[i] * m->st[i]) >> 8;
  }
  dot += ((int64_t)w[CM_MODELS] * 256) >> 8;
  int d = (int)(dot >> 8);
  m->pm = w3o_cm_squash(d);
  int s = w3o_cm_stretch(m->pm) + 2048;
  m->apm_w = s & 127;
  m->apm_idx = (int)m->c0 * 33 + (s >> 7);
  int pa = (m->apm[m->apm_idx] * (128 - m->apm_w) + m->apm[m->apm_idx + 1] * m->apm_w) >> 11;
  int p = (m->pm + 3 * pa + 2) >> 2;
  return p < 1 ? 1 : p > 4095 ? 4095 : p;
}
static void cm_update(cm_model *m, int bitpos, int j, unsigned nib_ctx, unsigned bit) {
  const int err = (((int)bit << 12) - m->pm) * 6;
  int32_t *w = m->w[bitpos];
  for (int i = 0; i <= CM_MODELS; i++) {
    const int st = i < CM_MODELS ? m->st[i] : 256;
    int32_t nw = w[i] + (int32_t)(((int64_t)st * err + (1 << 15)) >> 16);
    w[i] = nw > (1 << 18) ? (1 << 18) : nw < -(1 << 18) ? -(1 << 18) : nw;
  }
  const int g = (int)((bit << 16) + (bit << 7) - bit - bit);
  m->apm[m->apm_idx] = (uint16_t)(m->apm[m->apm_idx] + ((g - (int)m->apm[m->apm_idx]) >> 7));
  m->apm[m->apm_idx + 1] = (uint16_t)(m->apm[m->apm_idx + 1] + ((g - (int)m->apm[m->apm_idx + 1]) >> 7));
  for (int i = 0; i < CM_MODELS; i++)
    w3o_slot_set_state(m->cell[i], m->slot[i], j, (int)nib_ctx, w3o_state_next(m->st_state[i], (int)bit));
  m->c0 = (m->c0 << 1) | bit;
}

/* the byte writer / reader and the coder live in w3_oracle.c */
int w3o_ac_encode_probs(const uint8_t *in, size_t n_in, const uint16_t *probs, size_t n_probs, uint8_t *out,
                        size_t cap, size_t *out_len);

/* compress one block with the w3cm stack: probabilities first, then the reference coder over them
 * (the coder consumes (bit, p) pairs exactly as in src/main.rs:105-107) */
int w3o_cm_compress_block(const uint8_t *in, size_t n, int log_cells, uint8_t *out, size_t cap, size_t *out_len) {
  if ((!in && n) || !out_len || log_cells < 4 || log_cells > 24) return W3O_E_ARG;
  cm_model *m = cm_new(log_cells);
  uint16_t *probs = (uint16_t *)malloc((n * 8 + 1) * sizeof(uint16_t));
  if (!m || !probs) { cm_free(m); free(probs); return W3O_E_ARG; }
  size_t t = 0;
  for (size_t i = 0; i < n; i++) {
    const unsigned byte = in[i];
    for (int nb = 0; nb < 2; nb++) {
      const unsigned nib = nb ? (byte & 15u) : (byte >> 4);
      cm_nibble_start(m, nb, byte >> 4);
      unsigned nib_ctx = 0;
      for (int j = 0; j < 4; j++) {
        const unsigned bit = (nib >> (3 - j)) & 1u;
        probs[t++] = (uint16_t)(cm_predict(m, nb * 4 + j, j, nib_ctx) << 4);
        cm_update(m, nb * 4 + j, j, nib_ctx, bit);
        nib_ctx = (nib_ctx << 1) | bit;
      }
    }
    m->last8 = (m->last8 << 8) | byte;
    m->c0 = 1;
  }
  int rc = w3o_ac_encode_probs(in, n, probs, t, out, cap, out_len);
  cm_free(m);
  free(probs);
  return rc;
}

/* the decoder needs the coder's state machine bit by bit: restated here over the public reader */
int w3o_cm_decompress_block(const uint8_t *in, size_t n_in, int log_cells, uint8_t *out, size_t n_out) {
  if ((!in && n_in) || (!out && n_out) || log_cells < 4 || log_cells > 24) return W3O_E_ARG;
  cm_model *m = cm_new(log_cells);
  w3o_reader *r = w3o_reader_new(in, n_in);
  if (!m || !r) { cm_free(m); w3o_reader_free(r); return W3O_E_ARG; }
  uint32_t x1 = 0, x2 = 0xFFFFFFFFu, x = w3o_reader_read_u32(r); /* new_decoder — arithmetic_coder.rs:75-78 */
  for (size_t i = 0; i < n_out; i++) {
    unsigned byte = 0;
    for (int nb = 0; nb < 2; nb++) {
      cm_nibble_start(m, nb, byte & 15u); /* for the low nibble `byte` holds the high nibble so far */
      unsigned nib_ctx = 0;
      for (int j = 0; j < 4; j++) {
        const uint16_t prob = (uint16_t)(cm_predict(m, nb * 4 + j, j, nib_ctx) << 4);
        /* ArithmeticCoder::decode — arithmetic_coder.rs:80-105 */
        const uint64_t P = prob == 0 ? 1 : ((uint64_t)prob << 16);
        const uint32_t xmid = x1 + (uint32_t)(((uint64_t)(x2 - x1) * P) >> 32);
        const unsigned bit = x <= xmid;
        if (bit) x2 = xmid; else x1 = xmid + 1;
        while (((x1 ^ x2) >> 31) == 0) {
          x1 <<= 1; x2 = (x2 << 1) | 1; x = (x << 1) | (uint32_t)w3o_reader_read_bit(r);
        }
        while (x1 >= 0x40000000u && x2 < 0xC0000000u) {
          x1 = (x1 << 1) & 0x7FFFFFFFu; x2 = (x2 << 1) | 0x80000001u;
          x = ((x << 1) ^ 0x80000000u) | (uint32_t)w3o_reader_read_bit(r);
        }
        cm_update(m, nb * 4 + j, j, nib_ctx, bit);
        nib_ctx = (nib_ctx << 1) | bit;
        byte = (byte << 1) | bit;
      }
    }
    out[i] = (uint8_t)byte;
    m->last8 = (m->last8 << 8) | byte;
    m->c0 = 1;
  }
  w3o_reader_free(r);
  cm_free(m);
  return W3O_OK;
}
