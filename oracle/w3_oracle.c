/*
 * w3_oracle.c — CPU restatement of weath3rb0i's predict→update→code path.
 * TEST INFRASTRUCTURE ONLY (see w3_oracle.h).  Plain C11 + pthreads.
 *
 * The structure deliberately follows the reference's call stack
 *  rather than anything the GPU path does: one generic coder
 * step, two sinks (byte writer, 8-bit entropy sink), a model object with
 * predict/adapt/update.  All citations are relative to /root/reference/.
 */
#include "w3_oracle.h"

#include <pthread.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------------ */
/* Counter — src/models/counter.rs                                          */

uint16_t w3o_counter_p(uint16_t c0, uint16_t c1) {
  /* counter.rs:14-17: p = (1<<17)*(c1+1)/(c0+c1+2) in u64, then round */
  uint64_t n0 = c0, n1 = c1;
  uint64_t p = ((uint64_t)1 << 17) * (n1 + 1) / (n0 + n1 + 2);
  return (uint16_t)((p >> 1) + (p & 1));
}

void w3o_counter_update(uint16_t c[2], int bit) {
  /* counter.rs:21-25 */
  c[bit] = (uint16_t)(c[bit] + 1);
  if (c[bit] == UINT16_MAX) {
    c[0] = (uint16_t)((c[0] >> 1) + (c[0] & 1));
    c[1] = (uint16_t)((c[1] >> 1) + (c[1] & 1));
  }
}

/* ------------------------------------------------------------------------ */
/* Arithmetic coder constants — src/entropy_coding/arithmetic_coder.rs:4-9  */

#define PREC_SHIFT 31u
#define Q1 0x40000000u
#define Q2 0x80000000u
#define Q3 0xC0000000u
#define RLO_MOD 0x7FFFFFFFu
#define RHI_MOD 0x80000001u

/* lerp — arithmetic_coder.rs:109-119 */
static inline uint32_t lerp(uint32_t x1, uint32_t x2, uint16_t prob) {
  uint64_t p = prob == 0 ? 1 : ((uint64_t)prob << 16);
  uint64_t range = (uint64_t)(x2 - x1);
  uint64_t lerped = (range * p) >> 32;
  return x1 + (uint32_t)lerped;
}

/* ------------------------------------------------------------------------ */
/* ACWriter — src/entropy_coding/io.rs:52-101                               */

struct w3o_writer {
  uint8_t *out;
  size_t cap;
  size_t len; /* bytes that the reference would have written */
  uint8_t buf;
  uint8_t idx;
  uint64_t rev_bits;
};

static inline void writer_raw(w3o_writer *w, unsigned bit) {
  /* io.rs:74-81 */
  w->buf = (uint8_t)((w->buf << 1) | bit);
  w->idx = (uint8_t)((w->idx + 1) % 8);
  if (w->idx == 0) {
    if (w->len < w->cap) w->out[w->len] = w->buf;
    w->len++;
  }
}

static inline void writer_write_bit(w3o_writer *w, unsigned bit) {
  /* io.rs:83-88 */
  writer_raw(w, bit);
  while (w->rev_bits > 0) {
    w->rev_bits--;
    writer_raw(w, bit ^ 1u);
  }
}

static inline void writer_flush(w3o_writer *w, uint32_t state) {
  /* io.rs:91-100 */
  writer_write_bit(w, state >> 31);
  state <<= 1;
  while (w->idx > 0) {
    writer_write_bit(w, state >> 31);
    state <<= 1;
  }
}

static void writer_init(w3o_writer *w, uint8_t *buf, size_t cap) {
  w->out = buf;
  w->cap = cap;
  w->len = 0;
  w->buf = 0;
  w->idx = 0;
  w->rev_bits = 0;
}

w3o_writer *w3o_writer_new(uint8_t *buf, size_t cap) {
  w3o_writer *w = (w3o_writer *)malloc(sizeof *w);
  if (w) writer_init(w, buf, cap);
  return w;
}
void w3o_writer_free(w3o_writer *w) { free(w); }
void w3o_writer_write_bit(w3o_writer *w, int bit) { writer_write_bit(w, bit & 1); }
void w3o_writer_inc_parity(w3o_writer *w) { w->rev_bits++; }
void w3o_writer_flush(w3o_writer *w, uint32_t state) { writer_flush(w, state); }
size_t w3o_writer_len(const w3o_writer *w) { return w->len; }
int w3o_writer_overflowed(const w3o_writer *w) { return w->len > w->cap; }

/* ------------------------------------------------------------------------ */
/* ACReader — src/entropy_coding/io.rs:7-49                                 */

struct w3o_reader {
  const uint8_t *in;
  size_t n;
  size_t at;
  uint8_t buf;
  uint8_t mask;
};

static inline uint8_t reader_byte(w3o_reader *r) {
  /* io.rs:18-27: UnexpectedEof → 0 */
  if (r->at < r->n) return r->in[r->at++];
  return 0;
}

static inline unsigned reader_bit(w3o_reader *r) {
  /* io.rs:31-38 */
  r->mask >>= 1;
  if (r->mask == 0) {
    r->buf = reader_byte(r);
    r->mask = 1u << 7;
  }
  return (r->buf & r->mask) ? 1u : 0u;
}

static inline uint32_t reader_u32(w3o_reader *r) {
  /* io.rs:40-48 */
  uint32_t b0 = reader_byte(r), b1 = reader_byte(r), b2 = reader_byte(r),
           b3 = reader_byte(r);
  return (b0 << 24) | (b1 << 16) | (b2 << 8) | b3;
}

static void reader_init(w3o_reader *r, const uint8_t *buf, size_t n) {
  r->in = buf;
  r->n = n;
  r->at = 0;
  r->buf = 0;
  r->mask = 0;
}

w3o_reader *w3o_reader_new(const uint8_t *buf, size_t n) {
  w3o_reader *r = (w3o_reader *)malloc(sizeof *r);
  if (r) reader_init(r, buf, n);
  return r;
}
void w3o_reader_free(w3o_reader *r) { free(r); }
int w3o_reader_read_bit(w3o_reader *r) { return (int)reader_bit(r); }
uint32_t w3o_reader_read_u32(w3o_reader *r) { return reader_u32(r); }

/* ------------------------------------------------------------------------ */
/* ArithmeticCoder<ACWriter>::encode / flush — arithmetic_coder.rs:41-71    */

typedef struct {
  uint32_t x1, x2, x;
} coder;

static inline void coder_init(coder *c) {
  /* new_coder — arithmetic_coder.rs:37-39 */
  c->x1 = 0;
  c->x2 = UINT32_MAX;
  c->x = 0;
}

static inline void ac_encode(coder *c, unsigned bit, uint16_t prob,
                             w3o_writer *io) {
  uint32_t xmid = lerp(c->x1, c->x2, prob); /* :42 */
  if (bit == 0) c->x1 = xmid + 1;           /* :45-48 */
  else c->x2 = xmid;
  while (((c->x1 ^ c->x2) >> PREC_SHIFT) == 0) { /* :51-55 */
    writer_write_bit(io, c->x1 >> PREC_SHIFT);
    c->x1 <<= 1;
    c->x2 = (c->x2 << 1) | 1;
  }
  while (c->x1 >= Q1 && c->x2 < Q3) { /* :58-62 */
    io->rev_bits++;
    c->x1 = (c->x1 << 1) & RLO_MOD;
    c->x2 = (c->x2 << 1) | RHI_MOD;
  }
}

/* ArithmeticCoder<ACReader>::decode — arithmetic_coder.rs:80-105 */
static inline unsigned ac_decode(coder *c, uint16_t prob, w3o_reader *io) {
  uint32_t xmid = lerp(c->x1, c->x2, prob); /* :81 */
  unsigned bit = c->x <= xmid;              /* :82 */
  if (bit == 0) c->x1 = xmid + 1;           /* :85-88 */
  else c->x2 = xmid;
  while (((c->x1 ^ c->x2) >> PREC_SHIFT) == 0) { /* :91-95 */
    c->x1 <<= 1;
    c->x2 = (c->x2 << 1) | 1;
    c->x = (c->x << 1) | reader_bit(io);
  }
  while (c->x1 >= Q1 && c->x2 < Q3) { /* :98-102 */
    c->x1 = (c->x1 << 1) & RLO_MOD;
    c->x2 = (c->x2 << 1) | RHI_MOD;
    c->x = ((c->x << 1) ^ Q2) | reader_bit(io);
  }
  return bit;
}

/* ------------------------------------------------------------------------ */
/* EntropyWriter + ACHistory::hash — src/history/ac_history.rs:28-88        */

const uint16_t W3O_BOOK1_TABLE[8] = {1,     50188, 62497, 15819,
                                     22545, 31499, 22988, 29616};

typedef struct {
  uint32_t state;
  uint16_t rev_bits;
  uint8_t idx;
  uint8_t max_bits;
} entropy_sink;

/* write_bit_raw closure — ac_history.rs:63-71; returns 1 on Err */
static inline int sink_raw(entropy_sink *s, unsigned bit) {
  if (s->idx == s->max_bits) return 1;
  s->state = (s->state >> 1) | ((uint32_t)bit << 31);
  s->idx++;
  return 0;
}

/* ACWrite::write_bit for EntropyWriter — ac_history.rs:58-80 */
static inline int sink_write_bit(entropy_sink *s, unsigned bit) {
  if (sink_raw(s, bit)) return 1;
  while (s->rev_bits > 0) {
    s->rev_bits--;
    if (sink_raw(s, bit ^ 1u)) return 1;
  }
  return 0;
}

uint32_t w3o_achash_ex(uint64_t bits, uint64_t pos, const uint16_t table[8],
                       int max_bits, int *steps) {
  coder ac;
  coder_init(&ac); /* ac_history.rs:29 */
  entropy_sink w = {0, 0, 0, (uint8_t)max_bits}; /* :30-35 */
  uint8_t alignment = (uint8_t)(pos & 7);        /* :36, stationary.rs:50-52 */
  int i, full_at = 64; /* full_at: history bits consumed when idx hit max_bits */
  for (i = 0; i < 64; i++) { /* :37 */
    unsigned bit = (unsigned)((bits >> i) & 1); /* :38 */
    alignment = (uint8_t)((alignment + 7) & 7); /* stationary.rs:55 */
    uint16_t prob = table[alignment];           /* stationary.rs:56 */
    /* ac.encode(bit, prob, &mut writer) with `?` on every write_bit —
     * arithmetic_coder.rs:41-65 */
    uint32_t xmid = lerp(ac.x1, ac.x2, prob);
    int err = 0;
    if (bit == 0) ac.x1 = xmid + 1;
    else ac.x2 = xmid;
    while (((ac.x1 ^ ac.x2) >> PREC_SHIFT) == 0) {
      if (sink_write_bit(&w, ac.x1 >> PREC_SHIFT)) {
        err = 1;
        break;
      }
      ac.x1 <<= 1;
      ac.x2 = (ac.x2 << 1) | 1;
    }
    if (w.idx == w.max_bits && full_at == 64) full_at = i + 1;
    if (err) break; /* ac_history.rs:40-42 */
    while (ac.x1 >= Q1 && ac.x2 < Q3) {
      w.rev_bits++; /* inc_parity — ac_history.rs:82-84 */
      ac.x1 = (ac.x1 << 1) & RLO_MOD;
      ac.x2 = (ac.x2 << 1) | RHI_MOD;
    }
  }
  /* the sink can no longer change once idx == max_bits, so the hash is a
   * function of the first `full_at` history bits only (65 = never filled) */
  if (steps) *steps = (w.idx == w.max_bits) ? full_at : 65;
  /* ac_history.rs:45: state >> (32 - idx).  idx == 0 would be a shift by 32
   * (wraps to 0 in release Rust; state is 0 then) — define as 0. */
  if (w.idx == 0) return 0;
  return w.state >> (32 - w.idx);
}

uint32_t w3o_achash(uint64_t bits, uint64_t pos, int *steps) {
  /* main.rs:151: ACHistory::new(8, StationaryModel::for_book1()) */
  return w3o_achash_ex(bits, pos, W3O_BOOK1_TABLE, 8, steps);
}

/* ------------------------------------------------------------------------ */
/* OrderNEntropy<ACHistory<StationaryModel>> — src/models/ordern_entropy.rs */

#define BITS_IN_CONTEXT 11 /* main.rs:151 */
#define ALIGNMENT_BITS 3   /* main.rs:151 */
#define HASH_MAX_BITS 8    /* main.rs:151: ACHistory::new(8, ...) */

typedef struct {
  uint16_t (*stats)[2]; /* Vec<Counter>, 1 << bits_in_context entries — ordern_entropy.rs:17 */
  uint32_t ctx;
  uint8_t alignment;
  uint8_t bits_in_context; /* ordern_entropy.rs:10 */
  uint8_t alignment_bits;  /* :11 */
  uint8_t max_bits;   /* ACHistory.max_bits — ac_history.rs:12, a parameter of its own (:16-18) */
  uint64_t hist_pos;  /* ACHistory.pos  ac_history.rs:10 */
  uint64_t hist_bits; /* ACHistory.bits ac_history.rs:11 */
  const uint16_t *table; /* StationaryModel.table  stationary.rs:9 */
} model;

/* OrderNEntropy::new(bits_in_context, alignment_bits, ACHistory::new(max_bits, model)) —
 * ordern_entropy.rs:15-24, ac_history.rs:16-18 */
static model *model_new(const uint16_t *table, int bits_in_context, int alignment_bits, int max_bits) {
  model *m = (model *)calloc(1, sizeof *m);
  if (!m) return NULL;
  m->stats = (uint16_t(*)[2])calloc((size_t)1 << bits_in_context, sizeof(uint16_t[2]));
  if (!m->stats) {
    free(m);
    return NULL;
  }
  m->bits_in_context = (uint8_t)bits_in_context;
  m->alignment_bits = (uint8_t)alignment_bits;
  m->max_bits = (uint8_t)max_bits;
  m->table = table ? table : W3O_BOOK1_TABLE; /* main.rs:151 */
  return m;
}
static void model_free(model *m) {
  if (m) free(m->stats);
  free(m);
}
/* init_model — main.rs:146-152 */
static model *model_init(const uint16_t *table) {
  return model_new(table, BITS_IN_CONTEXT, ALIGNMENT_BITS, HASH_MAX_BITS);
}

static inline uint16_t model_predict(const model *m) {
  /* ordern_entropy.rs:28-30 */
  return w3o_counter_p(m->stats[m->ctx][0], m->stats[m->ctx][1]);
}

static inline void model_update(model *m, unsigned bit) {
  /* blanket impl models/mod.rs:28-31: adapt, then update */
  w3o_counter_update(m->stats[m->ctx], (int)bit); /* ordern_entropy.rs:32-34 */
  /* ordern_entropy.rs:36-45 */
  const uint32_t mask_bits = (uint32_t)m->bits_in_context - m->alignment_bits;
  const uint32_t mask = (uint32_t)(((uint64_t)1 << mask_bits) - 1);
  const uint8_t alignment_mask = (uint8_t)((1u << m->alignment_bits) - 1);
  m->hist_bits = (m->hist_bits << 1) | bit; /* ac_history.rs:24 */
  m->hist_pos += 1;                         /* ac_history.rs:25 */
  m->alignment = (uint8_t)((m->alignment + 1) & alignment_mask);
  uint32_t hash = w3o_achash_ex(m->hist_bits, m->hist_pos, m->table, (int)m->max_bits, NULL) & mask;
  m->ctx = (hash << m->alignment_bits) | m->alignment;
}

/* ------------------------------------------------------------------------ */
/* OrderN (raw bit history) — src/models/ordern.rs:4-46; compress() of the     */
/* ordern sweep binary — src/bin/ordern/main.rs:66-81                          */

/* bits ACStats counts for OrderN::new(ctx_bits, alignment_bits) over buf (csize printed = / 8) */
int w3o_sweep_exec_ordern(const uint8_t *buf, size_t n, int ctx_bits, int alignment_bits, uint64_t *bit_count);

/* ------------------------------------------------------------------------ */
/* compress / decompress — src/main.rs:89-144                               */

int w3o_compress_block(const uint8_t *in, size_t n, uint8_t *out, size_t cap,
                       size_t *out_len) {
  return w3o_compress_block_ex(in, n, NULL, out, cap, out_len);
}

/* StationaryModel::new — src/models/ac_hash/stationary.rs:14-34 */
void w3o_stationary_new(const uint8_t *buf, size_t n, uint16_t table[8]) {
  uint16_t model[8][2];
  memset(model, 0, sizeof model);
  for (size_t b = 0; b < n; b++) {
    int i = 7;                       /* :17 */
    for (int k = 7; k >= 0; k--) {   /* unroll_for!: MSB first, macros.rs:2-21 */
      i = (i + 1) & 7;               /* :19 */
      w3o_counter_update(model[i], (buf[b] >> k) & 1); /* :20 */
    }
  }
  for (int i = 0; i < 8; i++) table[i] = w3o_counter_p(model[i][0], model[i][1]); /* :23-32 */
}

int w3o_compress_block_ex(const uint8_t *in, size_t n, const uint16_t *table, uint8_t *out,
                          size_t cap, size_t *out_len) {
  return w3o_compress_block_geo(in, n, table, BITS_IN_CONTEXT, ALIGNMENT_BITS, HASH_MAX_BITS, out, cap, out_len);
}

static int geo_ok(int bits_in_context, int alignment_bits, int max_bits) {
  return bits_in_context >= 1 && bits_in_context <= 30 && alignment_bits >= 0 && alignment_bits <= 7 &&
         alignment_bits <= bits_in_context && max_bits >= 1 && max_bits <= 32;
}

/* compress loop (main.rs:99-111) for any OrderNEntropy::new(bits_in_context, alignment_bits,
 * ACHistory::new(max_bits, StationaryModel::from_table(table))) */
int w3o_compress_block_geo(const uint8_t *in, size_t n, const uint16_t *table, int bits_in_context,
                           int alignment_bits, int max_bits, uint8_t *out, size_t cap, size_t *out_len) {
  if ((!in && n) || (!out && cap) || !out_len || !geo_ok(bits_in_context, alignment_bits, max_bits)) return W3O_E_ARG;
  model *m = model_new(table, bits_in_context, alignment_bits, max_bits);
  if (!m) return W3O_E_ARG;
  w3o_writer w;
  writer_init(&w, out, cap);
  coder ac;
  coder_init(&ac);
  for (size_t i = 0; i < n; i++) { /* main.rs:103 */
    uint8_t byte = in[i];
    for (int k = 7; k >= 0; k--) { /* main.rs:104 */
      unsigned bit = (byte >> k) & 1;
      uint16_t p = model_predict(m); /* :105 */
      model_update(m, bit);          /* :106 */
      ac_encode(&ac, bit, p, &w);    /* :107 */
    }
  }
  writer_flush(&w, ac.x2); /* main.rs:111 → arithmetic_coder.rs:67-71 */
  model_free(m);
  *out_len = w.len;
  return w.len > cap ? W3O_E_CAP : W3O_OK;
}

int w3o_decompress_block(const uint8_t *in, size_t n_in, uint8_t *out,
                         size_t n_out) {
  return w3o_decompress_block_ex(in, n_in, NULL, out, n_out);
}

int w3o_decompress_block_ex(const uint8_t *in, size_t n_in, const uint16_t *table, uint8_t *out,
                            size_t n_out) {
  return w3o_decompress_block_geo(in, n_in, table, BITS_IN_CONTEXT, ALIGNMENT_BITS, HASH_MAX_BITS, out, n_out);
}

int w3o_decompress_block_geo(const uint8_t *in, size_t n_in, const uint16_t *table, int bits_in_context,
                             int alignment_bits, int max_bits, uint8_t *out, size_t n_out) {
  if ((!in && n_in) || (!out && n_out) || !geo_ok(bits_in_context, alignment_bits, max_bits)) return W3O_E_ARG;
  model *m = model_new(table, bits_in_context, alignment_bits, max_bits);
  if (!m) return W3O_E_ARG;
  w3o_reader r;
  reader_init(&r, in, n_in);
  coder ac;
  coder_init(&ac);
  ac.x = reader_u32(&r); /* new_decoder — arithmetic_coder.rs:75-78 */
  for (size_t i = 0; i < n_out; i++) { /* main.rs:131 */
    unsigned byte = 0;
    for (int k = 0; k < 8; k++) {           /* :133 */
      uint16_t p = model_predict(m);        /* :134 */
      unsigned bit = ac_decode(&ac, p, &r); /* :135 */
      model_update(m, bit);                 /* :136 */
      byte = (byte << 1) | bit;             /* :137 */
    }
    out[i] = (uint8_t)byte; /* :139 */
  }
  model_free(m);
  return W3O_OK;
}

/* ------------------------------------------------------------------------ */
/* ACStats — src/helpers.rs:60-90: the bit-counting ACWrite sink of the sweep */
/* binaries; exec() — src/bin/entropy-hashing-ac/main.rs:48-75                */

typedef struct {
  uint64_t bit_count, rev_bits;
} acstats;

/* ac.encode(bit, p, &mut ACStats) — arithmetic_coder.rs:41-65 with helpers.rs:76-89 as the sink */
static inline void ac_encode_stats(coder *c, unsigned bit, uint16_t prob, acstats *io) {
  uint32_t xmid = lerp(c->x1, c->x2, prob);
  if (bit == 0) c->x1 = xmid + 1;
  else c->x2 = xmid;
  while (((c->x1 ^ c->x2) >> PREC_SHIFT) == 0) {
    io->bit_count += 1 + io->rev_bits; /* write_bit — helpers.rs:81-85 */
    io->rev_bits = 0;
    c->x1 <<= 1;
    c->x2 = (c->x2 << 1) | 1;
  }
  while (c->x1 >= Q1 && c->x2 < Q3) {
    io->rev_bits += 1; /* inc_parity — helpers.rs:77-79 */
    c->x1 = (c->x1 << 1) & RLO_MOD;
    c->x2 = (c->x2 << 1) | RHI_MOD;
  }
}

int w3o_sweep_exec_ordern(const uint8_t *buf, size_t n, int ctx_bits, int alignment_bits, uint64_t *bit_count) {
  if ((!buf && n) || !bit_count || ctx_bits < 1 || ctx_bits > 30 || alignment_bits < 0 || alignment_bits > 7 ||
      alignment_bits > ctx_bits)
    return W3O_E_ARG;
  uint16_t(*stats)[2] = (uint16_t(*)[2])calloc((size_t)1 << ctx_bits, sizeof(uint16_t[2])); /* ordern.rs:16 */
  if (!stats) return W3O_E_ARG;
  uint32_t ctx = 0, history = 0;
  uint8_t alignment = 0;
  const uint32_t mask_bits = (uint32_t)(ctx_bits - alignment_bits);
  const uint32_t mask = (uint32_t)(((uint64_t)1 << mask_bits) - 1);
  const uint8_t alignment_mask = (uint8_t)((1u << alignment_bits) - 1);
  coder ac;
  coder_init(&ac);
  acstats w = {0, 0};
  for (size_t i = 0; i < n; i++)
    for (int k = 7; k >= 0; k--) {
      unsigned bit = (buf[i] >> k) & 1;
      uint16_t p = w3o_counter_p(stats[ctx][0], stats[ctx][1]); /* predict — ordern.rs:28-30 */
      w3o_counter_update(stats[ctx], (int)bit);                  /* adapt — :32-34 */
      history = ((history << 1) | bit) & mask;                   /* update — :36-45 */
      alignment = (uint8_t)((alignment + 1) & alignment_mask);
      ctx = (history << alignment_bits) | alignment;
      ac_encode_stats(&ac, bit, p, &w);
    }
  free(stats);
  *bit_count = w.bit_count;
  return W3O_OK;
}

/* exec(buf, ctx_bits, alignment_bits, ACHistory::new(ctx_bits - alignment_bits, model)) of the
 * entropy-hashing-ac sweep (src/bin/entropy-hashing-ac/main.rs:24-25, 48-75): returns the number of
 * bits ACStats counted; csize as printed is that / 8 (helpers.rs:71-73); ac.flush adds nothing
 * (ACStats::flush is a no-op, helpers.rs:87-89). */
int w3o_sweep_exec(const uint8_t *buf, size_t n, const uint16_t *table, int ctx_bits, int alignment_bits,
                   uint64_t *bit_count) {
  const int max_bits = ctx_bits - alignment_bits; /* main.rs:24 of that bin */
  if ((!buf && n) || !bit_count || !geo_ok(ctx_bits, alignment_bits, max_bits)) return W3O_E_ARG;
  model *m = model_new(table, ctx_bits, alignment_bits, max_bits);
  if (!m) return W3O_E_ARG;
  coder ac;
  coder_init(&ac);
  acstats w = {0, 0};
  for (size_t i = 0; i < n; i++)
    for (int k = 7; k >= 0; k--) { /* unroll_for!(bit in byte, …) — macros.rs:2-21 */
      unsigned bit = (buf[i] >> k) & 1;
      uint16_t p = model_predict(m);   /* :56 */
      model_update(m, bit);            /* :57 */
      ac_encode_stats(&ac, bit, p, &w); /* :58 */
    }
  model_free(m);
  *bit_count = w.bit_count;
  return W3O_OK;
}

void w3o_ctx_trace(const uint8_t *in, size_t n, uint16_t *ctx_out) {
  model *m = model_init(NULL);
  if (!m) return;
  size_t t = 0;
  for (size_t i = 0; i < n; i++)
    for (int k = 7; k >= 0; k--) {
      ctx_out[t++] = (uint16_t)m->ctx;
      model_update(m, (in[i] >> k) & 1);
    }
  model_free(m);
}

/* header — src/main.rs:14-15, 95-96, 119-126 */
static const uint8_t MAGIC_STR[4] = {'w', '3', '0', 'i'};

int w3o_compress_file(const uint8_t *in, size_t n, uint8_t *out, size_t cap,
                      size_t *out_len) {
  if (!out_len) return W3O_E_ARG;
  if (cap < 12) {
    size_t body = 0;
    w3o_compress_block(in, n, NULL, 0, &body);
    *out_len = 12 + body;
    return W3O_E_CAP;
  }
  memcpy(out, MAGIC_STR, 4);
  uint64_t len = (uint64_t)n;
  for (int i = 0; i < 8; i++) out[4 + i] = (uint8_t)(len >> (56 - 8 * i));
  size_t body = 0;
  int rc = w3o_compress_block(in, n, out + 12, cap - 12, &body);
  *out_len = 12 + body;
  return rc;
}

int w3o_decompress_file(const uint8_t *in, size_t n_in, uint8_t *out,
                        size_t cap, size_t *out_len) {
  if (!in || n_in < 12 || !out_len) return W3O_E_ARG; /* read_exact main.rs:121 */
  if (memcmp(in, MAGIC_STR, 4) != 0) return W3O_E_MAGIC;
  uint64_t len = 0;
  for (int i = 0; i < 8; i++) len = (len << 8) | in[4 + i];
  *out_len = (size_t)len;
  if (len > cap) return W3O_E_CAP;
  return w3o_decompress_block(in + 12, n_in - 12, out, (size_t)len);
}

/* ------------------------------------------------------------------------ */
/* coder driven by caller-supplied probabilities —                          */
/* src/entropy_coding/arithmetic_coder_tests.rs:8-42                        */

int w3o_ac_encode_probs(const uint8_t *in, size_t n_in, const uint16_t *probs,
                        size_t n_probs, uint8_t *out, size_t cap,
                        size_t *out_len) {
  if (!out_len) return W3O_E_ARG;
  w3o_reader r;
  reader_init(&r, in, n_in);
  w3o_writer w;
  writer_init(&w, out, cap);
  coder ac;
  coder_init(&ac);
  for (size_t i = 0; i < n_probs; i++) {
    unsigned bit = reader_bit(&r);       /* tests.rs:15 */
    ac_encode(&ac, bit, probs[i], &w);   /* tests.rs:16 */
  }
  writer_flush(&w, ac.x2); /* tests.rs:19 */
  *out_len = w.len;
  return w.len > cap ? W3O_E_CAP : W3O_OK;
}

int w3o_ac_decode_probs(const uint8_t *in, size_t n_in, const uint16_t *probs,
                        size_t n_probs, uint8_t *out, size_t cap) {
  if (cap * 8 < n_probs) return W3O_E_CAP;
  w3o_reader r;
  reader_init(&r, in, n_in);
  coder ac;
  coder_init(&ac);
  ac.x = reader_u32(&r); /* tests.rs:27 */
  memset(out, 0, (n_probs + 7) / 8);
  for (size_t i = 0; i < n_probs; i++) {
    unsigned bit = ac_decode(&ac, probs[i], &r); /* tests.rs:32 */
    out[i >> 3] |= (uint8_t)(bit << (7 - (i & 7)));
  }
  return W3O_OK;
}

/* ------------------------------------------------------------------------ */
/* multi-threaded CPU baseline: one block per thread at a time              */

typedef struct {
  int decode;
  const uint8_t *in;
  uint8_t *out;
  size_t n, block_size, stride;
  uint32_t *lens;
  const uint32_t *in_lens;
  size_t n_blocks;
  size_t *next;
  pthread_mutex_t *mu;
  int rc;
} mt_job;

static void *mt_worker(void *arg) {
  mt_job *j = (mt_job *)arg;
  for (;;) {
    pthread_mutex_lock(j->mu);
    size_t b = (*j->next)++;
    pthread_mutex_unlock(j->mu);
    if (b >= j->n_blocks) break;
    size_t off = b * j->block_size;
    size_t len = j->n - off < j->block_size ? j->n - off : j->block_size;
    if (!j->decode) {
      size_t ol = 0;
      int rc = w3o_compress_block(j->in + off, len, j->out + b * j->stride,
                                  j->stride, &ol);
      j->lens[b] = (uint32_t)ol;
      if (rc != W3O_OK) j->rc = rc;
    } else {
      int rc = w3o_decompress_block(j->in + b * j->stride, j->in_lens[b],
                                    j->out + off, len);
      if (rc != W3O_OK) j->rc = rc;
    }
  }
  return NULL;
}

static int run_mt(mt_job proto, int n_threads) {
  if (n_threads < 1) n_threads = 1;
  if ((size_t)n_threads > proto.n_blocks && proto.n_blocks > 0)
    n_threads = (int)proto.n_blocks;
  pthread_mutex_t mu = PTHREAD_MUTEX_INITIALIZER;
  size_t next = 0;
  pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t) * (size_t)n_threads);
  mt_job *jobs = (mt_job *)malloc(sizeof(mt_job) * (size_t)n_threads);
  if (!th || !jobs) {
    free(th);
    free(jobs);
    return W3O_E_ARG;
  }
  for (int i = 0; i < n_threads; i++) {
    jobs[i] = proto;
    jobs[i].next = &next;
    jobs[i].mu = &mu;
    jobs[i].rc = W3O_OK;
    pthread_create(&th[i], NULL, mt_worker, &jobs[i]);
  }
  int rc = W3O_OK;
  for (int i = 0; i < n_threads; i++) {
    pthread_join(th[i], NULL);
    if (jobs[i].rc != W3O_OK) rc = jobs[i].rc;
  }
  free(th);
  free(jobs);
  return rc;
}

int w3o_compress_blocks_mt(const uint8_t *in, size_t n, size_t block_size,
                           uint8_t *out, size_t out_stride, uint32_t *out_lens,
                           int n_threads) {
  if (!block_size) return W3O_E_ARG;
  mt_job j;
  memset(&j, 0, sizeof j);
  j.decode = 0;
  j.in = in;
  j.out = out;
  j.n = n;
  j.block_size = block_size;
  j.stride = out_stride;
  j.lens = out_lens;
  j.n_blocks = (n + block_size - 1) / block_size;
  return run_mt(j, n_threads);
}

int w3o_decompress_blocks_mt(const uint8_t *in, size_t in_stride,
                             const uint32_t *in_lens, uint8_t *out, size_t n,
                             size_t block_size, int n_threads) {
  if (!block_size) return W3O_E_ARG;
  mt_job j;
  memset(&j, 0, sizeof j);
  j.decode = 1;
  j.in = in;
  j.out = out;
  j.n = n;
  j.block_size = block_size;
  j.stride = in_stride;
  j.in_lens = in_lens;
  j.n_blocks = (n + block_size - 1) / block_size;
  return run_mt(j, n_threads);
}
