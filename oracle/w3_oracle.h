/*
 * w3_oracle.h — CPU restatement of weath3rb0i's predict→update→code path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product:
 * only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * `--impl reference` legs may load this library, and only as the checker or as
 * the reported CPU baseline.  The product (weath3rb0i_b200/) never links or
 * calls it and has no CPU fallback.
 *
 * Parity status:
 *   - arithmetic coder + bit I/O: PINNED by the reference's own known-answer
 *     tests (src/entropy_coding/arithmetic_coder_tests.rs:44-161,
 *     src/entropy_coding/io.rs:107-227), replayed in tests/test_oracle_kats.py.
 *   - OrderNEntropy / Counter / ACHistory / StationaryModel / header: the
 *     reference holds NO test or golden vector for these and no Rust toolchain
 *     exists in this image (oracle/_ref cannot be built) — "parity unpinned"
 *     by reference fixtures.  Mitigation: this C restatement and the
 *     independent Python restatement (oracle/w3_oracle.py) must agree
 *     byte-for-byte (tests/test_oracle_cross.py) and every function below
 *     cites the reference file:line it follows.
 *
 * All citations are relative to /root/reference/.
 */
#ifndef W3_ORACLE_H
#define W3_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define W3O_OK 0
#define W3O_E_CAP (-1) /* output buffer too small */
#define W3O_E_ARG (-2) /* bad argument */
#define W3O_E_MAGIC (-3) /* header magic mismatch (src/main.rs:123-124) */

/* ---- models ------------------------------------------------------------ */

/* Counter::p — src/models/counter.rs:13-18 */
uint16_t w3o_counter_p(uint16_t c0, uint16_t c1);
/* Counter::update — src/models/counter.rs:20-26; c = {c0, c1} in place */
void w3o_counter_update(uint16_t c[2], int bit);

/* StationaryModel::for_book1 table — src/models/ac_hash/stationary.rs:40-42 */
extern const uint16_t W3O_BOOK1_TABLE[8];

/* ACHistory::hash with max_bits = 8 and the book1 StationaryModel —
 * src/history/ac_history.rs:28-46 (+ EntropyWriter 50-88,
 * StationaryModel::{align,predict} src/models/ac_hash/stationary.rs:49-57).
 * `steps` (optional) receives the number of history bits consumed when the
 * 8-bit sink became full, i.e. the hash depends on the low `steps` bits only
 * (65 if the sink never filled within the 64 bits). */
uint32_t w3o_achash(uint64_t bits, uint64_t pos, int *steps);
/* same, arbitrary table / max_bits (for StationaryModel::from_table variants) */
uint32_t w3o_achash_ex(uint64_t bits, uint64_t pos, const uint16_t table[8],
                       int max_bits, int *steps);

/* ---- bit I/O (src/entropy_coding/io.rs) -------------------------------- */

typedef struct w3o_writer w3o_writer; /* ACWriter io.rs:52-101 */
w3o_writer *w3o_writer_new(uint8_t *buf, size_t cap);
void w3o_writer_free(w3o_writer *w);
void w3o_writer_write_bit(w3o_writer *w, int bit); /* io.rs:70-89 */
void w3o_writer_inc_parity(w3o_writer *w);         /* io.rs:66-68 */
void w3o_writer_flush(w3o_writer *w, uint32_t state); /* io.rs:91-100 */
size_t w3o_writer_len(const w3o_writer *w);  /* bytes produced so far */
int w3o_writer_overflowed(const w3o_writer *w);

typedef struct w3o_reader w3o_reader; /* ACReader io.rs:7-49 */
w3o_reader *w3o_reader_new(const uint8_t *buf, size_t n);
void w3o_reader_free(w3o_reader *r);
int w3o_reader_read_bit(w3o_reader *r);      /* io.rs:31-38, 0 past EOF */
uint32_t w3o_reader_read_u32(w3o_reader *r); /* io.rs:40-48, zero-padded */

/* ---- arithmetic coder with caller-supplied probabilities ---------------
 * Mirrors the helper pair in src/entropy_coding/arithmetic_coder_tests.rs:8-42:
 * `in` is read MSB-first one bit per probability. */
int w3o_ac_encode_probs(const uint8_t *in, size_t n_in, const uint16_t *probs,
                        size_t n_probs, uint8_t *out, size_t cap,
                        size_t *out_len);
/* decodes n_probs bits, packs them MSB-first into out (ceil(n/8) bytes) */
int w3o_ac_decode_probs(const uint8_t *in, size_t n_in, const uint16_t *probs,
                        size_t n_probs, uint8_t *out, size_t cap);

/* ---- the codec path (src/main.rs:89-152) ------------------------------- */

/* Stream bytes the reference writes AFTER its 12-byte header for a file equal
 * to in[0..n): compress loop main.rs:99-111 with init_model main.rs:146-152.
 * *out_len receives the full stream length even when it exceeds cap (then
 * the return is W3O_E_CAP and only cap bytes were stored). */
int w3o_compress_block(const uint8_t *in, size_t n, uint8_t *out, size_t cap,
                       size_t *out_len);
/* decompress loop main.rs:127-140 over one stream; bits past n_in read as 0 */
int w3o_decompress_block(const uint8_t *in, size_t n_in, uint8_t *out,
                         size_t n_out);

/* The same two loops with another StationaryModel table (NULL = for_book1):
 * StationaryModel::from_table — src/models/ac_hash/stationary.rs:36-38 */
int w3o_compress_block_ex(const uint8_t *in, size_t n, const uint16_t *table, uint8_t *out,
                          size_t cap, size_t *out_len);
int w3o_decompress_block_ex(const uint8_t *in, size_t n_in, const uint16_t *table, uint8_t *out,
                            size_t n_out);
/* The same two loops for any OrderNEntropy::new(bits_in_context, alignment_bits,
 * ACHistory::new(max_bits, StationaryModel::from_table(table))) — src/models/ordern_entropy.rs:15-24,
 * src/history/ac_history.rs:16-18.  max_bits is a parameter of its own (main.rs:151 passes 8 = 11 - 3,
 * the sweep binaries ctx_bits - alignment_bits, but nothing ties them).  bits_in_context <= 30. */
int w3o_compress_block_geo(const uint8_t *in, size_t n, const uint16_t *table, int bits_in_context,
                           int alignment_bits, int max_bits, uint8_t *out, size_t cap, size_t *out_len);
int w3o_decompress_block_geo(const uint8_t *in, size_t n_in, const uint16_t *table, int bits_in_context,
                             int alignment_bits, int max_bits, uint8_t *out, size_t n_out);
/* exec() of the entropy-hashing-ac sweep binary (src/bin/entropy-hashing-ac/main.rs:48-75) into the
 * ACStats sink (src/helpers.rs:60-90): *bit_count = bits counted; the printed csize is that / 8 */
int w3o_sweep_exec(const uint8_t *buf, size_t n, const uint16_t *table, int ctx_bits, int alignment_bits,
                   uint64_t *bit_count);
/* compress() of the ordern sweep binary (src/bin/ordern/main.rs:66-81): OrderN::new(ctx_bits,
 * alignment_bits) — the raw bit history as context (src/models/ordern.rs:4-46) — into ACStats */
int w3o_sweep_exec_ordern(const uint8_t *buf, size_t n, int ctx_bits, int alignment_bits, uint64_t *bit_count);
/* StationaryModel::new(buf) — src/models/ac_hash/stationary.rs:14-34: the 8
 * bit-position probabilities of a buffer (index 0 = a byte's MSB) */
void w3o_stationary_new(const uint8_t *buf, size_t n, uint16_t table[8]);

/* Whole reference file: "w30i" + u64 BE length + stream (main.rs:14-15,95-96) */
int w3o_compress_file(const uint8_t *in, size_t n, uint8_t *out, size_t cap,
                      size_t *out_len);
int w3o_decompress_file(const uint8_t *in, size_t n_in, uint8_t *out,
                        size_t cap, size_t *out_len);

/* Context trace: ctx used to code each of the 8*n bits (for kernel debugging
 * and the Appendix-B trace check). ctx_out has 8*n entries. */
void w3o_ctx_trace(const uint8_t *in, size_t n, uint16_t *ctx_out);

/* ---- multi-threaded CPU baseline (bench only) --------------------------
 * Runs w3o_compress_block / w3o_decompress_block over `n_blocks` independent
 * blocks with `n_threads` std threads, one block per thread at a time
 * (BASELINE.md §3).  Offsets/lengths arrays have n_blocks entries. */
int w3o_compress_blocks_mt(const uint8_t *in, size_t n, size_t block_size,
                           uint8_t *out, size_t out_stride, uint32_t *out_lens,
                           int n_threads);
int w3o_decompress_blocks_mt(const uint8_t *in, size_t in_stride,
                             const uint32_t *in_lens, uint8_t *out, size_t n,
                             size_t block_size, int n_threads);

/* ---- Tier B: state table and hash slots (w3_tierb_oracle.c) --------------
 * Dead code in the reference (no caller), pinned by its own dumps:
 * docs/state_table/state_table.csv, aux_state_table.csv, docs/hashslots.md:69-133. */
int w3o_state_table_size(void);      /* 3963 — src/state_table/naive.rs:11 */
int w3o_aux_state_table_size(void);  /* 990  — naive.rs:10 */
uint16_t w3o_state_next(uint16_t state, int bit); /* StateTable::next — src/state_table/mod.rs:41-43 */
uint16_t w3o_state_p(uint16_t state);             /* StateTable::p    — mod.rs:45-47 */
uint16_t w3o_aux_state_next(uint16_t state, int bit); /* gen_auxiliary_table — naive.rs:50-75 */
uint16_t w3o_aux_state_p(uint16_t state);
void w3o_state_next4(const uint16_t states[4], uint8_t nib, uint16_t out[4]); /* mod.rs:5-12 */
void w3o_state_p4(const uint16_t states[4], uint16_t out[4]);                 /* mod.rs:15-22 */
uint32_t w3o_hashmap_log_cell_count(uint64_t size);                    /* HashMap::new — src/hashmap.rs:7-10 */
uint64_t w3o_hashmap_cell_index(uint64_t hash, uint32_t log_cell_count); /* HashMap::get_slot — :25-28 */
int w3o_cell_get_slot(const uint8_t cell[96], uint64_t hash);          /* Cell::get_slot — :43-71 */
void w3o_slot_get_idx(int id, int bit_id, int nib_ctx, uint32_t *abs_idx, int *parity); /* :80-84 */
uint16_t w3o_slot_get_state(const uint8_t cell[96], int id, int bit_id, int nib_ctx);   /* :86-97 */
void w3o_slot_set_state(uint8_t cell[96], int id, int bit_id, int nib_ctx, uint16_t new_state); /* :99-112 */
void w3o_slot_set_nib(uint8_t cell[96], int id, uint8_t nib, const uint16_t s[4]); /* :114-119 */
void w3o_slot_get_nib(const uint8_t cell[96], int id, uint8_t nib, uint16_t s[4]); /* :121-128 */
uint16_t w3o_opinion_mix2(uint16_t p1, uint16_t p2); /* OpinionMixer2::mix — src/mixers/opinion_mixer2.rs:5-10 */

/* ---- Tier C: the hashed context-mixing stack "w3cm" (w3_cm_oracle.c) -----
 * PARITY UNPINNED by the reference (it has no logistic mixer, no APM, no caller of its hash map
 * or state table): w3_cm_oracle.c is the specification.  Judged on round trip and on identity of
 * the GPU path with this oracle. */
int w3o_cm_squash(int d);            /* 12-bit logistic, d in 1/256 units */
int w3o_cm_stretch(int p);           /* its inverse table */
int w3o_cm_state_bits(uint16_t s);   /* bits a state has seen (replacement policy) */
uint64_t w3o_cm_hash(uint64_t last8, int model, int low_nibble, unsigned hi);
int w3o_cm_find_slot(uint8_t cell[96], uint64_t hash); /* Cell::get_slot + replacement */
int w3o_cm_compress_block(const uint8_t *in, size_t n, int log_cells, uint8_t *out, size_t cap, size_t *out_len);
int w3o_cm_decompress_block(const uint8_t *in, size_t n_in, int log_cells, uint8_t *out, size_t n_out);

#ifdef __cplusplus
}
#endif
#endif
