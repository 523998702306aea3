/*
 * w3_codec.h — C ABI of the B200 block-parallel codec for weath3rb0i's
 * predict→update→code path.  libw3codec.so exports exactly these symbols.
 *
 * The reference has no FFI layer; its boundary for this path is the pair of
 * whole-stream functions private to the binary,
 *     fn compress(input_file: PathBuf, output_file: PathBuf) -> io::Result<()>   src/main.rs:89
 *     fn decompress(input_file: PathBuf, output_file: PathBuf) -> io::Result<()> src/main.rs:115
 * with the model fixed by init_model() (src/main.rs:146-152).  w3_compress /
 * w3_decompress are those two functions over memory buffers; INTEGRATION.md
 * shows the `extern "C"` block a maintainer adds to src/main.rs to call them.
 *
 * Plain pointers and sizes only; the caller owns every buffer.  All functions
 * return W3_OK or a negative W3_E_* and never abort.  There is no CPU
 * fallback: without a usable CUDA device w3_ctx_create fails with W3_E_CUDA.
 */
#ifndef W3_CODEC_H
#define W3_CODEC_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define W3_OK 0
#define W3_E_ARG (-1)    /* bad argument (null pointer, block size, device list) */
#define W3_E_CAP (-2)    /* output buffer too small; *out_len holds the size needed when known */
#define W3_E_MAGIC (-3)  /* not a w30i / w3Bk stream — the reference asserts here (src/main.rs:123-124) */
#define W3_E_CUDA (-4)   /* CUDA error or no device; w3_last_cuda_error() has the text */
#define W3_E_FORMAT (-5) /* container header or index is inconsistent / truncated */
#define W3_E_TABLE (-6)  /* threshold-table self-check failed at context creation */
#define W3_E_TOOBIG (-7) /* a single stream or block is larger than the kernels' 32-bit bit counters allow:
                            one block <= 256 MiB, one stream < 512 MiB.  The reference streams a file of any
                            length as ONE stream (src/main.rs:96, u64 length); files above the block limit
                            must use the w3Bk container (which the reference cannot read) */

/* flags for w3_compress */
#define W3_F_LEGACY 1u /* one block, reference file format: "w30i" + u64 BE length + stream
                          (src/main.rs:14-15, 95-96) — byte-identical to the reference's .bin */

#define W3_MAX_DEVICES 8
#define W3_DEFAULT_BLOCK (1u << 20)

typedef struct w3_ctx w3_ctx;

/* Creates a codec context on `n_devices` CUDA devices (ordinals in `devices`;
 * NULL = devices 0..n_devices-1).  Builds the entropy-hash threshold table
 * for StationaryModel::for_book1 (src/models/ac_hash/stationary.rs:40-42) on
 * each device.  Blocks are sharded over the devices as contiguous ranges, one
 * host thread per device, no collective. */
int w3_ctx_create(const int *devices, int n_devices, w3_ctx **out);
void w3_ctx_destroy(w3_ctx *ctx);
int w3_ctx_device_count(const w3_ctx *ctx);

/* Launch options of a context (no counterpart in the reference; tests and measurements use them
 * to pin a kernel family instead of the launcher's choice by block count). */
#define W3_OPT_KERNELS 1    /* 0 auto | 1 team: one CTA per block | 2 warp: one warp per block (round 1)
                               | 3 many-block: teams sharing a CTA (encode), one lane per block (decode) */
#define W3_OPT_ENC_TEAMS 2  /* blocks per CTA of the team encoder, 1..8; 0 auto */
#define W3_OPT_LANE_WARPS 3 /* warps per CTA of the lane decoder, 1..8 (split form, when asked for: blocks per CTA); 0 auto */
#define W3_OPT_ENC_CODER 5  /* team encoder with several teams per CTA: 0 one coder warp per CTA (lane b codes team b) | 1 one per team */
#define W3_OPT_SCRATCH_PERMILLE 6 /* first-pass stream region per block, in 1/1000 of the block size (0 = the default 1125):
                                     a test hook for the path that re-encodes the blocks whose stream did not fit */
#define W3_OPT_LANE_FORM 4  /* lane decoder: 0 auto | 1 one lane per block | 2 two lanes per block (speculating the next
                               context) | 3 the same with a probability table next to the counters | 4 split: a coder warp
                               and eight model warps per CTA of up to 8 blocks.  0 = split when an SM gets at most 8 blocks,
                               else two lanes per block */
int w3_ctx_set_option(w3_ctx *ctx, int option, int value);

/* ---- the StationaryModel of the entropy hash -------------------------------
 * init_model() fixes StationaryModel::for_book1 (src/main.rs:151) and that is
 * what a fresh context uses.  The reference also has from_table (src/models/
 * ac_hash/stationary.rs:36-38, e.g. for_enwik7 :44-46) and new(buf) (:14-34),
 * the latter used by its sweep binaries (src/bin/entropy-hashing-ac/main.rs:19).
 * A context on any other model than for_book1 writes container version 2,
 * which carries the 8 probabilities; the reference's own file format cannot
 * (W3_F_LEGACY then fails with W3_E_ARG).  w3_decompress switches the context
 * to the model of the container it is given. */

/* StationaryModel::from_table: rebuilds the threshold tables on every device. */
int w3_ctx_set_model(w3_ctx *ctx, const uint16_t table[8]);
int w3_ctx_get_model(const w3_ctx *ctx, uint16_t table[8]);
/* StationaryModel::new(buf): the 8 bit-position probabilities of in[0..n)
 * (index 0 = a byte's MSB), computed on the first device, halvings of
 * Counter::update included (src/models/counter.rs:20-26). */
int w3_train_model(w3_ctx *ctx, const uint8_t *in, size_t n, uint16_t table[8]);

/* ---- the geometry of the model stack -----------------------------------------
 * OrderNEntropy::new(bits_in_context, alignment_bits, ACHistory::new(max_bits, model)) —
 * src/models/ordern_entropy.rs:15-24, src/history/ac_history.rs:16-18.  init_model() fixes
 * (11, 3, 8) (src/main.rs:151) and that is what a fresh context uses; the reference's sweep
 * (src/bin/entropy-hashing-ac/main.rs:20-37) walks bits_in_context 8..30 x alignment_bits 0..4
 * with max_bits = bits_in_context - alignment_bits and finds 20 / 3 best on book1 (book1.log:76).
 * With any other geometry than (11, 3, 8) the counter table (4 << bits_in_context bytes per
 * block) lives in HBM, w3_compress writes container version 3 — byte 5..7 of the header =
 * bits_in_context, alignment_bits, max_bits, then the 8 model probabilities as in version 2 —
 * and W3_F_LEGACY fails (the reference's file format cannot say so).  w3_decompress reads the
 * geometry from the container.  bits_in_context <= 30, alignment_bits <= 7, max_bits <= 32. */
int w3_ctx_set_geometry(w3_ctx *ctx, int bits_in_context, int alignment_bits, int max_bits);
int w3_ctx_get_geometry(const w3_ctx *ctx, int *bits_in_context, int *alignment_bits, int *max_bits);

/* The parameter sweep of the reference's research binaries, on the GPU: every point is exec()
 * of src/bin/entropy-hashing-ac/main.rs:48-75 — the whole of in[0..n) coded as one stream by
 * OrderNEntropy::new(ctx_bits, alignment_bits, ACHistory::new(ctx_bits - alignment_bits, model))
 * into the bit-counting sink ACStats (src/helpers.rs:60-90).  bit_counts[i] = bits counted for
 * points[i]; the csize the reference prints is bit_counts[i] / 8 (helpers.rs:71-73).
 * table = the StationaryModel probabilities, or NULL for StationaryModel::new(in) as the sweep
 * binary does (main.rs:19 of that binary); table_used (optional) receives the 8 probabilities.
 * All points run concurrently, each with its own counter table in HBM (4 << ctx_bits bytes),
 * in as many batches as the first device's memory requires.  n <= 256 MiB. */
typedef struct w3_sweep_point {
  uint8_t ctx_bits, alignment_bits;
} w3_sweep_point;
/* history: which History feeds the context —
 *   W3_HISTORY_AC   ACHistory, the entropy hash (src/history/ac_history.rs): the `entropy-hashing-ac` binary
 *                   (and `entropy-hashing-ac-cached`, whose ACHistoryCached memoises the same hash)
 *   W3_HISTORY_RAW  the raw bits (RawHistory src/history/raw_history.rs = OrderN src/models/ordern.rs:36-45):
 *                   the `ordern` binary (src/bin/ordern/main.rs:16-33, 66-81); table is ignored */
#define W3_HISTORY_AC 0
#define W3_HISTORY_RAW 1
int w3_sweep(w3_ctx *ctx, int history, const uint8_t *in, size_t n, const uint16_t table[8],
             const w3_sweep_point *points, size_t n_points, uint64_t *bit_counts, uint16_t table_used[8]);

/* ---- the hashed context-mixing stack "w3cm" ----------------------------------
 * What BASELINE.json's north star describes and the reference only plans (README.md:6-10):
 * six hashed contexts per nibble in the reference's cell / slot layout (src/hashmap.rs,
 * docs/hashslots.md), bit histories through its 12-bit state table (src/state_table/naive.rs), a
 * logistic mixer, an APM, and the reference's arithmetic coder.  The reference has no such model
 * (no mixer, no APM, no caller of its hash map): there is nothing to be byte-identical to, the
 * specification is oracle/w3_cm_oracle.c and parity means identity with that oracle.
 * One warp per block, lanes = models; every block has its own table of 2^log_cells 96-byte cells
 * in HBM (log_cells 4..24).  Container:
 *   "w3Cm" | u8 1 | u8 models (6) | u8 log_cells | u8 0 | u64 BE n | u32 BE block_size | u32 BE n_blocks
 *   | n_blocks x u32 BE stream length | streams */
int w3_cm_compress(w3_ctx *ctx, const uint8_t *in, size_t n, size_t block_size, int log_cells, uint8_t *out,
                   size_t cap, size_t *out_len);
int w3_cm_decompress(w3_ctx *ctx, const uint8_t *in, size_t n_in, uint8_t *out, size_t cap, size_t *out_len);

/* Upper bound of w3_compress output for n input bytes. */
size_t w3_compress_bound(size_t n, size_t block_size);

/* compress() — src/main.rs:89-113 — over host buffers.
 * Splits in[0..n) into ceil(n / block_size) independent blocks (block_size a
 * multiple of 4, <= 256 MiB), codes every block exactly as the reference
 * codes a file of those bytes, and writes the container
 *   "w3Bk" | u8 version | u8[3] 0 | u64 BE n | u32 BE block_size | u32 BE n_blocks
 *   | versions 2 and 3: 8 x u16 BE StationaryModel probabilities (version 3: u8[3] = the geometry)
 *   | n_blocks x u32 BE stream length | streams back to back
 * or, with W3_F_LEGACY, the reference's own single-stream file. */
int w3_compress(w3_ctx *ctx, const uint8_t *in, size_t n, size_t block_size, uint32_t flags,
                uint8_t *out, size_t cap, size_t *out_len);

/* decompress() — src/main.rs:115-144 — accepts both formats. */
int w3_decompress(w3_ctx *ctx, const uint8_t *in, size_t n_in, uint8_t *out, size_t cap,
                  size_t *out_len);

/* The same for blocks [first_block, first_block + n_blocks) of a stream only: out receives those
 * blocks' bytes (block i of the range at (i - first_block) * block_size).  The streams of the
 * range must lie within in[0..n_in); what follows them may be missing.  This is what a process
 * that holds one shard of a job calls (contiguous block ranges, no exchange). */
int w3_decompress_blocks(w3_ctx *ctx, const uint8_t *in, size_t n_in, size_t first_block, size_t n_blocks,
                         uint8_t *out, size_t cap, size_t *out_len);

/* Reads only the header: uncompressed length (main.rs:119-126 reads the same field). */
int w3_decompressed_size(const uint8_t *in, size_t n_in, size_t *out_len);

/* ---- device-resident entry points (inputs already in HBM) ----------------
 * `dev` indexes the context's device list; `stream` is a cudaStream_t (NULL =
 * default stream); all work is enqueued on it and the call returns without
 * synchronising, so the caller can bracket it with CUDA events. */

/* Encodes n_blocks = ceil(n / block_size) blocks from d_in into per-block
 * scratch regions d_scratch + i * stride (stride a multiple of 4);
 * d_lens[i] = stream length; afterwards *d_overflow has bit 0 set if any stream
 * did not fit its region (bit 1: a kernel gave up, never expected).  d_in must be readable up to the next multiple of 4. */
int w3_encode_device(w3_ctx *ctx, int dev, const void *d_in, size_t n, size_t block_size,
                     void *d_scratch, size_t stride, uint32_t *d_lens, uint32_t *d_overflow,
                     void *stream);
/* d_offsets[0..n_blocks] = exclusive prefix sum of d_lens; streams copied
 * back to back into d_packed. */
int w3_pack_device(w3_ctx *ctx, int dev, const void *d_scratch, size_t stride,
                   const uint32_t *d_lens, size_t n_blocks, uint64_t *d_offsets, void *d_packed,
                   void *stream);
/* Decodes streams d_comp + d_offsets[i] .. d_offsets[i+1] into d_out (n bytes
 * total, block i at i * block_size).  d_comp needs 8 readable bytes of slack.
 * *d_err (device word, may be NULL) is OR-ed with a non-zero code if a kernel
 * gave up (never expected; a bounded inter-warp wait timed out). */
int w3_decode_device(w3_ctx *ctx, int dev, const void *d_comp, const uint64_t *d_offsets,
                     size_t n_blocks, size_t block_size, size_t n, void *d_out, uint32_t *d_err,
                     void *stream);

/* ---- self-tests of single stages, used by tests/ -------------------------- */
/* ACHistory::hash (src/history/ac_history.rs:28-46) for n (bits,pos) pairs, host arrays */
int w3_selftest_achash(w3_ctx *ctx, const uint64_t *bits, const uint32_t *pos, size_t n,
                       uint32_t *out);
/* Counter::p (src/models/counter.rs:13-18): all 65535^2 count pairs, float estimate
 * vs the u64 statement, on the device; *mismatches must come back 0 */
int w3_selftest_counter_p(w3_ctx *ctx, uint64_t *mismatches);

/* Measurement support (bench.py): random 32-byte-sector loads per second of the
 * context's first device from a scratch buffer of `buffer_bytes` (>= 4 GiB to
 * stay out of L2) — the denominator of the gather roofline SURVEY.md §8(d)
 * defines for this path (8 context lookups = 8 sectors per input byte).  No
 * counterpart in the reference. */
int w3_probe_sector_gather(w3_ctx *ctx, size_t buffer_bytes, double *sectors_per_second);
/* Measurement support: the name of the decode kernel the launcher takes for n_blocks blocks of block_size bytes
 * on a device of sm_count SMs under W3_OPT_KERNELS = kernels and W3_OPT_LANE_FORM = lane_form (0, 0 = a fresh
 * context).  A pure function of its arguments: no device needed. */
const char *w3_decode_kernel_name(unsigned n_blocks, unsigned block_size, int sm_count, int kernels, int lane_form);
/* the coder with caller-supplied probabilities — the harness of
 * src/entropy_coding/arithmetic_coder_tests.rs:8-42; `in` is read MSB-first */
int w3_selftest_ac_encode(w3_ctx *ctx, const uint8_t *in, size_t n_in, const uint16_t *probs,
                          size_t n_probs, uint8_t *out, size_t cap, size_t *out_len);
int w3_selftest_ac_decode(w3_ctx *ctx, const uint8_t *in, size_t n_in, const uint16_t *probs,
                          size_t n_probs, uint8_t *out, size_t cap);

/* ---- Tier B unit self-tests -------------------------
 * The reference's state table and hash-slot cells have no caller there (private modules,
 * src/lib.rs:11-13); these entry points expose the device implementations (csrc/w3_tierb.cuh)
 * so that tests can compare them with the reference's own dumps. */
/* NaiveStateTable (src/state_table/naive.rs:9-115) as built on the device: which = 0 the 3963-entry
 * table (docs/state_table/state_table.csv), 1 the 990-entry auxiliary table (aux_state_table.csv).
 * Each array receives n entries: StateTable::next(s, 0), next(s, 1), p(s) (src/state_table/mod.rs:41-49). */
int w3_selftest_state_table(w3_ctx *ctx, int which, uint16_t *next0, uint16_t *next1, uint16_t *prob, size_t n);
/* Replays n_ops cell operations in order on the device over n_cells 96-byte cells (src/hashmap.rs:32-129;
 * in/out: cells).  op = {code | cell << 8, a, b, c}, two result words per op:
 *   0 Slot::get_idx(id a, bit_id b, nib_ctx c) -> abs_idx, parity      1 Slot::get_state(a, b, c) -> state
 *   2 Slot::set_state(id | bit_id << 8 | nib_ctx << 16, state b)        3 Cell::get_slot(hash = b:a) -> slot id
 *   4 Slot::set_nib(id | nib << 8, s0 | s1 << 16, s2 | s3 << 16)        5 Slot::get_nib(id | nib << 8) -> packed states
 *   6 HashMap::get_slot's cell index (hash = b:a, log_cell_count c) -> index low, high */
int w3_selftest_hashslots(w3_ctx *ctx, const uint32_t *ops, size_t n_ops, uint8_t *cells, size_t n_cells,
                          uint32_t *results);

const char *w3_strerror(int code);
const char *w3_last_cuda_error(void);
const char *w3_version(void);

#ifdef __cplusplus
}
#endif
#endif
