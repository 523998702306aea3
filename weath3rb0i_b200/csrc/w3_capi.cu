/*
 * w3_capi.cu — the C ABI of include/w3_codec.h: block framing, sharding of
 * blocks over devices (one host thread per device, contiguous ranges, no
 * collective), staging buffers, and the launches in w3_kernels.cu.
 *
 * Mirrors compress()/decompress() of the reference (src/main.rs:89-144): same
 * header for the single-block "legacy" form, same model (init_model,
 * src/main.rs:146-152), same behaviour on a bad magic number (an error here,
 * an assert there).  There is no CPU path: every block is coded by a kernel.
 */
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include <algorithm>
#include <mutex>
#include <thread>
#include <vector>

#include "../../include/w3_codec.h"
#include "w3_core.cuh"
#include "w3_geo.h"
#include "w3_kernels.h"
#include "w3_tierb.cuh"

static thread_local char g_cuda_err[256] = "";
static char g_cuda_err_any[256] = ""; /* the last message of any thread: the per-device worker threads' */
static std::mutex g_err_mu;

/* records a message for w3_last_cuda_error(): in the calling thread and, for messages raised on
 * a per-device worker thread, where the thread that made the API call finds it */
static int fail_msg(int code, const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_cuda_err, sizeof g_cuda_err, fmt, ap);
  va_end(ap);
  std::lock_guard<std::mutex> lock(g_err_mu);
  snprintf(g_cuda_err_any, sizeof g_cuda_err_any, "%s", g_cuda_err);
  return code;
}
static int cuda_fail(cudaError_t e, const char *what) {
  return fail_msg(W3_E_CUDA, "%s: %s", what, cudaGetErrorString(e));
}
#define CK(call)                                              \
  do {                                                        \
    cudaError_t e_ = (call);                                  \
    if (e_ != cudaSuccess) return cuda_fail(e_, #call);       \
  } while (0)

#define W3_CM_MODELS_HOST 6 /* W3_CM_MODELS of w3_cm.cuh */
static const uint8_t MAGIC_LEGACY[4] = {'w', '3', '0', 'i'}; /* src/main.rs:14 */
static const uint8_t MAGIC_BLOCKS[4] = {'w', '3', 'B', 'k'};
static const size_t HDR_LEGACY = 12, HDR_BLOCKS = 24, HDR_MODEL = 16; /* v2 appends the 8 u16 BE probabilities */
static const size_t MAX_BLOCK = (size_t)256 << 20;
/* the kernels count stream bytes and bits in 32 bits: a stream is at most this long */
static const uint64_t MAX_STREAM = ((uint64_t)1 << 29) - 4096;

struct w3_buf {
  void *p = nullptr;
  size_t cap = 0;
  int ensure(size_t n) {
    if (n <= cap) return W3_OK;
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
    size_t want = n + (n >> 3) + 256;
    cudaError_t e = cudaMalloc(&p, want);
    if (e != cudaSuccess) { e = cudaMalloc(&p, n); want = n; }
    if (e != cudaSuccess) { p = nullptr; return cuda_fail(e, "cudaMalloc"); }
    cap = want;
    return W3_OK;
  }
  /* like ensure, but the first `keep` bytes survive a reallocation */
  int grow_keep(size_t n, size_t keep) {
    if (n <= cap) return W3_OK;
    void *q = nullptr;
    const size_t want = n + (n >> 2) + 256;
    cudaError_t e = cudaMalloc(&q, want);
    if (e != cudaSuccess) return cuda_fail(e, "cudaMalloc");
    if (p && keep) {
      e = cudaMemcpy(q, p, keep, cudaMemcpyDeviceToDevice);
      if (e != cudaSuccess) { cudaFree(q); return cuda_fail(e, "cudaMemcpy"); }
    }
    if (p) cudaFree(p);
    p = q;
    cap = want;
    return W3_OK;
  }
  void release() {
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
  }
};

struct w3_dev {
  int ordinal = 0, sm_count = 148;
  cudaStream_t stream = nullptr;
  w3_hashtab *d_tab = nullptr;
  w3_buf in, scratch, scratch2, lens, lens2, alt, offsets, packed, out;
  w3_buf geo_H, geo_counters, geo_jobs; /* run-time geometry (w3_geo.cu) */
  w3_buf cm_cells;                      /* hashed context mixing (w3_cm.cu) */
  uint32_t *d_flag = nullptr;
};

static const uint16_t BOOK1[8] = W3_BOOK1_TABLE; /* StationaryModel::for_book1 — stationary.rs:40-42 */

struct w3_ctx {
  int n = 0;
  w3_dev d[W3_MAX_DEVICES];
  uint16_t table[8]; /* the StationaryModel in force (init_model: for_book1, src/main.rs:151) */
  w3_launch_opts opts = {0, 0, 0, 0, 0, 0}; /* w3_ctx_set_option */
  /* OrderNEntropy::new(bits_in_context, alignment_bits, ACHistory::new(max_bits, ..)); init_model: 11, 3, 8 */
  int ctx_bits = W3_CTX_BITS, align_bits = W3_ALIGN_BITS, max_bits = W3_HASH_BITS;
  bool default_geometry() const { return ctx_bits == W3_CTX_BITS && align_bits == W3_ALIGN_BITS && max_bits == W3_HASH_BITS; }
  std::mutex mu;     /* one host call at a time per context */
};

static void put_be32(uint8_t *p, uint32_t v) { for (int i = 0; i < 4; i++) p[i] = (uint8_t)(v >> (24 - 8 * i)); }
static void put_be64(uint8_t *p, uint64_t v) { for (int i = 0; i < 8; i++) p[i] = (uint8_t)(v >> (56 - 8 * i)); }
static uint32_t get_be32(const uint8_t *p) { uint32_t v = 0; for (int i = 0; i < 4; i++) v = (v << 8) | p[i]; return v; }
static uint64_t get_be64(const uint8_t *p) { uint64_t v = 0; for (int i = 0; i < 8; i++) v = (v << 8) | p[i]; return v; }
static size_t round_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

static int dev_init(w3_dev &d, int ordinal) {
  d.ordinal = ordinal;
  CK(cudaSetDevice(ordinal));
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, ordinal));
  if (prop.major < 10) {
    return fail_msg(W3_E_CUDA, "device %d is sm_%d%d; this library is built for sm_100a only", ordinal, prop.major, prop.minor);
  }
  d.sm_count = prop.multiProcessorCount;
  CK(cudaStreamCreateWithFlags(&d.stream, cudaStreamNonBlocking));
  CK(w3_kernels_init());
  CK(w3_cm_init());
  CK(cudaMalloc(&d.d_tab, sizeof(w3_hashtab)));
  CK(cudaMalloc(&d.d_flag, sizeof(uint32_t)));
  return W3_OK;
}

/* (re)build the threshold tables of device d for a StationaryModel table */
static int dev_build_table(w3_dev &d, const uint16_t *table) {
  CK(cudaSetDevice(d.ordinal));
  /* the table is rebuilt in place: launches of the device-resident entry points still in flight
   * on the caller's streams read it, so let the device drain first (a model switch is rare) */
  CK(cudaDeviceSynchronize());
  std::vector<uint8_t> host(sizeof(w3_hashtab), 0);
  w3_hashtab *h = reinterpret_cast<w3_hashtab *>(host.data());
  for (int i = 0; i < 8; i++) h->table[i] = table[i];
  h->max_bits = W3_HASH_BITS;
  CK(cudaMemcpyAsync(d.d_tab, h, sizeof(w3_hashtab), cudaMemcpyHostToDevice, d.stream));
  CK(w3_launch_build_table(d.d_tab, d.stream));
  uint16_t bad = 1;
  CK(cudaMemcpyAsync(&bad, &d.d_tab->bad, sizeof bad, cudaMemcpyDeviceToHost, d.stream));
  CK(cudaStreamSynchronize(d.stream));
  return bad ? W3_E_TABLE : W3_OK;
}

static int ctx_set_model(w3_ctx *ctx, const uint16_t *table) {
  if (!memcmp(ctx->table, table, sizeof ctx->table)) return W3_OK;
  for (int g = 0; g < ctx->n; g++) {
    int rc = dev_build_table(ctx->d[g], table);
    if (rc != W3_OK) { /* leave the context on a usable model */
      for (int k = 0; k <= g; k++) dev_build_table(ctx->d[k], ctx->table);
      return rc;
    }
  }
  memcpy(ctx->table, table, sizeof ctx->table);
  return W3_OK;
}

static void dev_release(w3_dev &d) {
  cudaSetDevice(d.ordinal);
  d.in.release(); d.scratch.release(); d.scratch2.release(); d.lens.release(); d.lens2.release();
  d.alt.release(); d.offsets.release(); d.geo_H.release(); d.geo_counters.release(); d.geo_jobs.release(); d.cm_cells.release();
  d.packed.release(); d.out.release();
  if (d.d_tab) cudaFree(d.d_tab);
  if (d.d_flag) cudaFree(d.d_flag);
  if (d.stream) cudaStreamDestroy(d.stream);
  d.d_tab = nullptr; d.d_flag = nullptr; d.stream = nullptr;
}

/* run f(g) for every device of the context, on its own host thread when n > 1 */
template <class F>
static int for_each_device(w3_ctx *ctx, F &&f) {
  int rcs[W3_MAX_DEVICES] = {0};
  if (ctx->n == 1) return f(0);
  std::vector<std::thread> th;
  for (int g = 0; g < ctx->n; g++) th.emplace_back([&, g] { rcs[g] = f(g); });
  for (auto &t : th) t.join();
  for (int g = 0; g < ctx->n; g++) if (rcs[g] != W3_OK) return rcs[g];
  return W3_OK;
}

/* contiguous block range of device g */
static void block_range(size_t n_blocks, int n_dev, int g, size_t &b0, size_t &b1) {
  size_t per = (n_blocks + (size_t)n_dev - 1) / (size_t)n_dev;
  b0 = std::min(n_blocks, per * (size_t)g);
  b1 = std::min(n_blocks, b0 + per);
}

extern "C" {

const char *w3_version(void) { return "w3-b200 0.1 (sm_100a)"; }
const char *w3_last_cuda_error(void) {
  /* the calling thread's own message if it has one, else the last one any thread recorded (a
   * context on several devices raises them on its per-device worker threads) */
  static thread_local char out[256];
  std::lock_guard<std::mutex> lock(g_err_mu);
  snprintf(out, sizeof out, "%s", g_cuda_err_any[0] ? g_cuda_err_any : g_cuda_err);
  return out;
}
const char *w3_strerror(int code) {
  switch (code) {
    case W3_OK: return "ok";
    case W3_E_ARG: return "bad argument";
    case W3_E_CAP: return "output buffer too small";
    case W3_E_MAGIC: return "magic numbers don't match up - not a weath3rb0i stream";
    case W3_E_CUDA: return "CUDA error (see w3_last_cuda_error)";
    case W3_E_FORMAT: return "inconsistent or truncated block container";
    case W3_E_TABLE: return "entropy-hash threshold table failed its self-check";
    case W3_E_TOOBIG: return "stream or block too large for this build (one stream < 512 MiB, one block <= 256 MiB)";
    default: return "unknown error";
  }
}

int w3_ctx_create(const int *devices, int n_devices, w3_ctx **out) {
  if (!out || n_devices < 1 || n_devices > W3_MAX_DEVICES) return W3_E_ARG;
  *out = nullptr;
  int have = 0;
  cudaError_t e = cudaGetDeviceCount(&have);
  if (e != cudaSuccess) return cuda_fail(e, "cudaGetDeviceCount");
  if (have < 1) return fail_msg(W3_E_CUDA, "no CUDA device");
  w3_ctx *ctx = new w3_ctx();
  for (int i = 0; i < n_devices; i++) {
    int ord = devices ? devices[i] : i;
    if (ord < 0 || ord >= have) { w3_ctx_destroy(ctx); return W3_E_ARG; }
    ctx->n = i + 1;
    int rc = dev_init(ctx->d[i], ord);
    if (rc == W3_OK) rc = dev_build_table(ctx->d[i], BOOK1);
    if (rc != W3_OK) { w3_ctx_destroy(ctx); return rc; }
  }
  memcpy(ctx->table, BOOK1, sizeof ctx->table);
  *out = ctx;
  return W3_OK;
}

void w3_ctx_destroy(w3_ctx *ctx) {
  if (!ctx) return;
  for (int i = 0; i < ctx->n; i++) dev_release(ctx->d[i]);
  delete ctx;
}
int w3_ctx_device_count(const w3_ctx *ctx) { return ctx ? ctx->n : 0; }

int w3_ctx_set_model(w3_ctx *ctx, const uint16_t *table) {
  if (!ctx || !table) return W3_E_ARG;
  std::lock_guard<std::mutex> lock(ctx->mu);
  return ctx_set_model(ctx, table);
}

int w3_ctx_get_model(const w3_ctx *ctx, uint16_t *table) {
  if (!ctx || !table) return W3_E_ARG;
  std::lock_guard<std::mutex> lock(const_cast<w3_ctx *>(ctx)->mu);
  memcpy(table, ctx->table, sizeof ctx->table);
  return W3_OK;
}

int w3_ctx_set_option(w3_ctx *ctx, int option, int value) {
  if (!ctx || value < 0) return W3_E_ARG;
  std::lock_guard<std::mutex> lock(ctx->mu);
  switch (option) {
    case W3_OPT_KERNELS: if (value > 3) return W3_E_ARG; ctx->opts.kernels = value; return W3_OK;
    case W3_OPT_ENC_TEAMS: if (value > 8) return W3_E_ARG; ctx->opts.enc_teams = value; return W3_OK;
    case W3_OPT_LANE_WARPS: if (value > 8) return W3_E_ARG; ctx->opts.lane_warps = value; return W3_OK;
    case W3_OPT_SCRATCH_PERMILLE: if (value && (value < 10 || value > 20000)) return W3_E_ARG; ctx->opts.scratch_permille = value; return W3_OK;
    case W3_OPT_ENC_CODER: if (value > 1) return W3_E_ARG; ctx->opts.enc_coder = value; return W3_OK;
    case W3_OPT_LANE_FORM: if (value > 4) return W3_E_ARG; ctx->opts.lane_form = value; return W3_OK;
    default: return W3_E_ARG;
  }
}

int w3_train_model(w3_ctx *ctx, const uint8_t *in, size_t n, uint16_t *table) {
  if (!ctx || !table || (!in && n)) return W3_E_ARG;
  std::lock_guard<std::mutex> lock(ctx->mu);
  w3_dev &d = ctx->d[0];
  CK(cudaSetDevice(d.ordinal));
  const size_t n_chunks = (n + 1023) / 1024;
  int rc;
  if ((rc = d.in.ensure(n + 16)) != W3_OK) return rc;
  if ((rc = d.scratch.ensure(n_chunks * 16 + 16)) != W3_OK) return rc;
  if ((rc = d.offsets.ensure(64)) != W3_OK) return rc;
  if (n) CK(cudaMemcpyAsync(d.in.p, in, n, cudaMemcpyHostToDevice, d.stream));
  CK(w3_launch_train((const uint8_t *)d.in.p, n, (uint16_t *)d.scratch.p, (uint16_t *)d.offsets.p,
                     d.sm_count, d.stream));
  CK(cudaMemcpyAsync(table, d.offsets.p, 16, cudaMemcpyDeviceToHost, d.stream));
  CK(cudaStreamSynchronize(d.stream));
  return W3_OK;
}

size_t w3_compress_bound(size_t n, size_t block_size) {
  if (!block_size) block_size = W3_DEFAULT_BLOCK;
  size_t nb = (n + block_size - 1) / block_size;
  /* <= 16.001 coded bits per input bit (p >= 2^-16), + flush, per block */
  return HDR_BLOCKS + HDR_MODEL + 4 * nb + 17 * n + 8 * (nb + 1);
}

/* ---- device-resident entry points ---------------------------------------- */

int w3_encode_device(w3_ctx *ctx, int dev, const void *d_in, size_t n, size_t block_size,
                     void *d_scratch, size_t stride, uint32_t *d_lens, uint32_t *d_overflow,
                     void *stream) {
  if (!ctx || dev < 0 || dev >= ctx->n || !block_size || (block_size & 3) || block_size > MAX_BLOCK ||
      (stride & 3) || !d_lens || !d_overflow)
    return W3_E_ARG;
  size_t nb = (n + block_size - 1) / block_size;
  if (!nb) return W3_OK;
  if (!d_in || !d_scratch || nb > 0xFFFFFFFFull) return W3_E_ARG;
  std::lock_guard<std::mutex> lock(ctx->mu); /* no model switch (table rebuild) while this is being enqueued */
  w3_dev &d = ctx->d[dev];
  CK(cudaSetDevice(d.ordinal));
  CK(w3_launch_encode((const uint8_t *)d_in, n, (uint32_t)block_size, (uint32_t)nb,
                      (uint8_t *)d_scratch, stride, d_lens, d_overflow, d.d_tab, d.sm_count,
                      &ctx->opts, (cudaStream_t)stream));
  return W3_OK;
}

int w3_pack_device(w3_ctx *ctx, int dev, const void *d_scratch, size_t stride,
                   const uint32_t *d_lens, size_t n_blocks, uint64_t *d_offsets, void *d_packed,
                   void *stream) {
  if (!ctx || dev < 0 || dev >= ctx->n || !d_offsets || n_blocks > 0xFFFFFFFFull) return W3_E_ARG;
  if (n_blocks && (!d_scratch || !d_lens || !d_packed)) return W3_E_ARG;
  w3_dev &d = ctx->d[dev];
  CK(cudaSetDevice(d.ordinal));
  CK(w3_launch_pack((const uint8_t *)d_scratch, stride, d_lens, (uint32_t)n_blocks, d_offsets, nullptr,
                    (uint8_t *)d_packed, (cudaStream_t)stream));
  return W3_OK;
}

int w3_decode_device(w3_ctx *ctx, int dev, const void *d_comp, const uint64_t *d_offsets,
                     size_t n_blocks, size_t block_size, size_t n, void *d_out, uint32_t *d_err,
                     void *stream) {
  if (!ctx || dev < 0 || dev >= ctx->n || !block_size || (block_size & 3) || block_size > MAX_BLOCK)
    return W3_E_ARG;
  if (n_blocks != (n + block_size - 1) / block_size || n_blocks > 0xFFFFFFFFull) return W3_E_ARG;
  if (!n_blocks) return W3_OK;
  if (!d_comp || !d_offsets || !d_out) return W3_E_ARG;
  std::lock_guard<std::mutex> lock(ctx->mu);
  w3_dev &d = ctx->d[dev];
  CK(cudaSetDevice(d.ordinal));
  CK(w3_launch_decode((const uint8_t *)d_comp, d_offsets, (uint32_t)n_blocks, (uint32_t)block_size,
                      n, (uint8_t *)d_out, d_err ? d_err : d.d_flag, d.d_tab, d.sm_count,
                      &ctx->opts, (cudaStream_t)stream));
  return W3_OK;
}

/* ---- host-buffer entry points -------------------------------------------- */

struct enc_part {
  size_t b0 = 0, b1 = 0;
  uint64_t total = 0;
  std::vector<uint32_t> lens;
};

/* encode blocks [b0,b1) of `in` on device g; leaves the packed streams in
 * d.packed and their lengths in part.lens */
static int encode_part(w3_ctx *ctx, int g, const uint8_t *in, size_t n, size_t bs, enc_part &part) {
  w3_dev &d = ctx->d[g];
  const size_t nb = part.b1 - part.b0;
  part.lens.assign(nb, 0);
  part.total = 0;
  if (!nb) return W3_OK;
  CK(cudaSetDevice(d.ordinal));
  const size_t off = part.b0 * bs;
  const size_t bytes = std::min(n, part.b1 * bs) - off;
  int rc;
  if ((rc = d.in.ensure(round_up(bytes, 4) + 16)) != W3_OK) return rc;
  if ((rc = d.lens.ensure(nb * 4)) != W3_OK) return rc;
  if ((rc = d.offsets.ensure((nb + 1) * 8)) != W3_OK) return rc;
  CK(cudaMemcpyAsync(d.in.p, in + off, bytes, cudaMemcpyHostToDevice, d.stream));
  /* text codes to ~0.45 n; leave generous room.  A stream that does not fit is not stored beyond
   * its region but still measured (lens holds its full length: at most 17 n, since p >= 2^-16 caps
   * the cost of a bit at 16 + eps bits), so only the blocks that overflowed are coded again, each
   * into a region of exactly the size it asked for. */
  size_t stride = round_up(bs + (bs >> 3) + 64, 16);
  if (ctx->opts.scratch_permille) stride = round_up(bs / 1000 * (size_t)ctx->opts.scratch_permille + bs % 1000 + 16, 16);
  if ((rc = d.scratch.ensure(stride * nb)) != W3_OK) return rc;
  CK(cudaMemsetAsync(d.d_flag, 0, 4, d.stream));
  CK(w3_launch_encode((const uint8_t *)d.in.p, bytes, (uint32_t)bs, (uint32_t)nb,
                      (uint8_t *)d.scratch.p, stride, (uint32_t *)d.lens.p, d.d_flag, d.d_tab,
                      d.sm_count, &ctx->opts, d.stream));
  uint32_t flag = 0;
  CK(cudaMemcpyAsync(&flag, d.d_flag, 4, cudaMemcpyDeviceToHost, d.stream));
  CK(cudaMemcpyAsync(part.lens.data(), d.lens.p, nb * 4, cudaMemcpyDeviceToHost, d.stream));
  CK(cudaStreamSynchronize(d.stream));
  if (flag & W3_ERR_ABORT) return fail_msg(W3_E_CUDA, "encode kernel aborted (inter-warp wait timed out)");
  const uint64_t *d_alt = nullptr;
  if (flag & W3_ERR_OVERFLOW) {
    std::vector<size_t> ov;
    uint64_t need = 0;
    for (size_t i = 0; i < nb; i++)
      if (part.lens[i] > stride) { ov.push_back(i); need = std::max<uint64_t>(need, part.lens[i]); }
    if (ov.empty() || need > MAX_STREAM) return fail_msg(W3_E_TOOBIG, "a block's stream is %llu bytes", (unsigned long long)need);
    const size_t stride2 = round_up((size_t)need + 64, 16);
    if ((rc = d.scratch2.ensure(stride2 * ov.size())) != W3_OK) return rc;
    if ((rc = d.lens2.ensure(ov.size() * 4)) != W3_OK) return rc;
    if ((rc = d.alt.ensure(nb * 8)) != W3_OK) return rc;
    std::vector<uint64_t> alt(nb, 0);
    CK(cudaMemsetAsync(d.d_flag, 0, 4, d.stream));
    for (size_t j = 0; j < ov.size(); j++) {
      const size_t i = ov[j], boff = i * bs, blen = std::min(bs, bytes - boff);
      uint8_t *dst = (uint8_t *)d.scratch2.p + j * stride2;
      alt[i] = (uint64_t)(uintptr_t)dst;
      CK(w3_launch_encode((const uint8_t *)d.in.p + boff, blen, (uint32_t)bs, 1u, dst, stride2,
                          (uint32_t *)d.lens2.p + j, d.d_flag, d.d_tab, d.sm_count, &ctx->opts, d.stream));
    }
    CK(cudaMemcpyAsync(d.alt.p, alt.data(), nb * 8, cudaMemcpyHostToDevice, d.stream));
    CK(cudaMemcpyAsync(&flag, d.d_flag, 4, cudaMemcpyDeviceToHost, d.stream));
    CK(cudaStreamSynchronize(d.stream));
    if (flag) return fail_msg(W3_E_CUDA, "re-encoding an overflowed block failed (flag %u)", flag);
    d_alt = (const uint64_t *)d.alt.p;
  }
  for (uint32_t l : part.lens) part.total += l;
  if ((rc = d.packed.ensure(part.total + 16)) != W3_OK) return rc;
  CK(w3_launch_pack((const uint8_t *)d.scratch.p, stride, (const uint32_t *)d.lens.p, (uint32_t)nb,
                    (uint64_t *)d.offsets.p, d_alt, (uint8_t *)d.packed.p, d.stream));
  return W3_OK;
}


/* ---- run-time geometry (w3_geo.cu) ------------------------------------------ */

static bool geo_ok(int ctx_bits, int align_bits, int max_bits) {
  return ctx_bits >= 1 && ctx_bits <= 30 && align_bits >= 0 && align_bits <= 7 && align_bits <= ctx_bits &&
         max_bits >= 1 && max_bits <= 32;
}
/* blocks of one batch: the hash array costs 32 bytes per input byte, the counters 4 << ctx_bits per
 * block; a batch takes up to 40 % / 30 % of the device's free memory for them (every block of a batch
 * is one stream in flight: the more, the better the gathers overlap) */
static size_t geo_batch_blocks(const w3_dev &d, size_t bs, int ctx_bits, bool with_hash) {
  size_t free_b = 0, total_b = 0;
  if (cudaMemGetInfo(&free_b, &total_b) != cudaSuccess) free_b = (size_t)32 << 30;
  free_b += d.geo_H.cap + d.geo_counters.cap; /* what this path already holds is reusable */
  const size_t h_budget = std::max<size_t>((size_t)2 << 30, free_b / 10 * 4), c_budget = std::max<size_t>((size_t)2 << 30, free_b / 10 * 3);
  size_t by_cnt = std::max<size_t>(1, c_budget / ((size_t)4 << ctx_bits));
  size_t by_h = with_hash ? std::max<size_t>(1, h_budget / (32 * bs)) : by_cnt;
  return std::min<size_t>(std::min(by_cnt, by_h), 8192);
}

/* encode blocks [b0, b1) of `in` on device g with the context's geometry; same contract as encode_part */
static int encode_part_geo(w3_ctx *ctx, int g, const uint8_t *in, size_t n, size_t bs, enc_part &part) {
  w3_dev &d = ctx->d[g];
  const size_t nb = part.b1 - part.b0;
  part.lens.assign(nb, 0);
  part.total = 0;
  if (!nb) return W3_OK;
  CK(cudaSetDevice(d.ordinal));
  const size_t off = part.b0 * bs, bytes = std::min(n, part.b1 * bs) - off;
  int rc;
  if ((rc = d.in.ensure(round_up(bytes, 4) + 16)) != W3_OK) return rc;
  CK(cudaMemcpyAsync(d.in.p, in + off, bytes, cudaMemcpyHostToDevice, d.stream));
  const size_t per = geo_batch_blocks(d, bs, ctx->ctx_bits, true);
  const size_t nbat = std::min(per, nb);
  size_t stride = round_up(bs + (bs >> 3) + 64, 16);
  if ((rc = d.geo_H.ensure(nbat * bs * 32 + 64)) != W3_OK) return rc;
  if ((rc = d.geo_counters.ensure(nbat * ((size_t)4 << ctx->ctx_bits))) != W3_OK) return rc;
  if ((rc = d.geo_jobs.ensure(nbat * sizeof(w3_geo_job))) != W3_OK) return rc;
  if ((rc = d.scratch.ensure(stride * nbat)) != W3_OK) return rc;
  if ((rc = d.lens.ensure(nbat * 4)) != W3_OK) return rc;
  if ((rc = d.offsets.ensure((nbat + 1) * 8)) != W3_OK) return rc;
  if ((rc = d.packed.ensure(bytes + (bytes >> 3) + 64 * (nb + 1))) != W3_OK) return rc;
  const uint16_t *d_table = &d.d_tab->table[0];
  std::vector<w3_geo_job> jobs(nbat);
  uint64_t packed_at = 0;
  for (size_t first = 0; first < nb; first += nbat) {
    const size_t cnt = std::min(nbat, nb - first);
    const size_t boff = first * bs, blen = std::min(bytes, (first + cnt) * bs) - boff;
    for (int attempt = 0;; attempt++) {
      CK(w3_launch_geo_hash((const uint8_t *)d.in.p + boff, blen, (uint32_t)bs, d_table, (uint32_t)ctx->max_bits, 0u,
                            (uint32_t *)d.geo_H.p, d.sm_count, d.stream));
      CK(cudaMemsetAsync(d.geo_counters.p, 0, cnt * ((size_t)4 << ctx->ctx_bits), d.stream));
      for (size_t i = 0; i < cnt; i++) {
        w3_geo_job &j = jobs[i];
        memset(&j, 0, sizeof j);
        j.H = (const uint32_t *)d.geo_H.p + i * bs * 8;
        j.in = (const uint8_t *)d.in.p + boff + i * bs;
        j.out = (uint8_t *)d.scratch.p + i * stride;
        j.counters = (uint32_t *)d.geo_counters.p + i * ((size_t)1 << ctx->ctx_bits);
        j.nbytes = (uint32_t)std::min(bs, blen - i * bs);
        j.cap = (uint32_t)std::min<size_t>(stride, 0xFFFFFFF0u);
        j.ctx_bits = (uint32_t)ctx->ctx_bits;
        j.align_bits = (uint32_t)ctx->align_bits;
      }
      CK(cudaMemcpyAsync(d.geo_jobs.p, jobs.data(), cnt * sizeof(w3_geo_job), cudaMemcpyHostToDevice, d.stream));
      CK(w3_launch_geo_encode((w3_geo_job *)d.geo_jobs.p, (uint32_t)cnt, 1u, d.stream));
      CK(cudaMemcpyAsync(jobs.data(), d.geo_jobs.p, cnt * sizeof(w3_geo_job), cudaMemcpyDeviceToHost, d.stream));
      CK(cudaStreamSynchronize(d.stream));
      uint64_t need = 0;
      for (size_t i = 0; i < cnt; i++) if (jobs[i].flags) need = std::max<uint64_t>(need, jobs[i].out_len);
      if (!need) break;
      if (attempt || need > MAX_STREAM) return fail_msg(W3_E_TOOBIG, "a block's stream is %llu bytes", (unsigned long long)need);
      stride = round_up((size_t)need + 64, 16); /* every stream reports its full length: the second pass fits */
      if ((rc = d.scratch.ensure(stride * nbat)) != W3_OK) return rc;
    }
    std::vector<uint32_t> lens(cnt);
    for (size_t i = 0; i < cnt; i++) { lens[i] = jobs[i].out_len; part.lens[first + i] = lens[i]; }
    CK(cudaMemcpyAsync(d.lens.p, lens.data(), cnt * 4, cudaMemcpyHostToDevice, d.stream));
    uint64_t batch_total = 0;
    for (uint32_t l : lens) batch_total += l;
    if ((rc = d.packed.grow_keep(packed_at + batch_total + 16, packed_at)) != W3_OK) return rc; /* grows only after an overflow */
    CK(w3_launch_pack((const uint8_t *)d.scratch.p, stride, (const uint32_t *)d.lens.p, (uint32_t)cnt,
                      (uint64_t *)d.offsets.p, nullptr, (uint8_t *)d.packed.p + packed_at, d.stream));
    CK(cudaStreamSynchronize(d.stream));
    packed_at += batch_total;
  }
  part.total = packed_at;
  return W3_OK;
}

/* decode blocks [b0, b1) (streams at in + hdr + offs[]) on device g with the given geometry */
static int decode_part_geo(w3_ctx *ctx, int g, const uint8_t *streams, const std::vector<uint64_t> &offs, size_t b0,
                           size_t b1, size_t bs, size_t n, int ctx_bits, int align_bits, int max_bits, uint8_t *out_base,
                           size_t base) {
  w3_dev &d = ctx->d[g];
  CK(cudaSetDevice(d.ordinal));
  const size_t nbg = b1 - b0;
  const uint64_t c0 = offs[b0], clen = offs[b1] - c0;
  const size_t o0 = b0 * bs, obytes = std::min(n, b1 * bs) - o0;
  int rc;
  if ((rc = d.packed.ensure(clen + 16)) != W3_OK) return rc;
  if ((rc = d.out.ensure(round_up(obytes, 4) + 16)) != W3_OK) return rc;
  const size_t nbat = std::min(geo_batch_blocks(d, bs, ctx_bits, false), nbg);
  if ((rc = d.geo_counters.ensure(nbat * ((size_t)4 << ctx_bits))) != W3_OK) return rc;
  if ((rc = d.geo_jobs.ensure(nbat * sizeof(w3_geo_job))) != W3_OK) return rc;
  if (clen) CK(cudaMemcpyAsync(d.packed.p, streams + c0, clen, cudaMemcpyHostToDevice, d.stream));
  std::vector<w3_geo_job> jobs(nbat);
  for (size_t first = 0; first < nbg; first += nbat) {
    const size_t cnt = std::min(nbat, nbg - first);
    CK(cudaMemsetAsync(d.geo_counters.p, 0, cnt * ((size_t)4 << ctx_bits), d.stream));
    for (size_t i = 0; i < cnt; i++) {
      w3_geo_job &j = jobs[i];
      memset(&j, 0, sizeof j);
      const size_t b = b0 + first + i;
      j.in = (const uint8_t *)d.packed.p + (offs[b] - c0);
      j.out_len = (uint32_t)(offs[b + 1] - offs[b]);
      j.out = (uint8_t *)d.out.p + (first + i) * bs;
      j.counters = (uint32_t *)d.geo_counters.p + i * ((size_t)1 << ctx_bits);
      j.nbytes = (uint32_t)(std::min(n, (b + 1) * bs) - b * bs);
      j.ctx_bits = (uint32_t)ctx_bits;
      j.align_bits = (uint32_t)align_bits;
    }
    CK(cudaMemcpyAsync(d.geo_jobs.p, jobs.data(), cnt * sizeof(w3_geo_job), cudaMemcpyHostToDevice, d.stream));
    CK(w3_launch_geo_decode((w3_geo_job *)d.geo_jobs.p, (uint32_t)cnt, 1u, &d.d_tab->table[0], (uint32_t)max_bits, d.stream));
    CK(cudaStreamSynchronize(d.stream)); /* jobs[] is reused by the next batch */
  }
  CK(cudaMemcpyAsync(out_base + (o0 - base), d.out.p, obytes, cudaMemcpyDeviceToHost, d.stream));
  CK(cudaStreamSynchronize(d.stream));
  return W3_OK;
}

int w3_compress(w3_ctx *ctx, const uint8_t *in, size_t n, size_t block_size, uint32_t flags,
                uint8_t *out, size_t cap, size_t *out_len) {
  if (!ctx || !out_len || (!in && n) || (!out && cap)) return W3_E_ARG;
  const bool legacy = flags & W3_F_LEGACY;
  if (legacy) {
    /* the reference streams a file of any length (main.rs:96: u64); here one stream is one block */
    if (n > MAX_BLOCK) return fail_msg(W3_E_TOOBIG, "a reference-format (w30i) file holds one stream: %zu bytes exceed the %zu-byte block limit", n, MAX_BLOCK);
    block_size = std::max<size_t>(round_up(n, 4), 4);
  }
  if (!block_size) block_size = W3_DEFAULT_BLOCK;
  if (block_size & 3) return W3_E_ARG;
  if (block_size > MAX_BLOCK) return fail_msg(W3_E_TOOBIG, "block size %zu exceeds %zu", block_size, MAX_BLOCK);
  std::lock_guard<std::mutex> lock(ctx->mu);
  /* the reference codes an empty file as one empty stream (0xff); keep that
   * in legacy form, and an empty container has zero blocks */
  const size_t nb = legacy ? 1 : (n + block_size - 1) / block_size;
  if (nb > 0xFFFFFFFFull) return W3_E_ARG;
  /* the reference's file format has no room for a model: it implies for_book1 */
  const bool geo = !ctx->default_geometry(); /* container version 3: geometry and model in the header */
  const bool custom = geo || memcmp(ctx->table, BOOK1, sizeof BOOK1) != 0;
  if (legacy && custom) return W3_E_ARG;
  const size_t fixed = HDR_BLOCKS + (custom ? HDR_MODEL : 0);
  const size_t hdr = legacy ? HDR_LEGACY : fixed + 4 * nb;

  enc_part parts[W3_MAX_DEVICES];
  for (int g = 0; g < ctx->n; g++) block_range(nb, ctx->n, g, parts[g].b0, parts[g].b1);
  int rc;
  if (legacy && n == 0) {
    /* zero-length block: the kernel still runs (flush only) */
    parts[0].b0 = 0; parts[0].b1 = 1;
    for (int g = 1; g < ctx->n; g++) parts[g].b0 = parts[g].b1 = 1;
  }
  rc = for_each_device(ctx, [&](int g) {
    return geo ? encode_part_geo(ctx, g, in, n, block_size, parts[g]) : encode_part(ctx, g, in, n, block_size, parts[g]);
  });
  if (rc != W3_OK) return rc;

  uint64_t total = 0;
  for (int g = 0; g < ctx->n; g++) total += parts[g].total;
  *out_len = hdr + total;
  if (*out_len > cap) return W3_E_CAP;

  if (legacy) {
    memcpy(out, MAGIC_LEGACY, 4);   /* main.rs:95 */
    put_be64(out + 4, (uint64_t)n); /* main.rs:96 */
  } else {
    memcpy(out, MAGIC_BLOCKS, 4);
    out[4] = geo ? 3 : custom ? 2 : 1;
    out[5] = geo ? (uint8_t)ctx->ctx_bits : 0; out[6] = geo ? (uint8_t)ctx->align_bits : 0; out[7] = geo ? (uint8_t)ctx->max_bits : 0;
    put_be64(out + 8, (uint64_t)n);
    put_be32(out + 16, (uint32_t)block_size);
    put_be32(out + 20, (uint32_t)nb);
    if (custom)
      for (int i = 0; i < 8; i++) { out[HDR_BLOCKS + 2 * i] = (uint8_t)(ctx->table[i] >> 8); out[HDR_BLOCKS + 2 * i + 1] = (uint8_t)ctx->table[i]; }
    uint8_t *idx = out + fixed;
    for (int g = 0; g < ctx->n; g++)
      for (uint32_t l : parts[g].lens) { put_be32(idx, l); idx += 4; }
  }
  uint64_t base[W3_MAX_DEVICES], run = hdr;
  for (int g = 0; g < ctx->n; g++) { base[g] = run; run += parts[g].total; }
  return for_each_device(ctx, [&](int g) {
    w3_dev &d = ctx->d[g];
    if (!parts[g].total) return (int)W3_OK;
    CK(cudaSetDevice(d.ordinal));
    CK(cudaMemcpyAsync(out + base[g], d.packed.p, parts[g].total, cudaMemcpyDeviceToHost, d.stream));
    CK(cudaStreamSynchronize(d.stream));
    return (int)W3_OK;
  });
}

int w3_decompressed_size(const uint8_t *in, size_t n_in, size_t *out_len) {
  if (!in || !out_len) return W3_E_ARG;
  if (n_in < 4) return W3_E_MAGIC; /* read_exact fails in the reference (main.rs:121) */
  if (!memcmp(in, MAGIC_LEGACY, 4)) {
    if (n_in < HDR_LEGACY) return W3_E_FORMAT;
    *out_len = (size_t)get_be64(in + 4);
    return W3_OK;
  }
  if (!memcmp(in, MAGIC_BLOCKS, 4)) {
    if (n_in < HDR_BLOCKS || in[4] < 1 || in[4] > 3) return W3_E_FORMAT;
    if (in[4] >= 2 && n_in < HDR_BLOCKS + HDR_MODEL) return W3_E_FORMAT;
    if (in[4] == 3 && !geo_ok(in[5], in[6], in[7])) return W3_E_FORMAT;
    *out_len = (size_t)get_be64(in + 8);
    return W3_OK;
  }
  return W3_E_MAGIC;
}

/* decode blocks [B0, B1) of a stream (B1 = SIZE_MAX: all of them) into out; the caller holds ctx->mu */
static int decode_range(w3_ctx *ctx, const uint8_t *in, size_t n_in, size_t B0, size_t B1, uint8_t *out,
                        size_t cap, size_t *out_len) {
  size_t n = 0;
  int rc = w3_decompressed_size(in, n_in, &n);
  if (rc != W3_OK) return rc;
  const bool whole = B1 == SIZE_MAX;

  size_t bs, nb, hdr;
  int g_ctx = W3_CTX_BITS, g_align = W3_ALIGN_BITS, g_max = W3_HASH_BITS; /* version 3: from the header */
  std::vector<uint64_t> offs; /* nb + 1, relative to the first stream byte */
  if (!memcmp(in, MAGIC_LEGACY, 4)) {
    /* one stream running to the end of the file; a short file decodes with
     * zero bits past its end, as ACReader does (io.rs:23-25) */
    bs = std::max<size_t>(round_up(n, 4), 4);
    if (bs > MAX_BLOCK || n_in - HDR_LEGACY > MAX_STREAM)
      return fail_msg(W3_E_TOOBIG, "reference-format stream of %zu bytes for %zu bytes of output: one stream is one block here", n_in - HDR_LEGACY, n);
    nb = n ? 1 : 0;
    hdr = HDR_LEGACY;
    offs = {0, n_in - hdr};
    if ((rc = ctx_set_model(ctx, BOOK1)) != W3_OK) return rc; /* init_model — src/main.rs:151 */
  } else {
    bs = get_be32(in + 16);
    nb = get_be32(in + 20);
    if (!bs || (bs & 3) || bs > MAX_BLOCK || nb != (n + bs - 1) / bs) return W3_E_FORMAT;
    const size_t fixed = HDR_BLOCKS + (in[4] >= 2 ? HDR_MODEL : 0);
    hdr = fixed + 4 * nb;
    if (n_in < hdr) return W3_E_FORMAT;
    if (in[4] == 3) { g_ctx = in[5]; g_align = in[6]; g_max = in[7]; }
    /* version-2 and -3 containers carry their StationaryModel; version 1 means for_book1 */
    uint16_t table[8];
    for (int i = 0; i < 8; i++)
      table[i] = in[4] >= 2 ? (uint16_t)((in[HDR_BLOCKS + 2 * i] << 8) | in[HDR_BLOCKS + 2 * i + 1]) : BOOK1[i];
    if ((rc = ctx_set_model(ctx, table)) != W3_OK) return rc;
    offs.resize(nb + 1);
    uint64_t run = 0;
    for (size_t i = 0; i < nb; i++) {
      const uint32_t len = get_be32(in + fixed + 4 * i);
      if (len > MAX_STREAM) return W3_E_FORMAT;
      offs[i] = run;
      run += len;
    }
    offs[nb] = run;
    if (whole && hdr + run != n_in) return W3_E_FORMAT;
  }
  if (whole) { B0 = 0; B1 = nb; }
  if (B0 > B1 || B1 > nb) return W3_E_ARG;
  if (hdr + offs[B1] > n_in) return W3_E_FORMAT; /* the range's streams must be there */
  const size_t base = B0 * bs, bytes = std::min(n, B1 * bs) - std::min(n, base);
  *out_len = bytes;
  if (bytes > cap) return W3_E_CAP;
  if (B0 == B1) return W3_OK;

  return for_each_device(ctx, [&](int g) {
    size_t b0, b1;
    block_range(B1 - B0, ctx->n, g, b0, b1);
    if (b0 == b1) return (int)W3_OK;
    b0 += B0;
    b1 += B0;
    if (g_ctx != W3_CTX_BITS || g_align != W3_ALIGN_BITS || g_max != W3_HASH_BITS)
      return decode_part_geo(ctx, g, in + hdr, offs, b0, b1, bs, n, g_ctx, g_align, g_max, out, base);
    w3_dev &d = ctx->d[g];
    CK(cudaSetDevice(d.ordinal));
    const size_t nbg = b1 - b0;
    const uint64_t c0 = offs[b0], clen = offs[b1] - c0;
    const size_t o0 = b0 * bs, obytes = std::min(n, b1 * bs) - o0;
    int r;
    if ((r = d.packed.ensure(clen + 16)) != W3_OK) return r;
    if ((r = d.offsets.ensure((nbg + 1) * 8)) != W3_OK) return r;
    if ((r = d.out.ensure(round_up(obytes, 4) + 16)) != W3_OK) return r;
    std::vector<uint64_t> rel(nbg + 1);
    for (size_t i = 0; i <= nbg; i++) rel[i] = offs[b0 + i] - c0;
    CK(cudaMemsetAsync((uint8_t *)d.packed.p + clen, 0, 16, d.stream));
    if (clen) CK(cudaMemcpyAsync(d.packed.p, in + hdr + c0, clen, cudaMemcpyHostToDevice, d.stream));
    CK(cudaMemcpyAsync(d.offsets.p, rel.data(), (nbg + 1) * 8, cudaMemcpyHostToDevice, d.stream));
    CK(cudaMemsetAsync(d.d_flag, 0, 4, d.stream));
    CK(w3_launch_decode((const uint8_t *)d.packed.p, (const uint64_t *)d.offsets.p, (uint32_t)nbg,
                        (uint32_t)bs, obytes, (uint8_t *)d.out.p, d.d_flag, d.d_tab, d.sm_count,
                        &ctx->opts, d.stream));
    uint32_t flag = 0;
    CK(cudaMemcpyAsync(&flag, d.d_flag, 4, cudaMemcpyDeviceToHost, d.stream));
    CK(cudaMemcpyAsync(out + (o0 - base), d.out.p, obytes, cudaMemcpyDeviceToHost, d.stream));
    CK(cudaStreamSynchronize(d.stream));
    if (flag) return fail_msg(W3_E_CUDA, "decode kernel aborted (flag %u)", flag);
    return (int)W3_OK;
  });
}

int w3_decompress(w3_ctx *ctx, const uint8_t *in, size_t n_in, uint8_t *out, size_t cap,
                  size_t *out_len) {
  if (!ctx || !in || !out_len || (!out && cap)) return W3_E_ARG;
  std::lock_guard<std::mutex> lock(ctx->mu);
  return decode_range(ctx, in, n_in, 0, SIZE_MAX, out, cap, out_len);
}

int w3_decompress_blocks(w3_ctx *ctx, const uint8_t *in, size_t n_in, size_t first_block, size_t n_blocks,
                         uint8_t *out, size_t cap, size_t *out_len) {
  if (!ctx || !in || !out_len || (!out && cap) || first_block + n_blocks < first_block) return W3_E_ARG;
  std::lock_guard<std::mutex> lock(ctx->mu);
  return decode_range(ctx, in, n_in, first_block, first_block + n_blocks, out, cap, out_len);
}

int w3_ctx_set_geometry(w3_ctx *ctx, int bits_in_context, int alignment_bits, int max_bits) {
  if (!ctx || !geo_ok(bits_in_context, alignment_bits, max_bits)) return W3_E_ARG;
  std::lock_guard<std::mutex> lock(ctx->mu);
  ctx->ctx_bits = bits_in_context;
  ctx->align_bits = alignment_bits;
  ctx->max_bits = max_bits;
  return W3_OK;
}
int w3_ctx_get_geometry(const w3_ctx *ctx, int *bits_in_context, int *alignment_bits, int *max_bits) {
  if (!ctx || !bits_in_context || !alignment_bits || !max_bits) return W3_E_ARG;
  std::lock_guard<std::mutex> lock(const_cast<w3_ctx *>(ctx)->mu);
  *bits_in_context = ctx->ctx_bits;
  *alignment_bits = ctx->align_bits;
  *max_bits = ctx->max_bits;
  return W3_OK;
}

/* the sweep of src/bin/entropy-hashing-ac/main.rs:20-37: every point is exec() (:48-75) into ACStats */
int w3_sweep(w3_ctx *ctx, int history, const uint8_t *in, size_t n, const uint16_t *table, const w3_sweep_point *points,
             size_t n_points, uint64_t *bit_counts, uint16_t *table_used) {
  if (!ctx || (!in && n) || !points || !bit_counts || !n_points || n_points > 4096 || n > ((size_t)1 << 28)) return W3_E_ARG;
  if (history != W3_HISTORY_AC && history != W3_HISTORY_RAW) return W3_E_ARG;
  int hmax = 0;
  for (size_t i = 0; i < n_points; i++) {
    const int mb = (int)points[i].ctx_bits - (int)points[i].alignment_bits; /* ACHistory::new(ctx_bits - alignment_bits, ..) — :24 */
    if (!geo_ok(points[i].ctx_bits, points[i].alignment_bits, mb)) return W3_E_ARG;
    hmax = std::max(hmax, mb);
  }
  std::lock_guard<std::mutex> lock(ctx->mu);
  w3_dev &d = ctx->d[0];
  CK(cudaSetDevice(d.ordinal));
  int rc;
  const size_t n_chunks = (n + 1023) / 1024;
  if ((rc = d.in.ensure(round_up(n, 4) + 16)) != W3_OK) return rc;
  if ((rc = d.scratch.ensure(n_chunks * 16 + 16)) != W3_OK) return rc;
  if ((rc = d.offsets.ensure(64)) != W3_OK) return rc;
  if ((rc = d.geo_H.ensure(n * 32 + 64)) != W3_OK) return rc;
  if ((rc = d.geo_jobs.ensure(n_points * sizeof(w3_geo_job))) != W3_OK) return rc;
  CK(cudaMemsetAsync(d.in.p, 0, round_up(n, 4) + 16, d.stream));
  if (n) CK(cudaMemcpyAsync(d.in.p, in, n, cudaMemcpyHostToDevice, d.stream));
  uint16_t *d_table = (uint16_t *)d.offsets.p;
  if (table) CK(cudaMemcpyAsync(d_table, table, 16, cudaMemcpyHostToDevice, d.stream));
  else CK(w3_launch_train((const uint8_t *)d.in.p, n, (uint16_t *)d.scratch.p, d_table, d.sm_count, d.stream)); /* StationaryModel::new(&buf) — :19 */
  if (table_used) CK(cudaMemcpyAsync(table_used, d_table, 16, cudaMemcpyDeviceToHost, d.stream));
  /* one hash pass at the largest max_bits: a shorter hash is a prefix of a longer one */
  CK(w3_launch_geo_hash((const uint8_t *)d.in.p, n, (uint32_t)std::max<size_t>(round_up(n, 4), 4), d_table, (uint32_t)hmax,
                        history == W3_HISTORY_RAW ? 1u : 0u, (uint32_t *)d.geo_H.p, d.sm_count, d.stream));
  /* points in batches whose counter tables fit the device's free memory */
  size_t free_b = 0, total_b = 0;
  CK(cudaMemGetInfo(&free_b, &total_b));
  const size_t budget = std::max<size_t>((size_t)4 << 30, (free_b + d.geo_counters.cap) / 10 * 8);
  std::vector<size_t> order(n_points);
  for (size_t i = 0; i < n_points; i++) order[i] = i;
  std::sort(order.begin(), order.end(), [&](size_t a, size_t b) { return points[a].ctx_bits > points[b].ctx_bits; });
  std::vector<w3_geo_job> jobs;
  for (size_t at = 0; at < n_points;) {
    size_t used = 0, cnt = 0;
    while (at + cnt < n_points) {
      const size_t sz = (size_t)4 << points[order[at + cnt]].ctx_bits;
      if (cnt && used + sz > budget) break;
      used += sz;
      cnt++;
    }
    if ((rc = d.geo_counters.ensure(used)) != W3_OK) return rc;
    CK(cudaMemsetAsync(d.geo_counters.p, 0, used, d.stream));
    jobs.assign(cnt, w3_geo_job());
    size_t coff = 0;
    for (size_t i = 0; i < cnt; i++) {
      const w3_sweep_point &pt = points[order[at + i]];
      w3_geo_job &j = jobs[i];
      memset(&j, 0, sizeof j);
      j.H = (const uint32_t *)d.geo_H.p;
      j.in = (const uint8_t *)d.in.p;
      j.out = nullptr; /* ACStats: count only */
      j.counters = (uint32_t *)((uint8_t *)d.geo_counters.p + coff);
      coff += (size_t)4 << pt.ctx_bits;
      j.nbytes = (uint32_t)n;
      j.ctx_bits = pt.ctx_bits;
      j.align_bits = pt.alignment_bits;
    }
    CK(cudaMemcpyAsync(d.geo_jobs.p, jobs.data(), cnt * sizeof(w3_geo_job), cudaMemcpyHostToDevice, d.stream));
    CK(w3_launch_geo_encode((w3_geo_job *)d.geo_jobs.p, (uint32_t)cnt, 1u, d.stream));
    CK(cudaMemcpyAsync(jobs.data(), d.geo_jobs.p, cnt * sizeof(w3_geo_job), cudaMemcpyDeviceToHost, d.stream));
    CK(cudaStreamSynchronize(d.stream));
    for (size_t i = 0; i < cnt; i++) bit_counts[order[at + i]] = jobs[i].bit_count;
    at += cnt;
  }
  return W3_OK;
}

/* ---- the hashed context-mixing stack (w3_cm.cu; specification: oracle/w3_cm_oracle.c) -------- */

static const uint8_t MAGIC_CM[4] = {'w', '3', 'C', 'm'};
static const size_t HDR_CM = 24;

/* blocks whose cell tables (96 << log_cells bytes each) are resident at once */
static size_t cm_batch_blocks(int log_cells) {
  return std::max<size_t>(1, std::min<size_t>(4096, ((size_t)32 << 30) / ((size_t)W3_CELL_BYTES << log_cells)));
}

int w3_cm_compress(w3_ctx *ctx, const uint8_t *in, size_t n, size_t block_size, int log_cells, uint8_t *out,
                   size_t cap, size_t *out_len) {
  if (!ctx || !out_len || (!in && n) || (!out && cap) || log_cells < 4 || log_cells > 24) return W3_E_ARG;
  if (!block_size) block_size = W3_DEFAULT_BLOCK;
  if (block_size & 3) return W3_E_ARG;
  if (block_size > ((size_t)64 << 20)) return fail_msg(W3_E_TOOBIG, "block size %zu exceeds 64 MiB", block_size);
  std::lock_guard<std::mutex> lock(ctx->mu);
  const size_t nb = (n + block_size - 1) / block_size;
  if (nb > 0xFFFFFFFFull) return W3_E_ARG;
  const size_t hdr = HDR_CM + 4 * nb;
  enc_part parts[W3_MAX_DEVICES];
  for (int g = 0; g < ctx->n; g++) block_range(nb, ctx->n, g, parts[g].b0, parts[g].b1);
  int rc = for_each_device(ctx, [&](int g) {
    enc_part &part = parts[g];
    w3_dev &d = ctx->d[g];
    const size_t nbg = part.b1 - part.b0;
    part.lens.assign(nbg, 0);
    part.total = 0;
    if (!nbg) return (int)W3_OK;
    CK(cudaSetDevice(d.ordinal));
    const size_t off = part.b0 * block_size, bytes = std::min(n, part.b1 * block_size) - off;
    const size_t nbat = std::min(cm_batch_blocks(log_cells), nbg);
    size_t stride = round_up(block_size + (block_size >> 3) + 64, 16);
    int r;
    if ((r = d.in.ensure(round_up(bytes, 4) + 16)) != W3_OK) return r;
    if ((r = d.cm_cells.ensure(nbat * ((size_t)W3_CELL_BYTES << log_cells))) != W3_OK) return r;
    if ((r = d.scratch.ensure(stride * nbat)) != W3_OK) return r;
    if ((r = d.lens.ensure(nbat * 4)) != W3_OK) return r;
    if ((r = d.offsets.ensure((nbat + 1) * 8)) != W3_OK) return r;
    if ((r = d.packed.ensure(bytes + (bytes >> 3) + 64 * (nbg + 1))) != W3_OK) return r;
    CK(cudaMemcpyAsync(d.in.p, in + off, bytes, cudaMemcpyHostToDevice, d.stream));
    uint64_t packed_at = 0;
    for (size_t first = 0; first < nbg; first += nbat) {
      const size_t cnt = std::min(nbat, nbg - first);
      const size_t boff = first * block_size, blen = std::min(bytes, (first + cnt) * block_size) - boff;
      for (int attempt = 0;; attempt++) {
        CK(cudaMemsetAsync(d.cm_cells.p, 0, cnt * ((size_t)W3_CELL_BYTES << log_cells), d.stream));
        CK(cudaMemsetAsync(d.d_flag, 0, 4, d.stream));
        CK(w3_launch_cm_encode((const uint8_t *)d.in.p + boff, blen, (uint32_t)block_size, (uint32_t)cnt, (uint8_t *)d.scratch.p,
                               stride, (uint32_t *)d.lens.p, d.d_flag, (uint8_t *)d.cm_cells.p, (uint32_t)log_cells, d.stream));
        uint32_t flag = 0;
        CK(cudaMemcpyAsync(&flag, d.d_flag, 4, cudaMemcpyDeviceToHost, d.stream));
        CK(cudaMemcpyAsync(part.lens.data() + first, d.lens.p, cnt * 4, cudaMemcpyDeviceToHost, d.stream));
        CK(cudaStreamSynchronize(d.stream));
        if (!flag) break;
        if (attempt) return fail_msg(W3_E_CUDA, "w3cm: a stream exceeded 12 bits per bit (flag %u)", flag);
        /* the final probability stays within [1, 4095] / 4096: a coded bit costs at most 12 bits */
        stride = round_up(12 * block_size + block_size / 2 + 64, 16);
        if ((r = d.scratch.ensure(stride * nbat)) != W3_OK) return r;
      }
      uint64_t batch_total = 0;
      for (size_t i = 0; i < cnt; i++) batch_total += part.lens[first + i];
      if ((r = d.packed.grow_keep(packed_at + batch_total + 16, packed_at)) != W3_OK) return r;
      CK(w3_launch_pack((const uint8_t *)d.scratch.p, stride, (const uint32_t *)d.lens.p, (uint32_t)cnt,
                        (uint64_t *)d.offsets.p, nullptr, (uint8_t *)d.packed.p + packed_at, d.stream));
      CK(cudaStreamSynchronize(d.stream));
      packed_at += batch_total;
    }
    part.total = packed_at;
    return (int)W3_OK;
  });
  if (rc != W3_OK) return rc;
  uint64_t total = 0;
  for (int g = 0; g < ctx->n; g++) total += parts[g].total;
  *out_len = hdr + total;
  if (*out_len > cap) return W3_E_CAP;
  memcpy(out, MAGIC_CM, 4);
  out[4] = 1; out[5] = (uint8_t)W3_CM_MODELS_HOST; out[6] = (uint8_t)log_cells; out[7] = 0;
  put_be64(out + 8, (uint64_t)n);
  put_be32(out + 16, (uint32_t)block_size);
  put_be32(out + 20, (uint32_t)nb);
  uint8_t *idx = out + HDR_CM;
  for (int g = 0; g < ctx->n; g++)
    for (uint32_t l : parts[g].lens) { put_be32(idx, l); idx += 4; }
  uint64_t base[W3_MAX_DEVICES], run = hdr;
  for (int g = 0; g < ctx->n; g++) { base[g] = run; run += parts[g].total; }
  return for_each_device(ctx, [&](int g) {
    w3_dev &d = ctx->d[g];
    if (!parts[g].total) return (int)W3_OK;
    CK(cudaSetDevice(d.ordinal));
    CK(cudaMemcpyAsync(out + base[g], d.packed.p, parts[g].total, cudaMemcpyDeviceToHost, d.stream));
    CK(cudaStreamSynchronize(d.stream));
    return (int)W3_OK;
  });
}

int w3_cm_decompress(w3_ctx *ctx, const uint8_t *in, size_t n_in, uint8_t *out, size_t cap, size_t *out_len) {
  if (!ctx || !in || !out_len || (!out && cap)) return W3_E_ARG;
  if (n_in < 4 || memcmp(in, MAGIC_CM, 4)) return W3_E_MAGIC;
  if (n_in < HDR_CM || in[4] != 1 || in[5] != W3_CM_MODELS_HOST || in[6] < 4 || in[6] > 24) return W3_E_FORMAT;
  const int log_cells = in[6];
  const size_t n = (size_t)get_be64(in + 8), bs = get_be32(in + 16), nb = get_be32(in + 20);
  if (!bs || (bs & 3) || bs > ((size_t)64 << 20) || nb != (n + bs - 1) / bs) return W3_E_FORMAT;
  const size_t hdr = HDR_CM + 4 * nb;
  if (n_in < hdr) return W3_E_FORMAT;
  std::vector<uint64_t> offs(nb + 1);
  uint64_t run = 0;
  for (size_t i = 0; i < nb; i++) {
    const uint32_t len = get_be32(in + HDR_CM + 4 * i);
    if (len > MAX_STREAM) return W3_E_FORMAT;
    offs[i] = run;
    run += len;
  }
  offs[nb] = run;
  if (hdr + run != n_in) return W3_E_FORMAT;
  *out_len = n;
  if (n > cap) return W3_E_CAP;
  if (!nb) return W3_OK;
  std::lock_guard<std::mutex> lock(ctx->mu);
  return for_each_device(ctx, [&](int g) {
    size_t b0, b1;
    block_range(nb, ctx->n, g, b0, b1);
    if (b0 == b1) return (int)W3_OK;
    w3_dev &d = ctx->d[g];
    CK(cudaSetDevice(d.ordinal));
    const size_t nbg = b1 - b0;
    const uint64_t c0 = offs[b0], clen = offs[b1] - c0;
    const size_t o0 = b0 * bs, obytes = std::min(n, b1 * bs) - o0;
    const size_t nbat = std::min(cm_batch_blocks(log_cells), nbg);
    int r;
    if ((r = d.packed.ensure(clen + 16)) != W3_OK) return r;
    if ((r = d.offsets.ensure((nbg + 1) * 8)) != W3_OK) return r;
    if ((r = d.out.ensure(round_up(obytes, 4) + 16)) != W3_OK) return r;
    if ((r = d.cm_cells.ensure(nbat * ((size_t)W3_CELL_BYTES << log_cells))) != W3_OK) return r;
    std::vector<uint64_t> rel(nbg + 1);
    for (size_t i = 0; i <= nbg; i++) rel[i] = offs[b0 + i] - c0;
    CK(cudaMemsetAsync((uint8_t *)d.packed.p + clen, 0, 16, d.stream));
    if (clen) CK(cudaMemcpyAsync(d.packed.p, in + hdr + c0, clen, cudaMemcpyHostToDevice, d.stream));
    CK(cudaMemcpyAsync(d.offsets.p, rel.data(), (nbg + 1) * 8, cudaMemcpyHostToDevice, d.stream));
    for (size_t first = 0; first < nbg; first += nbat) {
      const size_t cnt = std::min(nbat, nbg - first);
      const size_t boff = first * bs, blen = std::min(obytes, (first + cnt) * bs) - boff;
      CK(cudaMemsetAsync(d.cm_cells.p, 0, cnt * ((size_t)W3_CELL_BYTES << log_cells), d.stream));
      CK(w3_launch_cm_decode((const uint8_t *)d.packed.p, (const uint64_t *)d.offsets.p + first, blen, (uint32_t)bs, (uint32_t)cnt,
                             (uint8_t *)d.out.p + boff, d.d_flag, (uint8_t *)d.cm_cells.p, (uint32_t)log_cells, d.stream));
    }
    CK(cudaMemcpyAsync(out + o0, d.out.p, obytes, cudaMemcpyDeviceToHost, d.stream));
    CK(cudaStreamSynchronize(d.stream));
    return (int)W3_OK;
  });
}

/* ---- self-tests ------------------------------------------------------------ */

int w3_selftest_achash(w3_ctx *ctx, const uint64_t *bits, const uint32_t *pos, size_t n,
                       uint32_t *out) {
  if (!ctx || !bits || !pos || !out || n > 0xFFFFFFFFull) return W3_E_ARG;
  if (!n) return W3_OK;
  std::lock_guard<std::mutex> lock(ctx->mu);
  w3_dev &d = ctx->d[0];
  CK(cudaSetDevice(d.ordinal));
  int rc;
  if ((rc = d.in.ensure(n * 8)) != W3_OK) return rc;
  if ((rc = d.lens.ensure(n * 4)) != W3_OK) return rc;
  if ((rc = d.out.ensure(n * 4)) != W3_OK) return rc;
  CK(cudaMemcpyAsync(d.in.p, bits, n * 8, cudaMemcpyHostToDevice, d.stream));
  CK(cudaMemcpyAsync(d.lens.p, pos, n * 4, cudaMemcpyHostToDevice, d.stream));
  CK(w3_launch_hash_array((const uint64_t *)d.in.p, (const uint32_t *)d.lens.p, (uint32_t)n,
                          (uint32_t *)d.out.p, d.d_tab, d.stream));
  CK(cudaMemcpyAsync(out, d.out.p, n * 4, cudaMemcpyDeviceToHost, d.stream));
  CK(cudaStreamSynchronize(d.stream));
  return W3_OK;
}

int w3_selftest_counter_p(w3_ctx *ctx, uint64_t *mismatches) {
  if (!ctx || !mismatches) return W3_E_ARG;
  std::lock_guard<std::mutex> lock(ctx->mu);
  w3_dev &d = ctx->d[0];
  CK(cudaSetDevice(d.ordinal));
  int rc;
  if ((rc = d.offsets.ensure(8)) != W3_OK) return rc;
  CK(cudaMemsetAsync(d.offsets.p, 0, 8, d.stream));
  CK(w3_launch_counter_p_check((unsigned long long *)d.offsets.p, d.stream));
  CK(cudaMemcpyAsync(mismatches, d.offsets.p, 8, cudaMemcpyDeviceToHost, d.stream));
  CK(cudaStreamSynchronize(d.stream));
  return W3_OK;
}

int w3_probe_sector_gather(w3_ctx *ctx, size_t buffer_bytes, double *sectors_per_second) {
  if (!ctx || !sectors_per_second || buffer_bytes < (1u << 20)) return W3_E_ARG;
  std::lock_guard<std::mutex> lock(ctx->mu);
  w3_dev &d = ctx->d[0];
  CK(cudaSetDevice(d.ordinal));
  void *buf = nullptr;
  unsigned long long *sink = nullptr;
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  int rc = W3_OK;
  float ms = 0.f;
  const uint64_t n_sectors = std::min<uint64_t>(buffer_bytes / 32, 0xFFFFFFFFull);
  const uint32_t blocks = (uint32_t)d.sm_count * 32u, per_thread = 64u;
  auto fin = [&](cudaError_t e) { if (e != cudaSuccess && rc == W3_OK) rc = cuda_fail(e, "w3_probe_sector_gather"); };
  fin(cudaMalloc(&buf, n_sectors * 32));
  if (rc == W3_OK) fin(cudaMalloc((void **)&sink, 8));
  if (rc == W3_OK) fin(cudaMemsetAsync(buf, 0x5A, n_sectors * 32, d.stream));
  if (rc == W3_OK) fin(cudaMemsetAsync(sink, 0, 8, d.stream));
  if (rc == W3_OK) fin(cudaEventCreate(&e0));
  if (rc == W3_OK) fin(cudaEventCreate(&e1));
  if (rc == W3_OK) fin(w3_launch_gather_probe(buf, n_sectors, blocks, per_thread, sink, d.stream)); /* warm-up */
  if (rc == W3_OK) fin(cudaEventRecord(e0, d.stream));
  for (int r = 0; r < 3 && rc == W3_OK; r++) fin(w3_launch_gather_probe(buf, n_sectors, blocks, per_thread, sink, d.stream));
  if (rc == W3_OK) fin(cudaEventRecord(e1, d.stream));
  if (rc == W3_OK) fin(cudaStreamSynchronize(d.stream));
  if (rc == W3_OK) fin(cudaEventElapsedTime(&ms, e0, e1));
  if (rc == W3_OK) *sectors_per_second = 3.0 * (double)blocks * 256.0 * per_thread / ((double)ms * 1e-3);
  if (e0) cudaEventDestroy(e0);
  if (e1) cudaEventDestroy(e1);
  if (sink) cudaFree(sink);
  if (buf) cudaFree(buf);
  return rc;
}

int w3_selftest_ac_encode(w3_ctx *ctx, const uint8_t *in, size_t n_in, const uint16_t *probs,
                          size_t n_probs, uint8_t *out, size_t cap, size_t *out_len) {
  if (!ctx || !in || !probs || !out || !out_len || n_in * 8 < n_probs || n_probs > 0x7FFFFFFF)
    return W3_E_ARG;
  std::lock_guard<std::mutex> lock(ctx->mu);
  w3_dev &d = ctx->d[0];
  CK(cudaSetDevice(d.ordinal));
  int rc;
  if ((rc = d.in.ensure(n_in + 16)) != W3_OK) return rc;
  if ((rc = d.scratch.ensure(n_probs * 2 + 16)) != W3_OK) return rc;
  if ((rc = d.out.ensure(cap + 16)) != W3_OK) return rc;
  if ((rc = d.lens.ensure(4)) != W3_OK) return rc;
  CK(cudaMemcpyAsync(d.in.p, in, n_in, cudaMemcpyHostToDevice, d.stream));
  CK(cudaMemcpyAsync(d.scratch.p, probs, n_probs * 2, cudaMemcpyHostToDevice, d.stream));
  CK(w3_launch_ac_encode_probs((const uint8_t *)d.in.p, (const uint16_t *)d.scratch.p,
                               (uint32_t)n_probs, (uint8_t *)d.out.p, (uint32_t)cap,
                               (uint32_t *)d.lens.p, d.stream));
  uint32_t len = 0;
  CK(cudaMemcpyAsync(&len, d.lens.p, 4, cudaMemcpyDeviceToHost, d.stream));
  CK(cudaStreamSynchronize(d.stream));
  *out_len = len;
  if (len > cap) return W3_E_CAP;
  CK(cudaMemcpy(out, d.out.p, len, cudaMemcpyDeviceToHost));
  return W3_OK;
}

int w3_selftest_ac_decode(w3_ctx *ctx, const uint8_t *in, size_t n_in, const uint16_t *probs,
                          size_t n_probs, uint8_t *out, size_t cap) {
  if (!ctx || !in || !probs || !out || cap * 8 < n_probs || n_probs > 0x7FFFFFFF) return W3_E_ARG;
  std::lock_guard<std::mutex> lock(ctx->mu);
  w3_dev &d = ctx->d[0];
  CK(cudaSetDevice(d.ordinal));
  int rc;
  if ((rc = d.in.ensure(n_in + 16)) != W3_OK) return rc;
  if ((rc = d.scratch.ensure(n_probs * 2 + 16)) != W3_OK) return rc;
  if ((rc = d.out.ensure(cap + 16)) != W3_OK) return rc;
  CK(cudaMemsetAsync(d.in.p, 0, n_in + 16, d.stream));
  CK(cudaMemcpyAsync(d.in.p, in, n_in, cudaMemcpyHostToDevice, d.stream));
  CK(cudaMemcpyAsync(d.scratch.p, probs, n_probs * 2, cudaMemcpyHostToDevice, d.stream));
  CK(w3_launch_ac_decode_probs((const uint8_t *)d.in.p, (uint32_t)n_in, (const uint16_t *)d.scratch.p,
                               (uint32_t)n_probs, (uint8_t *)d.out.p, d.stream));
  CK(cudaMemcpyAsync(out, d.out.p, (n_probs + 7) / 8, cudaMemcpyDeviceToHost, d.stream));
  CK(cudaStreamSynchronize(d.stream));
  return W3_OK;
}

int w3_selftest_state_table(w3_ctx *ctx, int which, uint16_t *next0, uint16_t *next1, uint16_t *prob, size_t n) {
  const size_t size = which == 0 ? W3_ST_SIZE : W3_ST_SUB;
  if (!ctx || (which != 0 && which != 1) || !next0 || !next1 || !prob || n > size) return W3_E_ARG;
  std::lock_guard<std::mutex> lock(ctx->mu);
  w3_dev &d = ctx->d[0];
  CK(cudaSetDevice(d.ordinal));
  int rc;
  if ((rc = d.scratch.ensure((W3_ST_SIZE + W3_ST_SUB) * sizeof(w3_state_entry))) != W3_OK) return rc;
  w3_state_entry *t = (w3_state_entry *)d.scratch.p, *aux = t + W3_ST_SIZE;
  CK(w3_launch_build_state_table(t, aux, d.stream));
  std::vector<w3_state_entry> h(size);
  CK(cudaMemcpyAsync(h.data(), which == 0 ? t : aux, size * sizeof(w3_state_entry), cudaMemcpyDeviceToHost, d.stream));
  CK(cudaStreamSynchronize(d.stream));
  for (size_t i = 0; i < n; i++) { next0[i] = h[i].next[0]; next1[i] = h[i].next[1]; prob[i] = h[i].prob; }
  return W3_OK;
}

int w3_selftest_hashslots(w3_ctx *ctx, const uint32_t *ops, size_t n_ops, uint8_t *cells, size_t n_cells,
                          uint32_t *results) {
  if (!ctx || !ops || !cells || !results || !n_cells || n_ops > 0x7FFFFFFF) return W3_E_ARG;
  for (size_t i = 0; i < n_ops; i++)
    if ((ops[4 * i] >> 8) >= n_cells || (ops[4 * i] & 0xFFu) > 6u) return W3_E_ARG;
  std::lock_guard<std::mutex> lock(ctx->mu);
  w3_dev &d = ctx->d[0];
  CK(cudaSetDevice(d.ordinal));
  int rc;
  if ((rc = d.in.ensure(n_ops * 16 + 16)) != W3_OK) return rc;
  if ((rc = d.scratch.ensure(n_cells * W3_CELL_BYTES)) != W3_OK) return rc;
  if ((rc = d.out.ensure(n_ops * 8 + 16)) != W3_OK) return rc;
  CK(cudaMemcpyAsync(d.in.p, ops, n_ops * 16, cudaMemcpyHostToDevice, d.stream));
  CK(cudaMemcpyAsync(d.scratch.p, cells, n_cells * W3_CELL_BYTES, cudaMemcpyHostToDevice, d.stream));
  CK(w3_launch_hashslot_ops((const uint32_t *)d.in.p, (uint32_t)n_ops, (uint8_t *)d.scratch.p, (uint32_t *)d.out.p, d.stream));
  CK(cudaMemcpyAsync(results, d.out.p, n_ops * 8, cudaMemcpyDeviceToHost, d.stream));
  CK(cudaMemcpyAsync(cells, d.scratch.p, n_cells * W3_CELL_BYTES, cudaMemcpyDeviceToHost, d.stream));
  CK(cudaStreamSynchronize(d.stream));
  return W3_OK;
}

} /* extern "C" */
