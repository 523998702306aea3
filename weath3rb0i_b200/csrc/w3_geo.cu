/*
 * w3_geo.cu — the model stack with its geometry as a run-time parameter
 * (SURVEY.md §8(f) ranks 2 and 3):
 *
 *     OrderNEntropy::new(bits_in_context, alignment_bits,
 *                        ACHistory::new(max_bits, StationaryModel::from_table(t)))
 *     (src/models/ordern_entropy.rs:15-24, src/history/ac_history.rs:16-18)
 *
 * init_model (src/main.rs:151) fixes (11, 3, 8), whose 8 KiB of counters live
 * in shared memory (w3_kernels.cu, w3_team.cuh, w3_lane.cuh).  The reference's
 * sweep binary (src/bin/entropy-hashing-ac/main.rs:20-37) walks ctx 8..30 x
 * align 0..4, and its own log says where that stack compresses best — ctx 20 /
 * align 3 on book1 (book1.log:76).  There the counter table is 4 MiB per
 * stream (4 GiB at ctx 30): it lives in HBM, every coded bit is one random
 * 32-byte-sector read-modify-write, and the gather roofline of SURVEY.md
 * §8(d).2 is what binds.
 *
 * Encode and the sweep are split in two kernels:
 *   w3_k_geo_hash    ACHistory::hash of EVERY bit position, one thread per
 *                    position (the literal loop of ac_history.rs:28-46, up to
 *                    64 nested coder steps: throughput work, no chain).  The
 *                    hash is a prefix code: with max_bits = M' < M it is the low
 *                    M' bits of the M-bit hash, so one pass at the largest
 *                    max_bits serves every point of a sweep.
 *   w3_k_geo_encode  one THREAD per stream (a block of a file, or a point of
 *                    the sweep over one file): context from the hash array,
 *                    counter read-modify-write in its own table in global
 *                    memory (prefetched into L2 W3_GEO_PF positions ahead —
 *                    the contexts are known in advance), Counter::p, coder
 *                    step, and either the byte writer (ACWriter) or the
 *                    bit-count sink of the sweep (ACStats — src/helpers.rs:60-90).
 * Decode (w3_k_geo_decode) is one thread per block with the literal hash on
 * its chain: correct for any geometry, not yet fast.
 */
#include <cuda_runtime.h>
#include <stdint.h>

#include "w3_core.cuh"
#include "w3_geo.h"
#include "w3_kernels.h"

#define W3_GEO_PF 24u /* positions the counter prefetch runs ahead */

/* ctx of position t from its hash — ordern_entropy.rs:36-45 (ctx_0 = 0: H[0] = 0) */
__device__ __forceinline__ uint32_t w3_geo_ctx(uint32_t h, uint32_t t, uint32_t ctx_bits, uint32_t align_bits) {
  const uint32_t mask = (1u << (ctx_bits - align_bits)) - 1u; /* ctx_bits <= 30 */
  return ((h & mask) << align_bits) | (t & ((1u << align_bits) - 1u));
}

/* H[first_bit + t] = ACHistory::hash after t bits of the block that bit belongs to; blocks of
 * block_size bytes restart the history (block_size a multiple of 4) */
/* raw = 1: H = the last 32 bits themselves — RawHistory::hash (src/history/raw_history.rs:18-20), the
 * context of OrderN (src/models/ordern.rs:36-45), for the `ordern` sweep binary */
__global__ void w3_k_geo_hash(const uint8_t *__restrict__ in, uint64_t n, uint32_t block_size,
                              const uint16_t *__restrict__ table, uint32_t max_bits, uint32_t raw, uint32_t *__restrict__ H) {
  __shared__ uint16_t tab[8];
  if (threadIdx.x < 8) tab[threadIdx.x] = table[threadIdx.x];
  __syncthreads();
  const uint64_t n_words = (n + 3) >> 2;
  const uint32_t lane = threadIdx.x & 31u;
  const uint64_t warps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
  const uint32_t wpb = block_size >> 2; /* words per block */
  for (uint64_t g = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; g < n_words; g += warps) {
    const uint32_t gw = (uint32_t)(g % wpb); /* word index inside its block */
    const uint32_t *src = reinterpret_cast<const uint32_t *>(in) + g;
    auto word = [&](int back) -> uint32_t { /* big-endian word `back` words before this one; zero before the block / past n */
      if ((int)gw < back) return 0u;
      uint32_t v = __byte_perm(__ldg(src - back), 0, 0x0123);
      const uint64_t at = (g - (uint64_t)back) * 4u;
      if (at + 4u > n) v &= 0xFFFFFFFFu << (8u * (uint32_t)(at + 4u - n));
      return v;
    };
    const uint32_t w2 = word(0), w1 = word(1), w0 = word(2);
    const uint32_t t = gw * 32u + lane; /* bits of the block before this position */
    const uint64_t bitpos = g * 32u + lane;
    if (bitpos >= n * 8u) continue;
    const uint32_t hist_hi = __funnelshift_l(w1, w0, lane), hist_lo = __funnelshift_l(w2, w1, lane);
    const uint64_t bits = ((uint64_t)hist_hi << 32) | hist_lo; /* ACHistory.bits: newest bit lowest */
    H[bitpos] = t == 0u ? 0u : raw ? hist_lo : w3_achash_exact(bits, t, tab, (int)max_bits).hash;
  }
}

/* one thread = one stream */
__global__ void w3_k_geo_encode(w3_geo_job *__restrict__ jobs, uint32_t n_jobs, uint32_t lanes) {
  const uint32_t lane = threadIdx.x & 31u;
  const uint32_t j = (blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * lanes + lane;
  if (lane >= lanes || j >= n_jobs) return;
  const w3_geo_job job = jobs[j];
  const uint32_t *H = job.H;
  uint32_t *cnt = job.counters;
  const uint32_t nbits = job.nbytes * 8u;
  uint32_t x1 = 0u, x2 = 0xFFFFFFFFu; /* new_coder — arithmetic_coder.rs:37-39 */
  uint64_t bit_count = 0, rev = 0;    /* ACStats — helpers.rs:60-69 */
  w3_writer wr;
  w3_writer_init(wr);
  uint32_t *dst = reinterpret_cast<uint32_t *>(job.out);
  const uint32_t cap_words = job.cap >> 2;
  uint32_t ovf = 0;
  auto store = [&](uint32_t idx, uint32_t v) {
    if (idx < cap_words) dst[idx] = __byte_perm(v, 0, 0x0123);
    else ovf = 1;
  };
  const bool count_only = job.out == nullptr;
  /* The contexts are known in advance (H), so the counter of position t + 2 is loaded while position t
   * is coded: two gathers in flight per stream.  A counter loaded early is stale if one of the two
   * positions in between used the same context — then the value they wrote is forwarded instead. */
  auto ctx_at = [&](uint32_t t) -> uint32_t { return w3_geo_ctx(__ldg(H + t), t, job.ctx_bits, job.align_bits); };
  uint32_t ctx0 = nbits > 0u ? ctx_at(0u) : 0u, ctx1 = nbits > 1u ? ctx_at(1u) : 0u; /* contexts of t, t + 1 */
  uint32_t c0 = cnt[ctx0], c1 = cnt[ctx1];                                            /* their counters as loaded */
  uint32_t pctx1 = 0xFFFFFFFFu, pnv1 = 0u, pctx2 = 0xFFFFFFFFu, pnv2 = 0u;            /* what t - 1 and t - 2 wrote */
  for (uint32_t i = 0; i < job.nbytes; i++) {
    const uint32_t byte = job.in[i];
#pragma unroll 1
    for (uint32_t kb = 0; kb < 8u; kb++) {
      const uint32_t t = i * 8u + kb;
      if (t + W3_GEO_PF < nbits) /* further ahead, pull the line into L2 */
        asm volatile("prefetch.global.L2 [%0];" ::"l"(cnt + ctx_at(t + W3_GEO_PF)));
      const uint32_t ctx2 = t + 2u < nbits ? ctx_at(t + 2u) : 0u;
      const uint32_t c2 = cnt[ctx2]; /* issued now, used two positions on */
      const uint32_t bit = (byte >> (7u - kb)) & 1u;
      const uint32_t c = ctx0 == pctx1 ? pnv1 : ctx0 == pctx2 ? pnv2 : c0;
      const uint32_t p = w3_counter_p1(c & 0xFFFFu, c >> 16); /* predict — main.rs:105 */
      const uint32_t nv = w3_counter_update(c, bit);          /* adapt — main.rs:106 */
      cnt[ctx0] = nv;
      pctx2 = pctx1; pnv2 = pnv1; pctx1 = ctx0; pnv1 = nv;
      ctx0 = ctx1; c0 = c1; ctx1 = ctx2; c1 = c2;
      w3_ac_narrow(x1, x2, bit, p << 16);                     /* encode — main.rs:107 */
      const uint32_t pre = x1;
      int s, k;
      w3_ac_renorm(x1, x2, s, k);
      if (count_only) { /* write_bit x s, inc_parity x k — helpers.rs:76-89 */
        if (s) { bit_count += (uint64_t)s + rev; rev = 0; }
        rev += (uint64_t)k;
      } else {
        w3_writer_step(wr, pre, s, k, store);
      }
    }
  }
  if (!count_only) { /* ac.flush — main.rs:111 (ACStats::flush counts nothing) */
    uint32_t tail, tail_bytes;
    const uint64_t len = w3_writer_flush(wr, x2, tail, tail_bytes, store);
    uint8_t *tb = job.out + (uint64_t)wr.words * 4u;
    for (uint32_t b = 0; b < tail_bytes; b++) {
      if ((uint64_t)wr.words * 4u + b < job.cap) tb[b] = (uint8_t)(tail >> (24u - 8u * b));
      else ovf = 1;
    }
    jobs[j].out_len = (uint32_t)len;
  }
  jobs[j].bit_count = bit_count;
  jobs[j].flags = ovf;
}

/* one thread = one block; job.in = the block's stream (out_len = its length), job.out = the bytes */
__global__ void w3_k_geo_decode(w3_geo_job *__restrict__ jobs, uint32_t n_jobs, uint32_t lanes,
                                const uint16_t *__restrict__ table, uint32_t max_bits) {
  __shared__ uint16_t tab[8];
  if (threadIdx.x < 8) tab[threadIdx.x] = table[threadIdx.x];
  __syncthreads();
  const uint32_t lane = threadIdx.x & 31u;
  const uint32_t j = (blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * lanes + lane;
  if (lane >= lanes || j >= n_jobs) return;
  const w3_geo_job job = jobs[j];
  uint32_t *cnt = job.counters;
  const uint8_t *sp = job.in;
  const uint64_t slen = job.out_len;
  uint64_t bp = 0;
  auto take = [&](int c) -> uint32_t { /* c <= 32 bits, MSB first, zero past the end — io.rs:18-48 */
    uint64_t acc = 0;
    const uint64_t by = bp >> 3;
    for (uint32_t b = 0; b < 5u; b++) acc = (acc << 8) | (by + b < slen ? sp[by + b] : 0u);
    const uint32_t v = c ? (uint32_t)((acc << (24 + (bp & 7u))) >> (64 - c)) : 0u;
    bp += (uint64_t)c;
    return v;
  };
  uint32_t x1 = 0u, x2 = 0xFFFFFFFFu, x = take(32);
  uint64_t hist = 0;
  uint32_t ctx = 0;
  for (uint32_t i = 0; i < job.nbytes; i++) {
    uint32_t byte = 0;
#pragma unroll 1
    for (uint32_t kb = 0; kb < 8u; kb++) {
      const uint32_t c = cnt[ctx];
      const uint32_t P = w3_counter_p1(c & 0xFFFFu, c >> 16) << 16;
      const uint32_t xmid = x1 + __umulhi(x2 - x1, P);
      const uint32_t bit = x <= xmid ? 1u : 0u; /* decode — arithmetic_coder.rs:80-89 */
      if (bit) x2 = xmid; else x1 = xmid + 1u;
      int s, k;
      const int nsh = w3_ac_renorm(x1, x2, s, k);
      if (nsh) x = ((x << nsh) | take(nsh)) ^ (k ? 0x80000000u : 0u);
      cnt[ctx] = w3_counter_update(c, bit);
      hist = (hist << 1) | bit;
      const uint32_t t = i * 8u + kb + 1u;
      ctx = w3_geo_ctx(w3_achash_exact(hist, t, tab, (int)max_bits).hash, t, job.ctx_bits, job.align_bits);
      byte = (byte << 1) | bit;
    }
    job.out[i] = (uint8_t)byte;
  }
}

extern "C" {
cudaError_t w3_launch_geo_hash(const uint8_t *d_in, uint64_t n, uint32_t block_size, const uint16_t *d_table,
                               uint32_t max_bits, uint32_t raw, uint32_t *d_H, int sm_count, cudaStream_t s) {
  if (!n) return cudaSuccess;
  w3_k_geo_hash<<<sm_count * 16, 256, 0, s>>>(d_in, n, block_size, d_table, max_bits, raw, d_H);
  return cudaGetLastError();
}
cudaError_t w3_launch_geo_encode(w3_geo_job *d_jobs, uint32_t n_jobs, uint32_t lanes, cudaStream_t s) {
  if (!n_jobs) return cudaSuccess;
  w3_k_geo_encode<<<(n_jobs + lanes - 1) / lanes, 32, 0, s>>>(d_jobs, n_jobs, lanes);
  return cudaGetLastError();
}
cudaError_t w3_launch_geo_decode(w3_geo_job *d_jobs, uint32_t n_jobs, uint32_t lanes, const uint16_t *d_table,
                                 uint32_t max_bits, cudaStream_t s) {
  if (!n_jobs) return cudaSuccess;
  w3_k_geo_decode<<<(n_jobs + lanes - 1) / lanes, 32, 0, s>>>(d_jobs, n_jobs, lanes, d_table, max_bits);
  return cudaGetLastError();
}
}
