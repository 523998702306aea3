/*
 * w3_kernels.cu — sm_100a kernels for weath3rb0i's predict→update→code path
 * (reference: src/main.rs:89-152 and everything it calls).
 *
 * One warp owns one block (an independent stream with fresh counters, history
 * and coder, exactly what the reference does for one file):
 *
 *   encode  lanes = 32 consecutive bit positions.  Each lane rebuilds its own
 *           64-bit history from the input words, finds its context through
 *           the shared-memory threshold table (w3_hash_lookup), resolves the
 *           counter it would have seen with __match_any_sync + popcounts,
 *           divides (Counter::p) — all 32 in parallel — and then the 32
 *           (bit, p) pairs go through the arithmetic coder in order.
 *   decode  strictly serial per stream (bit → history → context → p → bit).
 *           Lanes speculate: see w3_decode_blocks.
 *
 * Per-block adaptive state is 8 KiB of counters and lives in shared memory
 * together with one copy of the 50 KiB threshold table per CTA; HBM sees only
 * the input and output streams.
 */
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "w3_core.cuh"
#include "w3_kernels.h"

#define FULL 0xFFFFFFFFu

/* ------------------------------------------------------------------------ */
/* threshold-table construction (context creation)                          */

__global__ void w3_k_build_thr(w3_hashtab *tab) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < 8u * 256u) w3_build_thr(tab, i >> 8, i & 255u);
}
__global__ void w3_k_build_cell_rval(w3_hashtab *tab) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < 8u * W3_NCELL) w3_build_cell(tab, i / W3_NCELL, i % W3_NCELL);
  if (i < 8u * 256u) w3_build_rval(tab, i >> 8, i & 255u);
}
__global__ void w3_k_build_cell2(w3_hashtab *tab) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < 8u * W3_NCELL2) w3_build_cell2(tab, i / W3_NCELL2, i % W3_NCELL2);
}

/* ------------------------------------------------------------------------ */
/* shared-memory layout                                                     */

struct w3_smem_tab {
  uint64_t thr[8 * 256];
  uint16_t cell[8 * W3_NCELL];
  uint16_t table[8];
  uint32_t max_bits;
  uint32_t pad[3];
};
static_assert(sizeof(w3_smem_tab) % 16 == 0, "smem table must keep 16-byte alignment");

__device__ __forceinline__ void w3_load_tab(w3_smem_tab *s, const w3_hashtab *g) {
  const uint32_t tid = threadIdx.x, nt = blockDim.x;
  const uint4 *gt = reinterpret_cast<const uint4 *>(&g->thr[0][0]);
  uint4 *st = reinterpret_cast<uint4 *>(s->thr);
  for (uint32_t i = tid; i < sizeof(s->thr) / 16; i += nt) st[i] = gt[i];
  const uint4 *gc = reinterpret_cast<const uint4 *>(&g->cell[0][0]);
  uint4 *sc = reinterpret_cast<uint4 *>(s->cell);
  for (uint32_t i = tid; i < sizeof(s->cell) / 16; i += nt) sc[i] = gc[i];
  if (tid < 8) s->table[tid] = g->table[tid];
  if (tid == 0) s->max_bits = g->max_bits;
}

/* ACHistory::hash through the table, as a rank (see w3_rank_lookup); the
 * literal loop handles keys that sit exactly on a threshold */
__device__ __forceinline__ uint32_t w3_rank_dev(const w3_smem_tab *s, uint64_t key, uint32_t pos) {
  uint32_t r = w3_rank_lookup(key, pos & 7u, s->thr, s->cell);
  if (r == 0xFFFFFFFFu)
    r = w3_rank_of_hash(w3_achash_exact(w3_key_of_history(key), pos, s->table, (int)s->max_bits).hash);
  return r;
}

__device__ __forceinline__ uint32_t w3_bswap(uint32_t v) { return __byte_perm(v, 0, 0x0123); }

/* ------------------------------------------------------------------------ */
/* encode: one warp per block                                               */

__global__ void __launch_bounds__(W3_MAX_WARPS * 32)
w3_encode_blocks_warp(const uint8_t *__restrict__ in, uint64_t n_total, uint32_t block_size,
                 uint32_t n_blocks, uint8_t *__restrict__ scratch, uint64_t stride,
                 uint32_t *__restrict__ lens, uint32_t *__restrict__ overflow,
                 const w3_hashtab *__restrict__ tab) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  w3_smem_tab *st = reinterpret_cast<w3_smem_tab *>(smem_raw);
  const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31u;
  uint32_t *stats = reinterpret_cast<uint32_t *>(smem_raw + sizeof(w3_smem_tab)) + warp * W3_NCTX;

  w3_load_tab(st, tab);
  for (uint32_t i = lane; i < W3_NCTX; i += 32) stats[i] = 0u; /* ordern_entropy.rs:17 */
  __syncthreads();

  const uint32_t blk = blockIdx.x * (blockDim.x >> 5) + warp;
  if (blk >= n_blocks) return;
  const uint64_t off = (uint64_t)blk * block_size;
  const uint32_t n = (uint32_t)(n_total - off < block_size ? n_total - off : block_size);
  const uint32_t *src = reinterpret_cast<const uint32_t *>(in + off);
  uint32_t *dst = reinterpret_cast<uint32_t *>(scratch + (uint64_t)blk * stride);
  const uint32_t cap_words = (uint32_t)(stride >> 2);

  uint32_t ovf = 0;
  auto store = [&](uint32_t idx, uint32_t v) {
    if (lane == 0) {
      if (idx < cap_words) dst[idx] = w3_bswap(v);
      else ovf = 1;
    }
  };

  uint32_t x1 = 0, x2 = 0xFFFFFFFFu; /* new_coder — arithmetic_coder.rs:37-39 */
  w3_writer wr;
  w3_writer_init(wr);

  const uint32_t n_words = (n + 3u) >> 2;
  auto load_chunk = [&](uint32_t chunk) -> uint32_t {
    uint32_t widx = chunk * 32u + lane, v = 0;
    if (widx < n_words) {
      v = w3_bswap(__ldg(src + widx));
      uint32_t rem = n - widx * 4u; /* bytes of this word inside the block */
      if (rem < 4u) v &= 0xFFFFFFFFu << (8u * (4u - rem));
    }
    return v;
  };

  uint32_t w0 = 0, w1 = 0; /* the 64 stream bits before the current word */
  uint32_t cw_next = load_chunk(0);
  for (uint32_t chunk = 0; chunk * 32u < n_words; chunk++) {
    const uint32_t cw = cw_next;
    cw_next = load_chunk(chunk + 1);
    const uint32_t in_chunk = n_words - chunk * 32u < 32u ? n_words - chunk * 32u : 32u;
    for (uint32_t i = 0; i < in_chunk; i++) {
      const uint32_t w2 = __shfl_sync(FULL, cw, i);
      const uint32_t g = chunk * 32u + i;   /* word index = group of 32 bits */
      const uint32_t left = n * 8u - g * 32u; /* coded bits left in the block */
      const uint32_t nact = left < 32u ? left : 32u;
      const bool active = lane < nact;
      const uint32_t t = g * 32u + lane;    /* bits coded before this one */

      /* ACHistory.bits as seen by bit t: the 64 stream bits before it */
      const uint32_t hist_hi = __funnelshift_l(w1, w0, lane);
      const uint32_t hist_lo = __funnelshift_l(w2, w1, lane);
      const uint32_t bit = (w2 >> (31u - lane)) & 1u;
      const uint64_t key = ((uint64_t)__brev(hist_lo) << 32) | __brev(hist_hi);

      /* ctx = (hash << 3) | alignment — ordern_entropy.rs:43-44; ctx_0 = 0 (:20);
       * relabelled by rank, see W3_CTX0_RANKED in w3_core.cuh */
      uint32_t ctx = (w3_rank_dev(st, key, t) << W3_ALIGN_BITS) | (t & 7u);
      if (t == 0) ctx = W3_CTX0_RANKED;
      if (!active) ctx = 0x10000u | lane; /* unique: matches nobody */

      /* counter state each lane would see in sequential order */
      const uint32_t m = __match_any_sync(FULL, ctx);
      const uint32_t bm = __ballot_sync(FULL, active && bit);
      const uint32_t slot = active ? stats[ctx] : 0u;
      const uint32_t lower = m & ((1u << lane) - 1u);
      const uint32_t e1 = __popc(lower & bm), e0 = __popc(lower) - e1;
      const uint32_t t1 = __popc(m & bm), t0 = __popc(m) - t1;
      const uint32_t c0 = slot & 0xFFFFu, c1 = slot >> 16;
      const bool risky = active && (c0 + t0 >= 0xFFFFu || c1 + t1 >= 0xFFFFu);
      uint32_t p;
      __syncwarp();
      if (__any_sync(FULL, risky)) {
        /* a halving (counter.rs:22-25) may fall inside this group: replay it
         * in order, one lane at a time */
        p = 0;
        for (uint32_t j = 0; j < nact; j++) {
          if (lane == j) {
            uint32_t sl = stats[ctx];
            p = w3_counter_p(sl & 0xFFFFu, sl >> 16);
            stats[ctx] = w3_counter_update(sl, bit);
          }
          __syncwarp();
        }
      } else {
        p = w3_counter_p(c0 + e0, c1 + e1);                /* predict — main.rs:105 */
        if (active && lane == 31u - (uint32_t)__clz((int)m)) /* last of its group */
          stats[ctx] = (c0 + t0) | ((c1 + t1) << 16);      /* adapt — main.rs:106 */
      }
      __syncwarp();

      /* ac.encode(bit, p) in order — main.rs:107 */
      const uint32_t P = p << 16; /* p in [1,65535], so lerp's prob==0 arm is dead */
      for (uint32_t j = 0; j < nact; j++) {
        const uint32_t Pj = __shfl_sync(FULL, P, j);
        w3_ac_narrow(x1, x2, (bm >> j) & 1u, Pj);
        const uint32_t pre = x1;
        int s, k;
        w3_ac_renorm(x1, x2, s, k);
        w3_writer_step(wr, pre, s, k, store);
      }
      w0 = w1;
      w1 = w2;
    }
  }

  /* ac.flush — main.rs:111 */
  uint32_t tail, tail_bytes;
  const uint64_t len = w3_writer_flush(wr, x2, tail, tail_bytes, store);
  if (lane == 0) {
    uint8_t *tb = reinterpret_cast<uint8_t *>(dst) + (uint64_t)wr.words * 4u;
    for (uint32_t b = 0; b < tail_bytes; b++) {
      if ((uint64_t)wr.words * 4u + b < stride) tb[b] = (uint8_t)(tail >> (24u - 8u * b));
      else ovf = 1;
    }
    lens[blk] = (uint32_t)len;
    if (ovf) atomicOr(overflow, 1u);
  }
}

/* ------------------------------------------------------------------------ */
/* decode: one warp per block                                               */

/* MSB-first bit source over a byte range at any alignment; zero bits past the
 * end (ACReader — src/entropy_coding/io.rs:18-48).  Loads run two words ahead
 * of use so their latency stays off the decode chain. */
struct w3_reader {
  const uint32_t *ab; /* aligned base */
  uint32_t sel;       /* __byte_perm selector for the start misalignment */
  int64_t slen;       /* stream length in bytes */
  uint32_t widx;      /* next stream word to assemble */
  uint32_t cur, nxt;  /* ab[widx], ab[widx+1] */
  uint64_t win;       /* upcoming bits, left aligned */
  int have;
};

__device__ __forceinline__ uint32_t w3_reader_ld(const w3_reader &r, uint32_t i) {
  /* aligned word i of the range, 0 when it lies wholly past the stream's end
   * (the caller guarantees 8 bytes of slack after the buffer) */
  return ((int64_t)i * 4 < r.slen + 4) ? __ldg(r.ab + i) : 0u;
}
__device__ __forceinline__ uint32_t w3_reader_word(w3_reader &r) {
  uint32_t v = __byte_perm(r.cur, r.nxt, r.sel);
  int64_t rem = r.slen - (int64_t)r.widx * 4;
  if (rem < 4) v = rem <= 0 ? 0u : (v & (0xFFFFFFFFu << (8 * (4 - (int)rem))));
  r.widx++;
  r.cur = r.nxt;
  r.nxt = w3_reader_ld(r, r.widx + 1);
  return v;
}
__device__ __forceinline__ void w3_reader_init(w3_reader &r, const uint8_t *p, uint64_t slen) {
  uint32_t sh = (uint32_t)(reinterpret_cast<uintptr_t>(p) & 3u);
  r.ab = reinterpret_cast<const uint32_t *>(p - sh);
  r.sel = ((sh + 3u) | ((sh + 2u) << 4) | ((sh + 1u) << 8) | (sh << 12));
  r.slen = (int64_t)slen;
  r.widx = 0;
  r.cur = w3_reader_ld(r, 0);
  r.nxt = w3_reader_ld(r, 1);
  uint32_t a = w3_reader_word(r), b = w3_reader_word(r);
  r.win = ((uint64_t)a << 32) | b;
  r.have = 64;
}
__device__ __forceinline__ uint32_t w3_reader_take(w3_reader &r, int c) { /* c in [1,32] */
  uint32_t v = (uint32_t)(r.win >> (64 - c));
  r.win = c == 32 ? r.win << 32 : r.win << c;
  r.have -= c;
  if (r.have <= 32) {
    r.win |= (uint64_t)w3_reader_word(r) << (32 - r.have);
    r.have += 32;
  }
  return v;
}

/* per-warp adaptive state of the dense decode kernel: probabilities cached
 * next to the counts so that predict is one load (Counter::p is recomputed by
 * the update, off the decode chain) */
struct w3_dense_dec {
  uint32_t probs[W3_NCTX];  /* Counter::p << 16 */
  uint32_t counts[W3_NCTX]; /* c0 | c1 << 16 */
};
#define W3_DENSE_THI_ROW 264u

/* fixed-latency rank lookup on the W3_PB table: the next six thresholds' high
 * words are loaded at once and compared as a tree; declines (0xFFFFFFFF) when
 * the cell holds more or the high word ties, and the general search runs */
__device__ __forceinline__ uint32_t w3_rank_fixed_nat(uint32_t khi, uint32_t a, const w3_smem_tab *st,
                                                      const uint32_t *thi) {
  const uint32_t e = st->cell[a * W3_NCELL + (khi >> (32 - W3_PB))];
  const uint32_t base = e & 0xFFu;
  const uint32_t *t = thi + a * W3_DENSE_THI_ROW + base;
  uint32_t b[6];
#pragma unroll
  for (int i = 0; i < 6; i++) b[i] = t[i];
  bool bad = e > ((6u << 8) | 0xFFu);
  uint32_t g[6];
#pragma unroll
  for (int i = 0; i < 6; i++) {
    g[i] = khi > b[i] ? 1u : 0u;
    bad |= khi == b[i];
  }
  const uint32_t r = base + ((g[0] + g[1] + g[2]) + (g[3] + g[4] + g[5]));
  return bad ? 0xFFFFFFFFu : r;
}

/* decode, many-block regime: one warp per block, W3_MAX_WARPS blocks per CTA
 * sharing one table copy.  Lanes 0 and 1 evaluate the context and probability
 * of the NEXT bit for both values of the bit being decoded, in the same
 * instruction stream as the decode step itself, so the table lookup overlaps
 * the coder arithmetic instead of following it; a shuffle picks the survivor. */
__global__ void __launch_bounds__(W3_MAX_WARPS * 32)
w3_decode_blocks_warp(const uint8_t *__restrict__ comp, const uint64_t *__restrict__ offsets,
                 uint32_t n_blocks, uint32_t block_size, uint64_t n_total,
                 uint8_t *__restrict__ out, const w3_hashtab *__restrict__ tab) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  w3_smem_tab *st = reinterpret_cast<w3_smem_tab *>(smem_raw);
  uint32_t *thi = reinterpret_cast<uint32_t *>(smem_raw + sizeof(w3_smem_tab));
  const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31u;
  w3_dense_dec *ds = reinterpret_cast<w3_dense_dec *>(smem_raw + sizeof(w3_smem_tab) + 8u * W3_DENSE_THI_ROW * 4u) + warp;

  w3_load_tab(st, tab);
  for (uint32_t i = threadIdx.x; i < 8u * W3_DENSE_THI_ROW; i += blockDim.x) {
    const uint32_t row = i / W3_DENSE_THI_ROW, col = i % W3_DENSE_THI_ROW;
    thi[i] = col < 255u ? (uint32_t)(__ldg(&tab->thr[row][col]) >> 32) : 0xFFFFFFFFu;
  }
  for (uint32_t i = lane; i < W3_NCTX; i += 32) { ds->probs[i] = 0x80000000u; ds->counts[i] = 0u; }
  __syncthreads();

  const uint32_t blk = blockIdx.x * (blockDim.x >> 5) + warp;
  if (blk >= n_blocks) return;
  const uint64_t off = (uint64_t)blk * block_size;
  const uint32_t n = (uint32_t)(n_total - off < block_size ? n_total - off : block_size);
  uint8_t *dst = out + off;

  w3_reader rd;
  w3_reader_init(rd, comp + offsets[blk], offsets[blk + 1] - offsets[blk]);
  uint32_t x1 = 0, x2 = 0xFFFFFFFFu;
  uint32_t x = w3_reader_take(rd, 32); /* new_decoder — arithmetic_coder.rs:75-78 */
  uint64_t key = 0;                    /* ACHistory.bits, newest bit at bit 63 */
  uint32_t ctx = W3_CTX0_RANKED, P = 0x80000000u, t = 0, word = 0;
  const uint32_t guess = lane & 1u;

  for (uint32_t i = 0; i < n; i++) {
    uint32_t byte = 0;
#pragma unroll 2
    for (int kb = 0; kb < 8; kb++) {
      /* this lane's guess of the bit being decoded -> context and P of the next one */
      const uint64_t gkey = (key >> 1) | ((uint64_t)guess << 63);
      uint32_t gr = w3_rank_fixed_nat((uint32_t)(gkey >> 32), (t + 1u) & 7u, st, thi);
      if (gr == 0xFFFFFFFFu) gr = w3_rank_dev(st, gkey, t + 1u);
      const uint32_t gctx = (gr << W3_ALIGN_BITS) | ((t + 1u) & 7u);
      const uint32_t gP = ds->probs[gctx];
      /* decode bit t (main.rs:134-135) */
      const uint32_t xmid = x1 + __umulhi(x2 - x1, P);
      const uint32_t bit = x <= xmid ? 1u : 0u;
      if (bit) x2 = xmid; else x1 = xmid + 1u;
      int s, k;
      const int nsh = w3_ac_renorm(x1, x2, s, k);
      if (nsh) x = ((x << nsh) | w3_reader_take(rd, nsh)) ^ ((0u - (uint32_t)k) & 0x80000000u);
      /* adapt counter of bit t (main.rs:136); every lane writes the same words */
      const uint32_t nv = w3_counter_update(ds->counts[ctx], bit);
      ds->counts[ctx] = nv;
      ds->probs[ctx] = w3_counter_p(nv & 0xFFFFu, nv >> 16) << 16;
      /* the guess that came true */
      ctx = __shfl_sync(FULL, gctx, bit);
      P = __shfl_sync(FULL, gP, bit);
      key = (key >> 1) | ((uint64_t)bit << 63);
      t++;
      byte = (byte << 1) | bit;
    }
    word |= byte << (8u * (i & 3u));
    if ((i & 3u) == 3u) {
      if (lane == 0) *reinterpret_cast<uint32_t *>(dst + (i - 3u)) = word;
      word = 0;
    }
  }
  if (lane == 0)
    for (uint32_t b = 0; b < (n & 3u); b++) dst[(n & ~3u) + b] = (uint8_t)(word >> (8u * b));
}

#include "w3_team.cuh"
#include "w3_lane.cuh"

/* ------------------------------------------------------------------------ */
/* stream compaction: offsets = exclusive scan of lens, then a packed copy   */

__global__ void w3_k_scan_lens(const uint32_t *__restrict__ lens, uint32_t n_blocks,
                               uint64_t *__restrict__ offsets) {
  __shared__ uint64_t part[1024];
  const uint32_t tid = threadIdx.x, nt = blockDim.x;
  const uint32_t per = (n_blocks + nt - 1) / nt;
  const uint32_t lo = tid * per, hi = lo + per < n_blocks ? lo + per : n_blocks;
  uint64_t sum = 0;
  for (uint32_t i = lo; i < hi; i++) sum += lens[i];
  part[tid] = sum;
  __syncthreads();
  if (tid == 0) {
    uint64_t run = 0;
    for (uint32_t i = 0; i < nt; i++) { uint64_t v = part[i]; part[i] = run; run += v; }
    offsets[n_blocks] = run;
  }
  __syncthreads();
  uint64_t run = part[tid];
  for (uint32_t i = lo; i < hi; i++) { offsets[i] = run; run += lens[i]; }
}

/* alt (optional): per block, the address of its stream if it is not in its scratch region
 * (blocks re-encoded after an overflow), 0 otherwise */
__global__ void w3_k_pack(const uint8_t *__restrict__ scratch, uint64_t stride,
                          const uint64_t *__restrict__ offsets, const uint64_t *__restrict__ alt,
                          uint8_t *__restrict__ packed) {
  const uint32_t blk = blockIdx.x;
  const uint64_t len = offsets[blk + 1] - offsets[blk];
  const uint8_t *src = scratch + (uint64_t)blk * stride;
  if (alt && alt[blk]) src = reinterpret_cast<const uint8_t *>(alt[blk]);
  uint8_t *dst = packed + offsets[blk];
  for (uint64_t i = (uint64_t)blockIdx.y * blockDim.x + threadIdx.x; i < len;
       i += (uint64_t)gridDim.y * blockDim.x)
    dst[i] = src[i];
}

/* ------------------------------------------------------------------------ */
/* StationaryModel::new — src/models/ac_hash/stationary.rs:14-34             */
/* The 8 bit-position counters of a buffer.  Counter::update halves both     */
/* counts whenever one reaches 65535 (counter.rs:22-25), so this is not a    */
/* histogram: the halving points are found exactly.  Pass 1 counts the ones  */
/* of every bit position per 1 KiB chunk (any order); pass 2 walks the       */
/* chunks in order, one warp per bit position, 32 chunks per trip while no   */
/* halving can fall inside them and byte by byte where one does.             */

#define W3_TRAIN_CHUNK 1024u

__global__ void w3_k_bitpos_counts(const uint8_t *__restrict__ in, uint64_t n,
                                   uint16_t *__restrict__ cnt /*[chunks][8]*/) {
  const uint32_t lane = threadIdx.x & 31u;
  const uint64_t n_chunks = (n + W3_TRAIN_CHUNK - 1) / W3_TRAIN_CHUNK;
  const uint64_t warps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
  for (uint64_t chunk = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; chunk < n_chunks; chunk += warps) {
    const uint64_t base = chunk * W3_TRAIN_CHUNK + lane * 32u;
    uint32_t acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (uint32_t wd = 0; wd < 8; wd++) {
      const uint64_t at = base + wd * 4u;
      uint32_t v = 0;
      if (at + 4 <= n) v = __ldg(reinterpret_cast<const uint32_t *>(in + at));
      else for (uint32_t b = 0; b < 4 && at + b < n; b++) v |= (uint32_t)in[at + b] << (8u * b);
#pragma unroll
      for (int j = 0; j < 8; j++) acc[j] += __popc(v & (0x80808080u >> j)); /* j = 0: a byte's MSB */
    }
#pragma unroll
    for (int j = 0; j < 8; j++)
      for (uint32_t d = 16; d; d >>= 1) acc[j] += __shfl_xor_sync(FULL, acc[j], d);
    if (lane < 8) {
      uint32_t mine = acc[0];
#pragma unroll
      for (int j = 1; j < 8; j++) mine = lane == (uint32_t)j ? acc[j] : mine;
      cnt[chunk * 8u + lane] = (uint16_t)mine;
    }
  }
}

__global__ void __launch_bounds__(256) w3_k_stationary_scan(const uint8_t *__restrict__ in, uint64_t n,
                                                            const uint16_t *__restrict__ cnt,
                                                            uint16_t *__restrict__ table /*[8]*/) {
  const uint32_t j = threadIdx.x >> 5, lane = threadIdx.x & 31u; /* warp j = bit position j */
  const uint64_t n_chunks = (n + W3_TRAIN_CHUNK - 1) / W3_TRAIN_CHUNK;
  uint32_t c0 = 0, c1 = 0; /* Counter { data: [c0, c1] } — counter.rs:4-11 */
  auto update = [&](uint32_t bit) { /* Counter::update — counter.rs:20-26 */
    uint32_t slot = w3_counter_update(c0 | (c1 << 16), bit);
    c0 = slot & 0xFFFFu;
    c1 = slot >> 16;
  };
  uint64_t cursor = 0;
  while (cursor < n_chunks) {
    const uint64_t idx = cursor + lane;
    const bool valid = idx < n_chunks;
    const uint32_t len = valid ? (uint32_t)(n - idx * W3_TRAIN_CHUNK < W3_TRAIN_CHUNK ? n - idx * W3_TRAIN_CHUNK : W3_TRAIN_CHUNK) : 0u;
    const uint32_t ones = valid ? cnt[idx * 8u + j] : 0u;
    uint32_t po = ones, pz = len - ones; /* inclusive prefix sums over the 32 chunks */
    for (uint32_t d = 1; d < 32; d <<= 1) {
      const uint32_t uo = __shfl_up_sync(FULL, po, d), uz = __shfl_up_sync(FULL, pz, d);
      if (lane >= d) { po += uo; pz += uz; }
    }
    const uint32_t hit = __ballot_sync(FULL, valid && (c1 + po >= 0xFFFFu || c0 + pz >= 0xFFFFu));
    if (!hit) { /* no count can reach 65535 inside these chunks: plain sums */
      c1 += __shfl_sync(FULL, po, 31);
      c0 += __shfl_sync(FULL, pz, 31);
      cursor += 32;
      continue;
    }
    const uint32_t f = (uint32_t)__ffs((int)hit) - 1u; /* first chunk a halving may fall in */
    if (f) { c1 += __shfl_sync(FULL, po, f - 1u); c0 += __shfl_sync(FULL, pz, f - 1u); }
    const uint64_t chunk = cursor + f, base = chunk * W3_TRAIN_CHUNK;
    const uint32_t clen = (uint32_t)(n - base < W3_TRAIN_CHUNK ? n - base : W3_TRAIN_CHUNK);
    for (uint32_t r = 0; r * 32u < clen; r++) { /* 32 bytes per trip, lane = byte */
      const uint32_t o = r * 32u + lane;
      const bool act = o < clen;
      const uint32_t bit = act ? ((uint32_t)in[base + o] >> (7u - j)) & 1u : 0u;
      const uint32_t ma = __ballot_sync(FULL, act), mb = __ballot_sync(FULL, act && bit);
      const uint32_t so = __popc(mb), sz = __popc(ma) - so;
      if (c1 + so >= 0xFFFFu || c0 + sz >= 0xFFFFu) {
        for (uint32_t k = 0; k < 32; k++) if ((ma >> k) & 1u) update((mb >> k) & 1u);
      } else {
        c1 += so;
        c0 += sz;
      }
    }
    cursor = chunk + 1;
  }
  if (lane == 0) table[j] = (uint16_t)w3_counter_p(c0, c1); /* stationary.rs:23-32 */
}

/* ------------------------------------------------------------------------ */
/* self-test kernels (called through w3_selftest_* in the C ABI by tests/)   */

__global__ void w3_k_hash_array(const uint64_t *bits, const uint32_t *pos, uint32_t n,
                                uint32_t *out, const w3_hashtab *tab) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  w3_smem_tab *st = reinterpret_cast<w3_smem_tab *>(smem_raw);
  w3_load_tab(st, tab);
  __syncthreads();
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
    out[i] = w3_hash_of_rank(w3_rank_dev(st, w3_key_of_history(bits[i]), pos[i]));
}

/* exhaustive Counter::p check: float estimate + fix-up vs the u64 statement */
__global__ void w3_k_counter_p_check(unsigned long long *mismatches) {
  const uint32_t c0 = blockIdx.x; /* 0..65534 */
  unsigned long long bad = 0;
  for (uint32_t c1 = threadIdx.x; c1 <= 65534u; c1 += blockDim.x) {
    bad += w3_counter_p(c0, c1) != w3_counter_p_u64(c0, c1);
    bad += w3_counter_p1(c0, c1) != w3_counter_p_u64(c0, c1); /* the lane decoder's single-correction form */
    /* the speculators' pair form: both updates of this state */
    const uint32_t slot = c0 | (c1 << 16);
    uint32_t p0, p1;
    w3_counter_p_both(slot, p0, p1);
    const uint32_t n0 = w3_counter_update(slot, 0u), n1 = w3_counter_update(slot, 1u);
    bad += p0 != w3_counter_p_u64(n0 & 0xFFFFu, n0 >> 16);
    bad += p1 != w3_counter_p_u64(n1 & 0xFFFFu, n1 >> 16);
  }
  if (bad) atomicAdd(mismatches, bad);
}

/* Random 32-byte-sector gather probe: the denominator of the gather roofline
 * (SURVEY.md §8(d)): every thread loads `per_thread` independent sectors at
 * hashed indices of a buffer much larger than L2 and folds them into a sink. */
__global__ void w3_k_gather_probe(const uint4 *__restrict__ buf, uint64_t n_sectors, uint32_t per_thread,
                                  unsigned long long *sink) {
  uint64_t x = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) * 0x9E3779B97F4A7C15ull + 0x632BE59BD9B4E019ull;
  uint32_t acc = 0;
#pragma unroll 8 /* eight independent sectors in flight per thread */
  for (uint32_t i = 0; i < per_thread; i++) {
    x ^= x >> 29; x *= 0xBF58476D1CE4E5B9ull; x ^= x >> 32; /* independent of the loaded data: loads overlap */
    const uint64_t sct = ((x >> 32) * (n_sectors & 0xFFFFFFFFull)) >> 32; /* < n_sectors <= 2^32 - 1 (capped by the caller) */
    const uint4 a = __ldg(buf + 2 * sct), b = __ldg(buf + 2 * sct + 1);
    acc ^= a.x ^ a.y ^ a.z ^ a.w ^ b.x ^ b.y ^ b.z ^ b.w;
  }
  if (acc == 0x12345677u) atomicAdd(sink, 1ull); /* keeps the loads alive */
}

/* the coder alone with caller-supplied probabilities (the reference's KAT
 * harness, arithmetic_coder_tests.rs:8-42), one thread.  Handles prob == 0 /
 * x1 == x2 (32 shifts), which the codec kernels never meet. */
__global__ void w3_k_ac_encode_probs(const uint8_t *in, const uint16_t *probs, uint32_t n_probs,
                                     uint8_t *out, uint32_t cap, uint32_t *out_len) {
  if (threadIdx.x || blockIdx.x) return;
  uint32_t x1 = 0, x2 = 0xFFFFFFFFu;
  w3_writer wr;
  w3_writer_init(wr);
  auto store = [&](uint32_t idx, uint32_t v) {
    for (int b = 0; b < 4; b++)
      if (idx * 4u + b < cap) out[idx * 4u + b] = (uint8_t)(v >> (24 - 8 * b));
  };
  for (uint32_t i = 0; i < n_probs; i++) {
    uint32_t bit = (in[i >> 3] >> (7u - (i & 7u))) & 1u;
    w3_ac_narrow(x1, x2, bit, w3_prob32(probs[i]));
    uint32_t pre = x1;
    int s, k;
    if (x1 == x2) { s = 32; k = 0; x1 = 0; x2 = 0xFFFFFFFFu; }
    else w3_ac_renorm(x1, x2, s, k);
    w3_writer_step(wr, pre, s, k, store);
  }
  uint32_t tail, tb;
  uint64_t len = w3_writer_flush(wr, x2, tail, tb, store);
  for (uint32_t b = 0; b < tb; b++)
    if (wr.words * 4u + b < cap) out[wr.words * 4u + b] = (uint8_t)(tail >> (24 - 8 * b));
  *out_len = (uint32_t)len;
}

__global__ void w3_k_ac_decode_probs(const uint8_t *in, uint32_t n_in, const uint16_t *probs,
                                     uint32_t n_probs, uint8_t *out) {
  if (threadIdx.x || blockIdx.x) return;
  w3_reader rd;
  w3_reader_init(rd, in, n_in);
  uint32_t x1 = 0, x2 = 0xFFFFFFFFu, x = w3_reader_take(rd, 32);
  for (uint32_t i = 0; i < n_probs; i++) {
    uint32_t xmid = x1 + __umulhi(x2 - x1, w3_prob32(probs[i]));
    uint32_t bit = x <= xmid ? 1u : 0u;
    if (bit) x2 = xmid; else x1 = xmid + 1u;
    if (x1 == x2) { x1 = 0; x2 = 0xFFFFFFFFu; x = w3_reader_take(rd, 32); }
    else {
      int s, k;
      int nsh = w3_ac_renorm(x1, x2, s, k);
      if (nsh) x = ((x << nsh) | w3_reader_take(rd, nsh)) ^ (k ? 0x80000000u : 0u);
    }
    if ((i & 7u) == 0) out[i >> 3] = 0;
    out[i >> 3] |= (uint8_t)(bit << (7u - (i & 7u)));
  }
}

/* ------------------------------------------------------------------------ */
/* launchers                                                                */

static inline size_t w3_smem_bytes(int warps) {
  return sizeof(w3_smem_tab) + (size_t)warps * W3_NCTX * sizeof(uint32_t);
}
static inline size_t w3_dense_dec_smem_bytes(int warps) {
  return sizeof(w3_smem_tab) + 8u * W3_DENSE_THI_ROW * 4u + (size_t)warps * sizeof(w3_dense_dec);
}

/* warps per CTA: spread blocks over all 148 SMs first, then stack warps */
static inline int w3_pick_warps(uint32_t n_blocks, int sm_count) {
  int w = (int)((n_blocks + (uint32_t)sm_count - 1) / (uint32_t)sm_count);
  if (w < 1) w = 1;
  if (w > W3_MAX_WARPS) w = W3_MAX_WARPS;
  return w;
}

/* Kernel families (w3_launch_opts.kernels forces one; 0 = by block count):
 *   team   one CTA per block, the coder's chain in a warp of its own (w3_team.cuh): up to one
 *          block per SM nothing is faster (69 / 109 cycles per coded bit, encode / decode).
 *   many   more blocks than SMs.  Encode: the same teams, up to eight per CTA sharing one table
 *          copy, so that every block is resident at once.  Decode: one LANE per block
 *          (w3_lane.cuh) — an instruction costs the same for 1 or 32 active lanes.
 *   warp   round 1's one-warp-per-block kernels, kept for comparison. */
enum { W3_K_AUTO = 0, W3_K_TEAM = 1, W3_K_WARP = 2, W3_K_MANY = 3 };
enum { W3_DEC_TEAM = 1, W3_DEC_WARP = 2, W3_DEC_LANE = 3, W3_DEC_PAIR = 4, W3_DEC_PAIR_P = 5, W3_DEC_SPLIT = 6 }; /* w3_decode_family */

static inline uint32_t w3_ceil_div(uint32_t a, uint32_t b) { return (a + b - 1u) / b; }

/* blocks a CTA takes at a time when every SM gets ceil(n_blocks / sm_count) blocks in as few,
 * equally filled rounds as `cap` (what fits a CTA) allows */
static inline uint32_t w3_per_cta(uint32_t n_blocks, int sm_count, uint32_t cap) {
  const uint32_t q = w3_ceil_div(n_blocks, (uint32_t)sm_count);
  const uint32_t rounds = w3_ceil_div(q, cap);
  return w3_ceil_div(q, rounds);
}

#define W3_LANE_SMEM_MAX (227u * 1024u)
/* Measured on 1 GiB (profiles/r02_many_block_kernels.txt): the pair form of the lane decoder takes
 * ~330 cycles per coded bit of every resident block whatever their number; the team decoder
 * 109 cycles per bit with one block per SM, in as many waves as it takes.  Two waves of teams
 * (218) still beat the lanes; from the third on the lanes win.  The pair form beat the lane form
 * at every block count tried (1024 x 1 MiB: 766 against 668 MB/s; 4096 x 256 KiB: 1476 against
 * 1312), so the lane form is only used when asked for.  The split form (coder and models in different
 * warps, up to 8 blocks per SM) takes ~236 cycles per bit: it takes over from the third wave of teams up
 * to 8 blocks per SM, the pair form beyond. */
#define W3_TEAM_MAX_WAVES 2u
static inline uint32_t w3_lane_cap(void) { return (W3_LANE_SMEM_MAX - (uint32_t)sizeof(w3_lane_tab)) / (W3_NCTX * 4u); }

extern "C" {

cudaError_t w3_launch_build_table(w3_hashtab *d_tab, cudaStream_t s) {
  w3_k_build_thr<<<8, 256, 0, s>>>(d_tab);
  w3_k_build_cell_rval<<<(8 * W3_NCELL + 255) / 256, 256, 0, s>>>(d_tab);
  w3_k_build_cell2<<<(8 * W3_NCELL2 + 255) / 256, 256, 0, s>>>(d_tab);
  return cudaGetLastError();
}

cudaError_t w3_kernels_init(void) {
  cudaError_t e;
  size_t mx = w3_smem_bytes(W3_MAX_WARPS);
  e = cudaFuncSetAttribute(w3_encode_blocks_warp, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)mx);
  if (e != cudaSuccess) return e;
  e = cudaFuncSetAttribute(w3_decode_blocks_warp, cudaFuncAttributeMaxDynamicSharedMemorySize,
                           (int)w3_dense_dec_smem_bytes(W3_MAX_WARPS));
  if (e != cudaSuccess) return e;
  e = cudaFuncSetAttribute(w3_encode_blocks_team<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                           (int)(sizeof(w3_smem_tab) + W3_ENC_TEAM_MAXB * sizeof(w3_enc_team)));
  if (e != cudaSuccess) return e;
  e = cudaFuncSetAttribute(w3_encode_blocks_team<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                           (int)(sizeof(w3_smem_tab) + W3_ENC_TEAM_MAXB * sizeof(w3_enc_team)));
  if (e != cudaSuccess) return e;
  e = cudaFuncSetAttribute(w3_decode_blocks_team, cudaFuncAttributeMaxDynamicSharedMemorySize,
                           (int)(sizeof(w3_dec_tab) + sizeof(w3_dec_team)));
  if (e != cudaSuccess) return e;
  e = cudaFuncSetAttribute(w3_decode_blocks_lane, cudaFuncAttributeMaxDynamicSharedMemorySize,
                           (int)(sizeof(w3_lane_tab) + w3_lane_cap() * W3_NCTX * 4u));
  if (e != cudaSuccess) return e;
  e = cudaFuncSetAttribute(w3_decode_blocks_pair<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                           (int)(sizeof(w3_lane_tab) + w3_lane_cap() * W3_NCTX * 4u));
  if (e != cudaSuccess) return e;
  e = cudaFuncSetAttribute(w3_decode_blocks_pair<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                           (int)(sizeof(w3_lane_tab) + (w3_lane_cap() / 2u) * W3_NCTX * 8u));
  if (e != cudaSuccess) return e;
  e = cudaFuncSetAttribute(w3_decode_blocks_split, cudaFuncAttributeMaxDynamicSharedMemorySize,
                           (int)(sizeof(w3_lane_tab) + sizeof(w3_split_ring) + W3_SPLIT_MAXB * W3_SPLIT_WORDS * 4u));
  if (e != cudaSuccess) return e;
  return cudaFuncSetAttribute(w3_k_hash_array, cudaFuncAttributeMaxDynamicSharedMemorySize,
                              (int)sizeof(w3_smem_tab));
}

cudaError_t w3_launch_encode(const uint8_t *d_in, uint64_t n_total, uint32_t block_size,
                             uint32_t n_blocks, uint8_t *d_scratch, uint64_t stride,
                             uint32_t *d_lens, uint32_t *d_overflow, const w3_hashtab *d_tab,
                             int sm_count, const w3_launch_opts *opts, cudaStream_t s) {
  if (!n_blocks) return cudaSuccess;
  const int k = opts ? opts->kernels : W3_K_AUTO;
  if (k == W3_K_WARP) {
    int warps = w3_pick_warps(n_blocks, sm_count);
    uint32_t grid = (n_blocks + (uint32_t)warps - 1) / (uint32_t)warps;
    w3_encode_blocks_warp<<<grid, warps * 32, w3_smem_bytes(warps), s>>>(
        d_in, n_total, block_size, n_blocks, d_scratch, stride, d_lens, d_overflow, d_tab);
    return cudaGetLastError();
  }
  uint32_t teams = 1u;
  if (k != W3_K_TEAM) teams = w3_per_cta(n_blocks, sm_count, W3_ENC_TEAM_MAXB);
  if (opts && opts->enc_teams > 0) teams = (uint32_t)opts->enc_teams < W3_ENC_TEAM_MAXB ? (uint32_t)opts->enc_teams : W3_ENC_TEAM_MAXB;
  /* from three teams per CTA on, one coder warp serves them all (lane b = team b); measured on 1 GiB
   * (profiles/r02_encode_lane_coder.txt): 7 teams 402 against 579 ms, 4 teams 369 / 372, 2 teams 347 / 323 */
  const bool lane_coder = teams >= 3u && !(opts && opts->enc_coder == 1);
  if (lane_coder)
    w3_encode_blocks_team<true><<<w3_ceil_div(n_blocks, teams), W3_ENC_LANE_THREADS(teams),
                                  sizeof(w3_smem_tab) + teams * sizeof(w3_enc_team), s>>>(
        d_in, n_total, block_size, n_blocks, d_scratch, stride, d_lens, d_overflow, d_tab, teams);
  else
    w3_encode_blocks_team<false><<<w3_ceil_div(n_blocks, teams), W3_ENC_TEAM_THREADS * teams,
                                   sizeof(w3_smem_tab) + teams * sizeof(w3_enc_team), s>>>(
        d_in, n_total, block_size, n_blocks, d_scratch, stride, d_lens, d_overflow, d_tab, teams);
  return cudaGetLastError();
}

cudaError_t w3_launch_pack(const uint8_t *d_scratch, uint64_t stride, const uint32_t *d_lens,
                           uint32_t n_blocks, uint64_t *d_offsets, const uint64_t *d_alt,
                           uint8_t *d_packed, cudaStream_t s) {
  w3_k_scan_lens<<<1, 1024, 0, s>>>(d_lens, n_blocks, d_offsets);
  if (n_blocks) {
    dim3 grid(n_blocks, 32);
    w3_k_pack<<<grid, 256, 0, s>>>(d_scratch, stride, d_offsets, d_alt, d_packed);
  }
  return cudaGetLastError();
}

cudaError_t w3_launch_scan(const uint32_t *d_lens, uint32_t n_blocks, uint64_t *d_offsets,
                           cudaStream_t s) {
  w3_k_scan_lens<<<1, 1024, 0, s>>>(d_lens, n_blocks, d_offsets);
  return cudaGetLastError();
}

/* The decode kernel w3_launch_decode takes for `n_blocks` blocks of `block_size` bytes on a device of `sm_count`
 * SMs under the options (kernels = W3_OPT_KERNELS, lane_form = W3_OPT_LANE_FORM): the launcher's policy in one
 * place (bench.py names the dominant kernel of its roofline through w3_decode_kernel_name).
 *   team   up to two waves of one CTA per block, and for blocks the 32-bit bit counters of the other forms cannot index
 *   split  beyond, while the blocks of an SM fit one CTA of it (w3_lane.cuh: 1 GiB / 1024 blocks 952 ms against the pair
 *          form's 1331, profiles/r02_spin_variants.txt, r02_split_fix_variants.txt); its pace does not depend on the number
 *          of blocks in the CTA, so more than 8 per SM would mean rounds
 *   pair   beyond that (4096 x 256 KiB: 681 ms against the split form's ~1000, profiles/r02_decode_policy.txt) */
int w3_decode_family(uint32_t n_blocks, uint32_t block_size, int sm_count, int kernels, int lane_form) {
  int k = kernels;
  /* the lane decoders count stream bits in 32 bits: blocks they cannot be sure about go to the team kernel */
  const bool lane_ok = block_size <= (16u << 20);
  if (k == W3_K_AUTO) k = (n_blocks > W3_TEAM_MAX_WAVES * (uint32_t)sm_count && lane_ok) ? W3_K_MANY : W3_K_TEAM;
  if (k == W3_K_MANY && !lane_ok) k = W3_K_TEAM;
  if (k == W3_K_TEAM) return W3_DEC_TEAM;
  if (k == W3_K_WARP) return W3_DEC_WARP;
  const uint32_t q = w3_ceil_div(n_blocks, (uint32_t)sm_count);
  if (lane_form == 4 || (lane_form == 0 && q <= W3_SPLIT_MAXB)) return W3_DEC_SPLIT;
  if (lane_form == 1) return W3_DEC_LANE;
  return lane_form == 3 ? W3_DEC_PAIR_P : W3_DEC_PAIR; /* the probability table measured slower (profiles/r02_pair_variants.txt): on request only */
}
const char *w3_decode_kernel_name(uint32_t n_blocks, uint32_t block_size, int sm_count, int kernels, int lane_form) {
  switch (w3_decode_family(n_blocks, block_size, sm_count, kernels, lane_form)) {
    case W3_DEC_TEAM: return "w3_decode_blocks_team";
    case W3_DEC_WARP: return "w3_decode_blocks_warp";
    case W3_DEC_LANE: return "w3_decode_blocks_lane";
    case W3_DEC_SPLIT: return "w3_decode_blocks_split";
    default: return "w3_decode_blocks_pair";
  }
}

cudaError_t w3_launch_decode(const uint8_t *d_comp, const uint64_t *d_offsets, uint32_t n_blocks,
                             uint32_t block_size, uint64_t n_total, uint8_t *d_out,
                             uint32_t *d_err, const w3_hashtab *d_tab, int sm_count,
                             const w3_launch_opts *opts, cudaStream_t s) {
  if (!n_blocks) return cudaSuccess;
  const int form = opts ? opts->lane_form : 0;
  const int fam = w3_decode_family(n_blocks, block_size, sm_count, opts ? opts->kernels : W3_K_AUTO, form);
  if (fam == W3_DEC_TEAM) {
    w3_decode_blocks_team<<<n_blocks, W3_DEC_TEAM_THREADS,
                            sizeof(w3_dec_tab) + sizeof(w3_dec_team), s>>>(
        d_comp, d_offsets, n_blocks, block_size, n_total, d_out, d_err, d_tab);
    return cudaGetLastError();
  }
  if (fam == W3_DEC_WARP) {
    int warps = w3_pick_warps(n_blocks, sm_count);
    uint32_t grid = (n_blocks + (uint32_t)warps - 1) / (uint32_t)warps;
    w3_decode_blocks_warp<<<grid, warps * 32, w3_dense_dec_smem_bytes(warps), s>>>(
        d_comp, d_offsets, n_blocks, block_size, n_total, d_out, d_tab);
    return cudaGetLastError();
  }
  if (fam == W3_DEC_SPLIT) {
    uint32_t group = w3_per_cta(n_blocks, sm_count, W3_SPLIT_MAXB);
    if (form == 4 && opts && opts->lane_warps > 0) group = (uint32_t)opts->lane_warps < W3_SPLIT_MAXB ? (uint32_t)opts->lane_warps : W3_SPLIT_MAXB;
    uint32_t grid = w3_ceil_div(n_blocks, group);
    if (grid > (uint32_t)sm_count) grid = (uint32_t)sm_count;
    w3_decode_blocks_split<<<grid, W3_SPLIT_THREADS, sizeof(w3_lane_tab) + sizeof(w3_split_ring) + (size_t)group * W3_SPLIT_WORDS * 4u, s>>>(
        d_comp, d_offsets, n_blocks, block_size, n_total, d_out, d_err, d_tab, group);
    return cudaGetLastError();
  }
  /* lanes: `resident` blocks per CTA at a time, spread over `warps` warps (one per SM
   * sub-partition by default) of `lanes` blocks each; the pair form (two lanes per block, the next
   * position's lookup overlapping the current bit's coder step) unless the lane form is asked for.
   * W3_DEC_PAIR_P keeps Counter::p of every counter in a second table (16 KiB per block). */
  const bool pair = fam != W3_DEC_LANE;
  const bool cache_p = fam == W3_DEC_PAIR_P;
  const uint32_t cap = cache_p ? w3_lane_cap() / 2u : w3_lane_cap();
  const uint32_t per = w3_per_cta(n_blocks, sm_count, cap);
  uint32_t warps = per < 4u ? per : 4u;
  if (opts && opts->lane_warps > 0) warps = (uint32_t)opts->lane_warps < W3_LANE_MAX_WARPS ? (uint32_t)opts->lane_warps : W3_LANE_MAX_WARPS;
  if (warps > per) warps = per;
  uint32_t lanes = w3_ceil_div(per, warps);
  while (warps * lanes > cap) lanes--; /* rounding up may overshoot what fits */
  if (pair && lanes > 16u) lanes = 16u;
  const uint32_t resident = warps * lanes;
  uint32_t grid = w3_ceil_div(n_blocks, resident);
  if (grid > (uint32_t)sm_count) grid = (uint32_t)sm_count;
  const size_t smem = sizeof(w3_lane_tab) + (size_t)resident * W3_NCTX * (cache_p ? 8u : 4u);
  if (cache_p)
    w3_decode_blocks_pair<true><<<grid, warps * 32u, smem, s>>>(d_comp, d_offsets, n_blocks, block_size, n_total, d_out, d_err, d_tab, lanes);
  else if (pair)
    w3_decode_blocks_pair<false><<<grid, warps * 32u, smem, s>>>(d_comp, d_offsets, n_blocks, block_size, n_total, d_out, d_err, d_tab, lanes);
  else
    w3_decode_blocks_lane<<<grid, warps * 32u, smem, s>>>(d_comp, d_offsets, n_blocks, block_size, n_total, d_out, d_err, d_tab, lanes);
  return cudaGetLastError();
}

cudaError_t w3_launch_train(const uint8_t *d_in, uint64_t n, uint16_t *d_cnt, uint16_t *d_table,
                            int sm_count, cudaStream_t s) {
  if (n) w3_k_bitpos_counts<<<sm_count * 8, 256, 0, s>>>(d_in, n, d_cnt);
  w3_k_stationary_scan<<<1, 256, 0, s>>>(d_in, n, d_cnt, d_table);
  return cudaGetLastError();
}

cudaError_t w3_launch_hash_array(const uint64_t *d_bits, const uint32_t *d_pos, uint32_t n,
                                 uint32_t *d_out, const w3_hashtab *d_tab, cudaStream_t s) {
  w3_k_hash_array<<<64, 256, sizeof(w3_smem_tab), s>>>(d_bits, d_pos, n, d_out, d_tab);
  return cudaGetLastError();
}
cudaError_t w3_launch_gather_probe(const void *d_buf, uint64_t n_sectors, uint32_t blocks, uint32_t per_thread,
                                   unsigned long long *d_sink, cudaStream_t s) {
  w3_k_gather_probe<<<blocks, 256, 0, s>>>((const uint4 *)d_buf, n_sectors, per_thread, d_sink);
  return cudaGetLastError();
}

cudaError_t w3_launch_counter_p_check(unsigned long long *d_mismatches, cudaStream_t s) {
  w3_k_counter_p_check<<<65535, 256, 0, s>>>(d_mismatches);
  return cudaGetLastError();
}
cudaError_t w3_launch_ac_encode_probs(const uint8_t *d_in, const uint16_t *d_probs,
                                      uint32_t n_probs, uint8_t *d_out, uint32_t cap,
                                      uint32_t *d_out_len, cudaStream_t s) {
  w3_k_ac_encode_probs<<<1, 32, 0, s>>>(d_in, d_probs, n_probs, d_out, cap, d_out_len);
  return cudaGetLastError();
}
cudaError_t w3_launch_ac_decode_probs(const uint8_t *d_in, uint32_t n_in, const uint16_t *d_probs,
                                      uint32_t n_probs, uint8_t *d_out, cudaStream_t s) {
  w3_k_ac_decode_probs<<<1, 32, 0, s>>>(d_in, n_in, d_probs, n_probs, d_out);
  return cudaGetLastError();
}

} /* extern "C" */
