/*
 * w3_cm.cu — kernels of the hashed context-mixing stack (see w3_cm.cuh; the
 * specification is oracle/w3_cm_oracle.c).  One warp per block of input; lanes
 * 0..5 are the context models, lane 6 carries the mixer's bias input, lane 0
 * also runs the APM and the arithmetic coder.
 *
 * Per nibble and model one 96-byte cell (src/hashmap.rs:32-35) travels from the
 * block's table in HBM to a staging row in shared memory and back: 3 sectors in,
 * 3 out.  Two models of one nibble can land in the same cell (6 models into
 * 2^L cells); the oracle applies their operations to the one cell in model
 * order, so lanes that share a cell share one staging row and take turns
 * (__match_any_sync finds them; the common case runs fully parallel).
 */
#include <cuda_runtime.h>
#include <stdint.h>

#include "w3_cm.cuh"
#include "w3_kernels.h"

#define FULL 0xFFFFFFFFu
#define W3_CM_WARPS 8u

struct w3_cm_tabs { /* per CTA */
  w3_state_entry st[W3_ST_SIZE + 1];
  int16_t stretch[4096];
  uint8_t bits_of[4096];
  int16_t sq33[40];
};
struct w3_cm_warp { /* per block of input */
  uint16_t apm[256 * W3_CM_APM_BINS];
  int32_t w[8][8];
  __align__(16) uint8_t stage[W3_CM_MODELS][W3_CELL_BYTES];
};

template <bool kDecode>
__global__ void __launch_bounds__(W3_CM_WARPS * 32u)
w3_cm_blocks(const uint8_t *__restrict__ src, const uint64_t *__restrict__ offsets /* decode: stream offsets */,
             uint64_t n_total, uint32_t block_size, uint32_t n_blocks, uint8_t *__restrict__ dst,
             uint64_t stride /* encode: scratch stride */, uint32_t *__restrict__ lens, uint32_t *__restrict__ err,
             uint8_t *__restrict__ cells, uint32_t log_cells) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  w3_cm_tabs *tb = reinterpret_cast<w3_cm_tabs *>(smem_raw);
  const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31u;
  w3_cm_warp *ws = reinterpret_cast<w3_cm_warp *>(smem_raw + sizeof(w3_cm_tabs)) + warp;

  { /* tables: NaiveStateTable, squash anchors, stretch, bits per state */
    const int16_t sq[33] = W3_CM_SQ33;
    if (threadIdx.x < 33) tb->sq33[threadIdx.x] = sq[threadIdx.x];
    for (uint32_t i = threadIdx.x; i < W3_ST_SIZE; i += blockDim.x) {
      tb->st[i] = w3_state_entry_of(i);
      tb->bits_of[i] = (uint8_t)w3_cm_state_bits(i);
    }
    __syncthreads();
    for (uint32_t i = threadIdx.x; i < 4096u; i += blockDim.x) tb->stretch[i] = (int16_t)w3_cm_stretch_entry((int)i, tb->sq33);
  }
  for (uint32_t i = lane; i < 256u * W3_CM_APM_BINS; i += 32u)
    ws->apm[i] = (uint16_t)(w3_cm_squash(((int)(i % W3_CM_APM_BINS) - 16) * 128, tb->sq33) * 16);
  for (uint32_t i = lane; i < 64u; i += 32u) ws->w[i >> 3][i & 7u] = (i & 7u) < W3_CM_MODELS ? 0x4000 : 0;
  __syncthreads();

  const uint32_t blk = blockIdx.x * W3_CM_WARPS + warp;
  if (blk >= n_blocks) return;
  const uint64_t off = (uint64_t)blk * block_size;
  const uint32_t n = (uint32_t)(n_total - off < block_size ? n_total - off : block_size);
  uint8_t *my_cells = cells + ((size_t)blk * W3_CELL_BYTES << log_cells);

  /* coder state (lane 0's is the one that counts) */
  uint32_t x1 = 0u, x2 = 0xFFFFFFFFu, x = 0u;
  w3_writer wr;
  w3_writer_init(wr);
  uint32_t *out_w = reinterpret_cast<uint32_t *>(dst + (uint64_t)blk * stride);
  const uint32_t cap_words = (uint32_t)(stride >> 2);
  uint32_t ovf = 0;
  auto store = [&](uint32_t idx, uint32_t v) {
    if (lane == 0) {
      if (idx < cap_words) out_w[idx] = __byte_perm(v, 0, 0x0123);
      else ovf = 1;
    }
  };
  const uint8_t *sp = kDecode ? src + offsets[blk] : src + off;
  const uint64_t slen = kDecode ? offsets[blk + 1] - offsets[blk] : 0;
  uint64_t bp = 0;
  auto take = [&](int c) -> uint32_t { /* c <= 32 stream bits, MSB first, zero past the end — io.rs:18-48 */
    uint64_t acc = 0;
    const uint64_t by = bp >> 3;
    for (uint32_t b = 0; b < 5u; b++) acc = (acc << 8) | ((kDecode && by + b < slen) ? sp[by + b] : 0u);
    const uint32_t v = c ? (uint32_t)((acc << (24 + (bp & 7u))) >> (64 - c)) : 0u;
    bp += (uint64_t)c;
    return v;
  };
  if (kDecode) x = take(32);
  uint8_t *out_b = dst + off; /* decode */

  uint64_t last8 = 0;
  for (uint32_t i = 0; i < n; i++) {
    uint32_t byte = kDecode ? 0u : sp[i];
    uint32_t c0 = 1u;
    for (uint32_t nb = 0; nb < 2u; nb++) {
      const uint32_t hi = kDecode ? (byte & 15u) : (byte >> 4);
      /* ---- nibble start: hash, fetch the cell, find the slot ---------------------------------- */
      uint64_t h = 0;
      uint32_t cidx = 0xFFFFFF00u | lane; /* lanes without a model: unique */
      if (lane < W3_CM_MODELS) {
        h = w3_cm_hash(last8, lane, nb, hi);
        cidx = (uint32_t)(h >> (64u - log_cells));
      }
      const uint32_t grp = __match_any_sync(FULL, cidx);
      const uint32_t leader = (uint32_t)__ffs((int)grp) - 1u;
      const bool shared_cell = __any_sync(FULL, lane < W3_CM_MODELS && leader != lane);
      uint8_t *g_cell = my_cells + (size_t)cidx * W3_CELL_BYTES;
      if (lane < W3_CM_MODELS && leader == lane) {
        const uint4 *g4 = reinterpret_cast<const uint4 *>(g_cell);
        uint4 *s4 = reinterpret_cast<uint4 *>(ws->stage[lane]);
#pragma unroll
        for (int k = 0; k < 6; k++) s4[k] = g4[k];
      }
      __syncwarp();
      uint8_t *cell = ws->stage[lane < W3_CM_MODELS ? leader : 0u];
      uint32_t slot = 0;
      if (!shared_cell) {
        if (lane < W3_CM_MODELS) slot = w3_cm_find_slot(cell, h, tb->bits_of);
      } else { /* the oracle finds the slots in model order on the one cell */
        for (uint32_t k = 0; k < W3_CM_MODELS; k++) {
          if (lane == k) slot = w3_cm_find_slot(cell, h, tb->bits_of);
          __syncwarp();
        }
      }
      __syncwarp();
      uint32_t nib_ctx = 0;
#pragma unroll
      for (uint32_t j = 0; j < 4u; j++) {
        const uint32_t sel = nb * 4u + j;
        /* ---- predict ------------------------------------------------------------------------ */
        uint32_t s = 0;
        int st = 256; /* lane 6: the bias input */
        if (lane < W3_CM_MODELS) {
          s = w3_slot_get_state(cell, slot, j, nib_ctx);
          st = tb->stretch[tb->st[s].prob >> 4];
        }
        int term = lane <= W3_CM_MODELS ? (ws->w[sel][lane] * st) >> 8 : 0;
        term += __shfl_xor_sync(FULL, term, 1);
        term += __shfl_xor_sync(FULL, term, 2);
        term += __shfl_xor_sync(FULL, term, 4);
        const int dot = __shfl_sync(FULL, term, 0);
        const int pm = w3_cm_squash(dot >> 8, tb->sq33);
        const int sx = tb->stretch[pm] + 2048;
        const int aw = sx & 127;
        const uint32_t aidx = c0 * W3_CM_APM_BINS + (uint32_t)(sx >> 7);
        const int a0 = ws->apm[aidx], a1 = ws->apm[aidx + 1u];
        const int pa = (a0 * (128 - aw) + a1 * aw) >> 11;
        int p = (pm + 3 * pa + 2) >> 2;
        p = p < 1 ? 1 : p > 4095 ? 4095 : p;
        const uint32_t P = (uint32_t)p << 20; /* lerp's 32-bit probability: (p << 4) << 16 */
        /* ---- code (the reference's coder: arithmetic_coder.rs:41-65 / 80-105) ------------------ */
        uint32_t bit;
        if (!kDecode) {
          bit = (byte >> (7u - sel)) & 1u;
          w3_ac_narrow(x1, x2, bit, P);
          const uint32_t pre = x1;
          int sh_s, sh_k;
          w3_ac_renorm(x1, x2, sh_s, sh_k);
          w3_writer_step(wr, pre, sh_s, sh_k, store);
        } else {
          const uint32_t xmid = x1 + __umulhi(x2 - x1, P);
          bit = x <= xmid ? 1u : 0u;
          if (bit) x2 = xmid; else x1 = xmid + 1u;
          int sh_s, sh_k;
          const int nsh = w3_ac_renorm(x1, x2, sh_s, sh_k);
          if (nsh) x = ((x << nsh) | take(nsh)) ^ (sh_k ? 0x80000000u : 0u);
          byte = (byte << 1) | bit;
        }
        /* ---- update -------------------------------------------------------------------------- */
        const int e = (((int)bit << 12) - pm) * W3_CM_LR;
        if (lane <= W3_CM_MODELS) ws->w[sel][lane] = w3_cm_train(ws->w[sel][lane], st, e);
        if (lane == 0) {
          const int g = (int)((bit << 16) + (bit << 7) - bit - bit);
          ws->apm[aidx] = (uint16_t)(a0 + ((g - a0) >> 7));
          ws->apm[aidx + 1u] = (uint16_t)(a1 + ((g - a1) >> 7));
        }
        const uint32_t ns = lane < W3_CM_MODELS ? tb->st[s].next[bit] : 0u;
        if (!shared_cell) {
          if (lane < W3_CM_MODELS) w3_slot_set_state(cell, slot, j, nib_ctx, ns);
        } else {
          for (uint32_t k = 0; k < W3_CM_MODELS; k++) {
            if (lane == k) w3_slot_set_state(cell, slot, j, nib_ctx, ns);
            __syncwarp();
          }
        }
        __syncwarp();
        c0 = (c0 << 1) | bit;
        nib_ctx = (nib_ctx << 1) | bit;
      }
      /* ---- the cell goes back ------------------------------------------------------------------ */
      if (lane < W3_CM_MODELS && leader == lane) {
        uint4 *g4 = reinterpret_cast<uint4 *>(g_cell);
        const uint4 *s4 = reinterpret_cast<const uint4 *>(ws->stage[lane]);
#pragma unroll
        for (int k = 0; k < 6; k++) g4[k] = s4[k];
      }
      __syncwarp();
    }
    if (kDecode && lane == 0) out_b[i] = (uint8_t)byte;
    last8 = (last8 << 8) | byte;
  }
  if (!kDecode) { /* ac.flush — arithmetic_coder.rs:67-71 */
    uint32_t tail, tail_bytes;
    const uint64_t len = w3_writer_flush(wr, x2, tail, tail_bytes, store);
    if (lane == 0) {
      uint8_t *tbp = reinterpret_cast<uint8_t *>(out_w) + (uint64_t)wr.words * 4u;
      for (uint32_t b = 0; b < tail_bytes; b++) {
        if ((uint64_t)wr.words * 4u + b < stride) tbp[b] = (uint8_t)(tail >> (24u - 8u * b));
        else ovf = 1;
      }
      lens[blk] = (uint32_t)len;
      if (ovf) atomicOr(err, W3_ERR_OVERFLOW);
    }
  }
}

extern "C" {
size_t w3_cm_smem_bytes(void) { return sizeof(w3_cm_tabs) + W3_CM_WARPS * sizeof(w3_cm_warp); }
cudaError_t w3_cm_init(void) {
  cudaError_t e = cudaFuncSetAttribute(w3_cm_blocks<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)w3_cm_smem_bytes());
  if (e != cudaSuccess) return e;
  return cudaFuncSetAttribute(w3_cm_blocks<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)w3_cm_smem_bytes());
}
cudaError_t w3_launch_cm_encode(const uint8_t *d_in, uint64_t n_total, uint32_t block_size, uint32_t n_blocks,
                                uint8_t *d_scratch, uint64_t stride, uint32_t *d_lens, uint32_t *d_err, uint8_t *d_cells,
                                uint32_t log_cells, cudaStream_t s) {
  if (!n_blocks) return cudaSuccess;
  w3_cm_blocks<false><<<(n_blocks + W3_CM_WARPS - 1) / W3_CM_WARPS, W3_CM_WARPS * 32u, w3_cm_smem_bytes(), s>>>(
      d_in, nullptr, n_total, block_size, n_blocks, d_scratch, stride, d_lens, d_err, d_cells, log_cells);
  return cudaGetLastError();
}
cudaError_t w3_launch_cm_decode(const uint8_t *d_comp, const uint64_t *d_offsets, uint64_t n_total, uint32_t block_size,
                                uint32_t n_blocks, uint8_t *d_out, uint32_t *d_err, uint8_t *d_cells, uint32_t log_cells,
                                cudaStream_t s) {
  if (!n_blocks) return cudaSuccess;
  w3_cm_blocks<true><<<(n_blocks + W3_CM_WARPS - 1) / W3_CM_WARPS, W3_CM_WARPS * 32u, w3_cm_smem_bytes(), s>>>(
      d_comp, d_offsets, n_total, block_size, n_blocks, d_out, 0, nullptr, d_err, d_cells, log_cells);
  return cudaGetLastError();
}
}
