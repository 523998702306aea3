/*
 * w3_team.cuh — "team" kernels: one CTA per block, the serial chain of the
 * coder isolated in a warp of its own.  Included by w3_kernels.cu.
 *
 * The path is bound by one dependent chain per block (bit -> range -> renorm
 * -> next bit).  A single warp that also hashes, looks up, divides and updates
 * pays for all of those instructions on the same chain (measured round 1:
 * 244 / 1042 cycles per bit for encode / decode, profiles/r01_v1_*).  With few
 * blocks per GPU (64 x 1 MiB on 148 SMs) the idle SM resources are spent on
 * taking everything else off that chain:
 *
 *   encode   warp 0  coder: arithmetic-coder steps only, (bit, P) pairs read
 *                    from a shared-memory ring
 *            warp 1  model: the warp-parallel predict/update of the v1 kernel
 *                    (history -> threshold-table hash -> __match_any counter
 *                    resolution -> Counter::p), 32 bit positions per pass
 *
 *   decode   warp 0        decoder: decode steps only.  For bit t it picks one
 *                          of 32 prepared candidates by its last 5 decoded bits
 *            warps 1,2,3,5 speculators: warp j owns positions == j (mod 4).
 *                          For position pos its 32 lanes evaluate context and
 *                          probability for all 2^5 values of bits pos-5..pos-1
 *                          as soon as bit pos-6 is known, and apply the counter
 *                          update of position pos-8.  Two positions that share
 *                          a counter are a multiple of 8 apart, hence owned by
 *                          the same warp: program order alone keeps every
 *                          counter read behind the update it depends on.
 *            warp 4        idle (keeps the decoder alone on its scheduler)
 *
 * Warps of one CTA talk through volatile shared-memory words; every wait is
 * bounded (W3_SPIN_LIMIT) and trips an abort flag instead of hanging.
 */
#pragma once

#define W3_SPIN_LIMIT (1u << 25)
#define W3_ENC_RING 64u /* groups of 32 bits in flight between model and coder */
#define W3_SPEC_D 5     /* speculated bits per position: 2^5 = 32 lanes */
#define W3_CAND_RING 16u

#define W3_ERR_OVERFLOW 1u
#define W3_ERR_ABORT 2u

__device__ __forceinline__ uint32_t w3_lds_volatile(const volatile uint32_t *p) { return *p; }

/* wait until (int)(*flag - v) > 0, bounded */
__device__ __forceinline__ bool w3_wait_gt(const volatile uint32_t *flag, uint32_t v,
                                           volatile uint32_t *abort_flag) {
  uint32_t spins = 0;
  while ((int32_t)(*flag - v) <= 0) {
    if (++spins > W3_SPIN_LIMIT || *abort_flag) { *abort_flag = 1u; return false; }
  }
  return true;
}

/* ------------------------------------------------------------------------ */
/* encode                                                                   */

struct w3_enc_team {
  uint32_t stats[W3_NCTX];          /* c0 | c1 << 16 */
  uint32_t ring[W3_ENC_RING][32];   /* (p << 16) | bit per coded bit */
  volatile uint32_t prod_done;      /* groups written by the model warp */
  volatile uint32_t cons_done;      /* groups consumed by the coder warp */
  volatile uint32_t abort_flag;
  uint32_t pad;
};

__global__ void __launch_bounds__(64)
w3_encode_blocks_team(const uint8_t *__restrict__ in, uint64_t n_total, uint32_t block_size,
                      uint32_t n_blocks, uint8_t *__restrict__ scratch, uint64_t stride,
                      uint32_t *__restrict__ lens, uint32_t *__restrict__ err,
                      const w3_hashtab *__restrict__ tab) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  w3_smem_tab *st = reinterpret_cast<w3_smem_tab *>(smem_raw);
  w3_enc_team *tm = reinterpret_cast<w3_enc_team *>(smem_raw + sizeof(w3_smem_tab));
  const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31u;

  w3_load_tab(st, tab);
  for (uint32_t i = threadIdx.x; i < W3_NCTX; i += blockDim.x) tm->stats[i] = 0u;
  if (threadIdx.x == 0) { tm->prod_done = 0; tm->cons_done = 0; tm->abort_flag = 0; }
  __syncthreads();

  const uint32_t blk = blockIdx.x;
  if (blk >= n_blocks) return;
  const uint64_t off = (uint64_t)blk * block_size;
  const uint32_t n = (uint32_t)(n_total - off < block_size ? n_total - off : block_size);
  const uint32_t n_words = (n + 3u) >> 2; /* groups of 32 coded bits */

  if (warp == 1) {
    /* ---- model warp: (bit, p) for 32 positions per pass -------------------- */
    const uint32_t *src = reinterpret_cast<const uint32_t *>(in + off);
    uint32_t *stats = tm->stats;
    auto load_chunk = [&](uint32_t chunk) -> uint32_t {
      uint32_t widx = chunk * 32u + lane, v = 0;
      if (widx < n_words) {
        v = w3_bswap(__ldg(src + widx));
        uint32_t rem = n - widx * 4u;
        if (rem < 4u) v &= 0xFFFFFFFFu << (8u * (4u - rem));
      }
      return v;
    };
    uint32_t w0 = 0, w1 = 0;
    uint32_t cw_next = load_chunk(0);
    for (uint32_t chunk = 0; chunk * 32u < n_words; chunk++) {
      const uint32_t cw = cw_next;
      cw_next = load_chunk(chunk + 1);
      const uint32_t in_chunk = n_words - chunk * 32u < 32u ? n_words - chunk * 32u : 32u;
      for (uint32_t i = 0; i < in_chunk; i++) {
        const uint32_t w2 = __shfl_sync(FULL, cw, i);
        const uint32_t g = chunk * 32u + i;
        const uint32_t left = n * 8u - g * 32u;
        const uint32_t nact = left < 32u ? left : 32u;
        const bool active = lane < nact;
        const uint32_t t = g * 32u + lane;

        const uint32_t hist_hi = __funnelshift_l(w1, w0, lane);
        const uint32_t hist_lo = __funnelshift_l(w2, w1, lane);
        const uint32_t bit = (w2 >> (31u - lane)) & 1u;
        const uint64_t key = ((uint64_t)__brev(hist_lo) << 32) | __brev(hist_hi);
        uint32_t ctx = (w3_hash_dev(st, key, t) << W3_ALIGN_BITS) | (t & 7u);
        if (t == 0) ctx = 0;
        if (!active) ctx = 0x10000u | lane;

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
          p = w3_counter_p(c0 + e0, c1 + e1);
          if (active && lane == 31u - (uint32_t)__clz((int)m))
            stats[ctx] = (c0 + t0) | ((c1 + t1) << 16);
        }
        /* hand the group to the coder */
        if (g >= W3_ENC_RING && !w3_wait_gt(&tm->cons_done, g - W3_ENC_RING, &tm->abort_flag)) return;
        ((volatile uint32_t *)tm->ring[g & (W3_ENC_RING - 1u)])[lane] = active ? ((p << 16) | bit) : 0u;
        __syncwarp();
        if (lane == 0) {
          __threadfence_block();
          tm->prod_done = g + 1u;
        }
        w0 = w1;
        w1 = w2;
      }
    }
    return;
  }

  /* ---- coder warp: ac.encode(bit, p) in order (main.rs:107) ----------------- */
  uint32_t *dst = reinterpret_cast<uint32_t *>(scratch + (uint64_t)blk * stride);
  const uint32_t cap_words = (uint32_t)(stride >> 2);
  uint32_t ovf = 0;
  auto store = [&](uint32_t idx, uint32_t v) {
    if (lane == 0) {
      if (idx < cap_words) dst[idx] = w3_bswap(v);
      else ovf = 1;
    }
  };
  uint32_t x1 = 0, x2 = 0xFFFFFFFFu;
  w3_writer wr;
  w3_writer_init(wr);
  bool dead = false;
  for (uint32_t g = 0; g < n_words; g++) {
    if (!w3_wait_gt(&tm->prod_done, g, &tm->abort_flag)) { dead = true; break; }
    const uint32_t v = ((const volatile uint32_t *)tm->ring[g & (W3_ENC_RING - 1u)])[lane];
    const uint32_t left = n * 8u - g * 32u;
    const uint32_t nact = left < 32u ? left : 32u;
#pragma unroll 4
    for (uint32_t j = 0; j < nact; j++) {
      const uint32_t vj = __shfl_sync(FULL, v, j);
      w3_ac_narrow(x1, x2, vj & 1u, vj & 0xFFFF0000u);
      const uint32_t pre = x1;
      int s, k;
      w3_ac_renorm(x1, x2, s, k);
      w3_writer_step(wr, pre, s, k, store);
    }
    if (lane == 0) tm->cons_done = g + 1u;
  }
  uint32_t tail, tail_bytes;
  const uint64_t len = w3_writer_flush(wr, x2, tail, tail_bytes, store);
  if (lane == 0) {
    uint8_t *tb = reinterpret_cast<uint8_t *>(dst) + (uint64_t)wr.words * 4u;
    for (uint32_t b = 0; b < tail_bytes; b++) {
      if ((uint64_t)wr.words * 4u + b < stride) tb[b] = (uint8_t)(tail >> (24u - 8u * b));
      else ovf = 1;
    }
    lens[blk] = (uint32_t)len;
    if (ovf) atomicOr(err, W3_ERR_OVERFLOW);
    if (dead || tm->abort_flag) atomicOr(err, W3_ERR_ABORT);
  }
}

/* ------------------------------------------------------------------------ */
/* decode                                                                   */

struct w3_dec_team {
  uint2 stats[W3_NCTX];                  /* .x = c0 | c1 << 16, .y = Counter::p << 16 */
  uint32_t cand[W3_CAND_RING][32];       /* (p << 16) | ctx for every guess of the last 5 bits */
  volatile uint32_t spec_flag[W3_CAND_RING]; /* pos + 1 once cand[pos % 16] is complete */
  volatile uint32_t pub;                 /* (bits decoded << 16) | last 16 decoded bits */
  volatile uint32_t abort_flag;
  uint32_t pad[2];
};

#define W3_DEC_TEAM_THREADS 192

__global__ void __launch_bounds__(W3_DEC_TEAM_THREADS)
w3_decode_blocks_team(const uint8_t *__restrict__ comp, const uint64_t *__restrict__ offsets,
                      uint32_t n_blocks, uint32_t block_size, uint64_t n_total,
                      uint8_t *__restrict__ out, uint32_t *__restrict__ err,
                      const w3_hashtab *__restrict__ tab) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  w3_smem_tab *st = reinterpret_cast<w3_smem_tab *>(smem_raw);
  w3_dec_team *tm = reinterpret_cast<w3_dec_team *>(smem_raw + sizeof(w3_smem_tab));
  const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31u;

  w3_load_tab(st, tab);
  for (uint32_t i = threadIdx.x; i < W3_NCTX; i += blockDim.x)
    tm->stats[i] = make_uint2(0u, 0x80000000u); /* Counter::new: p = 32768 */
  if (threadIdx.x < W3_CAND_RING) tm->spec_flag[threadIdx.x] = 0;
  if (threadIdx.x == 0) { tm->pub = 0; tm->abort_flag = 0; }
  __syncthreads();

  const uint32_t blk = blockIdx.x;
  if (blk >= n_blocks || warp == 4) return;
  const uint64_t off = (uint64_t)blk * block_size;
  const uint32_t n = (uint32_t)(n_total - off < block_size ? n_total - off : block_size);
  const uint32_t nbits = n * 8u;

  if (warp != 0) {
    /* ---- speculator j: positions j, j+4, j+8, … ----------------------------- */
    const uint32_t j = warp < 4 ? warp - 1u : 3u;
    uint64_t hist = 0;   /* decoded bits, newest at bit 0 */
    uint32_t known = 0;  /* how many of them */
    uint32_t ctx_m4 = 0, ctx_m8 = 0; /* this lane's context guess for pos-4 / pos-8 */
    for (uint32_t pos = j; pos < nbits; pos += 4) {
      /* wait for bit pos-6 (bits 0 .. pos-6 known = pos-5 bits) */
      const uint32_t need = pos > W3_SPEC_D ? pos - W3_SPEC_D : 0u;
      uint32_t spins = 0;
      for (;;) {
        const uint32_t pv = tm->pub;
        const uint32_t delta = ((pv >> 16) - known) & 0xFFFFu; /* new bits, <= 16 by construction */
        if (known + delta >= need) {
          hist = (hist << delta) | (uint64_t)(pv & ((1u << delta) - 1u));
          known += delta;
          break;
        }
        if (++spins > W3_SPIN_LIMIT || tm->abort_flag) { tm->abort_flag = 1u; return; }
      }
      /* adapt (main.rs:136 -> ordern_entropy.rs:32-34) for position pos-8, whose
       * counter this warp owns; its context is the guess that came true */
      if (pos >= 8u) {
        const uint32_t q = pos - 8u;
        const uint32_t hq = (uint32_t)(hist >> (known - q)) & 31u;      /* bits q-5 .. q-1 */
        const uint32_t bq = (uint32_t)(hist >> (known - 1u - q)) & 1u;  /* bit q */
        const uint32_t cq = __shfl_sync(FULL, ctx_m8, hq);
        const uint32_t nv = w3_counter_update(tm->stats[cq].x, bq);
        const uint32_t np = w3_counter_p(nv & 0xFFFFu, nv >> 16) << 16;
        if (lane == 0) tm->stats[cq] = make_uint2(nv, np);
        __syncwarp();
      }
      /* predict for position pos under guess `lane` of bits pos-5 .. pos-1 */
      const uint64_t base = pos >= W3_SPEC_D ? hist >> (known - (pos - W3_SPEC_D)) : 0ull;
      const uint64_t full = (base << W3_SPEC_D) | lane;
      uint32_t ctx = (w3_hash_dev(st, w3_key_of_history(full), pos) << W3_ALIGN_BITS) | (pos & 7u);
      if (pos == 0) ctx = 0;
      const uint32_t P = tm->stats[ctx].y;
      ((volatile uint32_t *)tm->cand[pos & (W3_CAND_RING - 1u)])[lane] = P | ctx;
      __syncwarp();
      if (lane == 0) {
        __threadfence_block();
        tm->spec_flag[pos & (W3_CAND_RING - 1u)] = pos + 1u;
      }
      ctx_m8 = ctx_m4;
      ctx_m4 = ctx;
    }
    return;
  }

  /* ---- decoder: decode steps only (main.rs:131-140) ------------------------- */
  uint8_t *dst = out + off;
  w3_reader rd;
  w3_reader_init(rd, comp + offsets[blk], offsets[blk + 1] - offsets[blk]);
  uint32_t x1 = 0, x2 = 0xFFFFFFFFu;
  uint32_t x = w3_reader_take(rd, 32);
  uint32_t hist = 0;
  bool dead = false;
  uint32_t flag_next = tm->spec_flag[0];
  for (uint32_t t = 0; t < nbits; t++) {
    /* candidate set for bit t (normally long since published) */
    if (flag_next != t + 1u) {
      if (!w3_wait_gt(&tm->spec_flag[t & (W3_CAND_RING - 1u)], t, &tm->abort_flag)) { dead = true; break; }
    }
    const uint32_t c = ((const volatile uint32_t *)tm->cand[t & (W3_CAND_RING - 1u)])[hist & 31u];
    flag_next = tm->spec_flag[(t + 1u) & (W3_CAND_RING - 1u)];
    const uint32_t xmid = x1 + __umulhi(x2 - x1, c & 0xFFFF0000u);
    const uint32_t bit = x <= xmid ? 1u : 0u; /* arithmetic_coder.rs:82 */
    if (bit) x2 = xmid; else x1 = xmid + 1u;
    hist = (hist << 1) | bit;
    tm->pub = ((t + 1u) << 16) | (hist & 0xFFFFu);
    int s, k;
    const int nsh = w3_ac_renorm(x1, x2, s, k);
    /* take(0) must read nothing: shift in two steps */
    const uint32_t fresh = (uint32_t)((rd.win >> 1) >> (63 - nsh));
    x = ((x << nsh) | fresh) ^ (k ? 0x80000000u : 0u);
    rd.win <<= nsh;
    rd.have -= nsh;
    if (rd.have <= 32) {
      rd.win |= (uint64_t)w3_reader_word(rd) << (32 - rd.have);
      rd.have += 32;
    }
    if ((t & 31u) == 31u && lane == 0)
      *reinterpret_cast<uint32_t *>(dst + (t >> 5) * 4u) = w3_bswap(hist);
  }
  if (lane == 0) {
    const uint32_t rem = n & 3u; /* bytes after the last whole word */
    for (uint32_t b = 0; b < rem; b++) dst[(n & ~3u) + b] = (uint8_t)(hist >> (8u * (rem - 1u - b)));
    if (dead || tm->abort_flag) atomicOr(err, W3_ERR_ABORT);
  }
}
