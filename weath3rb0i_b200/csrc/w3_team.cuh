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
 *            warp 1  model: the warp-parallel predict/update of the warp kernel
 *                    (history -> threshold-table rank -> __match_any counter
 *                    resolution -> Counter::p), 32 bit positions per pass
 *            warp 2  writer: turns the coder's (prefix, shift, pending) records
 *                    into stream words
 *
 *   decode   warp 0   decoder: decode steps only, branch-free ("attempts").
 *                     For bit t it picks one of 32 prepared probabilities by
 *                     its last 5 decoded bits; the entry's tag says whether it
 *                     is there yet.  Alone on scheduler 0.
 *            16 warps speculators: warp j owns positions == j (mod 16).  For
 *                     position pos its 32 lanes evaluate context and
 *                     probability for all 2^5 values of bits pos-5..pos-1 as
 *                     soon as bit pos-6 is known (sleeping on the mbarrier of
 *                     that bit until then), publish them, prepare the
 *                     counter's update for either value of bit pos, and
 *                     commit it when bit pos arrives.  Positions that share a
 *                     counter are a multiple of 8 apart; a per-alignment
 *                     `done` word orders the next predict behind the commit.
 *            1 warp   feeder: stages the block's stream in shared memory
 *
 * Warps of one CTA talk through volatile shared-memory words and mbarriers;
 * every wait is bounded and trips an abort flag instead of hanging.  Rare paths
 * are out of line so that the steady-state loops are straight-line code: a
 * branch costs a lone warp 20-40 cycles (profiles/r01_microbench5_*).
 *
 * Build-time switches (scripts/variants.sh builds A/B libraries with them; what
 * each measured is in profiles/r01_{encode,decode}_team_experiments.txt):
 *   W3_SPEC_WARPS 8|16, W3_DEC_FIX 3|4|6, W3_ATTEMPTS_PER_TRIP, W3_OUT_RING  sizes
 *   W3_FETCH_EARLY, W3_FETCH_BEFORE_ARRIVE, W3_P_AFTER_COMPARE, W3_LOADS_PREDICATED  schedule variants
 *   W3_STATS, W3_DEBUG                                                           counters / clock64 probes + printf
 *                      (a W3_STATS build runs up to 5 % faster or slower than the plain one: compare like builds)
 *   W3_UNSAFE_*  take a part of the work out (wrong output!) to bound its cost; never shipped
 */
#pragma once

#define W3_SPIN_LIMIT (1u << 25)
#define W3_ENC_RING 64u /* groups of 32 bits in flight between model and coder */
#define W3_SPEC_D 5u    /* speculated bits per position: 2^D guesses, 2^(D-5) per lane */
#define W3_GPL (1u << (W3_SPEC_D - 5u)) /* guesses per lane */
#ifndef W3_SPEC_WARPS
#define W3_SPEC_WARPS 16u /* speculator warps: warp j owns positions = j mod W3_SPEC_WARPS */
#endif
#define W3_CAND_RING 16u
#ifndef W3_ATTEMPTS_PER_TRIP
#define W3_ATTEMPTS_PER_TRIP 16u /* decoder: straight-line attempts per loop trip (<= 32) */
#endif


/* shared-memory words by 32-bit address: keeps the address in a register
 * (no generic-pointer arithmetic per access) and fixes the issue order */
__device__ __forceinline__ uint32_t w3_saddr(const volatile void *p) {
  return (uint32_t)__cvta_generic_to_shared(const_cast<const void *>(p));
}
__device__ __forceinline__ uint32_t w3_lds(uint32_t a) {
  uint32_t v;
  asm volatile("ld.volatile.shared.u32 %0, [%1];" : "=r"(v) : "r"(a) : "memory");
  return v;
}
__device__ __forceinline__ uint2 w3_lds64(uint32_t a) {
  uint2 v;
  asm volatile("ld.volatile.shared.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(a) : "memory");
  return v;
}
__device__ __forceinline__ void w3_sts(uint32_t a, uint32_t v) {
  asm volatile("st.volatile.shared.u32 [%0], %1;" ::"r"(a), "r"(v) : "memory");
}
__device__ __forceinline__ void w3_sts64(uint32_t a, uint2 v) {
  asm volatile("st.volatile.shared.v2.u32 [%0], {%1, %2};" ::"r"(a), "r"(v.x), "r"(v.y) : "memory");
}

/* wait until (int)(word at `flag_a` - v) > 0; false (and abort raised) on timeout */
__device__ __noinline__ bool w3_wait_gt(uint32_t flag_a, uint32_t v, uint32_t abort_a) {
  for (uint32_t spins = 0; spins < W3_SPIN_LIMIT; spins++) {
    if ((int32_t)(w3_lds(flag_a) - v) > 0) return true;
    if (w3_lds(abort_a)) break;
  }
#ifdef W3_DEBUG
  if ((threadIdx.x & 31) == 0) printf("[w3] wait_gt timeout warp %d flag_a %u v %u cur %u\n", threadIdx.x >> 5, flag_a, v, w3_lds(flag_a));
#endif
  w3_sts(abort_a, 1u);
  return false;
}

/* ------------------------------------------------------------------------ */
/* encode                                                                   */

#ifndef W3_OUT_RING
#define W3_OUT_RING 16u /* groups of 32 coder steps in flight between coder and writer */
#endif

struct w3_enc_team {
  uint32_t stats[W3_NCTX];            /* c0 | c1 << 16 */
  uint32_t ring[W3_ENC_RING][32];     /* model -> coder: (p << 16) | bit per coded bit */
  uint2 oring[W3_OUT_RING][32];       /* coder -> writer: the bounds (x1, x2) after narrowing, per step */
  uint32_t prod_done;                 /* groups written by the model warp */
  uint32_t cons_done;                 /* groups consumed by the coder warp */
  uint32_t out_done;                  /* groups of steps published by the coder warp */
  uint32_t wr_done;                   /* groups consumed by the writer warp */
  uint32_t final_x2;                  /* x2 at the end of the block, for ac.flush */
  uint32_t abort_flag;
  uint32_t pad[28];                   /* the size in words is 2 mod 32: lane b of the lane coder addresses team b,
                                         and the eight teams' entries fall into different banks (8-byte accesses too) */
};
static_assert((sizeof(w3_enc_team) / 4u) % 32u == 2u && sizeof(w3_enc_team) % 8u == 0u, "team stride: bank spread of the lane coder");

#define W3_ENC_TEAM_THREADS 96 /* per team: coder, model and writer warp */
/* Teams per CTA.  One team per CTA leaves an SM at three resident blocks (the
 * 48 KiB table copy is per CTA); with up to eight teams sharing one copy, 1024
 * blocks (1 GiB in 1 MiB blocks) are resident at once instead of running in
 * 2.3 waves, and the 21-24 warps of a CTA fill the issue slots a lone team of
 * three leaves idle. */
#define W3_ENC_TEAM_MAXB 8u

/* kLaneCoder: ONE coder warp for all the CTA's teams, lane b coding block b — the coder's 22
 * instructions per coded bit are then issued once per CTA instead of once per team (47 warp
 * instructions per coded bit and team made seven teams per SM pipe-bound: 134 cycles per bit;
 * the lone coder warp paces at ~60-70).  Warp 0 is the coder; B model warps and B writer warps follow.  The per-team rings and flags are unchanged; the coder warp waits until every
 * active lane's group is there. */
/* The coder warp has scheduler 0 to itself: worker k (model warps first, then writer warps) is warp
 * k + 1 + k / 3, the other warps = 0 mod 4 exit at once (as in the team decoder). */
#define W3_ENC_LANE_WARP_OF(k) ((k) + 1u + (k) / 3u)
#define W3_ENC_LANE_THREADS(teams) (32u * (W3_ENC_LANE_WARP_OF(2u * (teams) - 1u) + 1u))

template <bool kLaneCoder>
__global__ void __launch_bounds__(W3_ENC_TEAM_THREADS * W3_ENC_TEAM_MAXB, 1)
w3_encode_blocks_team(const uint8_t *__restrict__ in, uint64_t n_total, uint32_t block_size,
                      uint32_t n_blocks, uint8_t *__restrict__ scratch, uint64_t stride,
                      uint32_t *__restrict__ lens, uint32_t *__restrict__ err,
                      const w3_hashtab *__restrict__ tab, uint32_t n_teams) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  w3_smem_tab *st = reinterpret_cast<w3_smem_tab *>(smem_raw);
  w3_enc_team *t0 = reinterpret_cast<w3_enc_team *>(smem_raw + sizeof(w3_smem_tab));
  const uint32_t lane = threadIdx.x & 31u, wid = threadIdx.x >> 5;
  /* role: 0 coder, 1 model, 2 writer; lane coder: worker k = wid - 1 - wid / 4 for the warps that are not 0 mod 4 */
  const uint32_t worker = wid - 1u - (wid >> 2);
  const uint32_t team = kLaneCoder ? (wid == 0u ? 0u : worker % n_teams) : wid / 3u;
  const uint32_t warp = kLaneCoder ? (wid == 0u ? 0u : 1u + worker / n_teams) : wid % 3u;
  w3_enc_team *tm = t0 + team;

  w3_load_tab(st, tab);
  for (uint32_t i = threadIdx.x; i < n_teams * W3_NCTX; i += blockDim.x) t0[i / W3_NCTX].stats[i % W3_NCTX] = 0u;
  if (threadIdx.x < n_teams) {
    w3_enc_team *t = t0 + threadIdx.x;
    t->prod_done = 0; t->cons_done = 0; t->out_done = 0; t->wr_done = 0; t->abort_flag = 0; t->final_x2 = 0xFFFFFFFFu;
  }
  __syncthreads();

  /* one abort flag per CTA when the coder warp is shared: a wait that gives up must release everybody */
  const uint32_t abort_a = w3_saddr(kLaneCoder ? &t0->abort_flag : &tm->abort_flag);

  if (kLaneCoder && wid == 0u) {
    /* ---- lane coder: lane b runs the coder of team b (see the coder warp below for the step) ---- */
    const uint32_t blk_l = blockIdx.x * n_teams + lane;
    const bool has = lane < n_teams && blk_l < n_blocks;
    w3_enc_team *my = t0 + (has ? lane : 0u);
    const uint64_t off_l = (uint64_t)blk_l * block_size;
    const uint32_t n_l = has ? (uint32_t)(n_total - off_l < block_size ? n_total - off_l : block_size) : 0u;
    const uint32_t words_l = (n_l + 3u) >> 2;
    uint32_t max_words = words_l;
#pragma unroll
    for (uint32_t d = 16; d; d >>= 1) {
      const uint32_t o = __shfl_xor_sync(FULL, max_words, d);
      max_words = o > max_words ? o : max_words;
    }
    const uint32_t prod_l = w3_saddr(&my->prod_done), cons_l = w3_saddr(&my->cons_done);
    const uint32_t out_l = w3_saddr(&my->out_done), wr_l = w3_saddr(&my->wr_done), fx2_l = w3_saddr(&my->final_x2);
    const uint32_t ring_l = w3_saddr(&my->ring[0][0]), oring_l = w3_saddr(&my->oring[0][0]);
    w3_coder cw = w3_coder_new();
    const uint32_t ones = w3_lds(w3_saddr(&t0->final_x2)); /* 0xFFFFFFFF in a register ptxas cannot fold */
    auto coder_step = [&](uint32_t vj) -> uint2 {
      uint32_t flip, x1p, x2p;
      (void)w3_coder_step<false>(cw, vj & 1u, vj & 0xFFFF0000u, flip, x1p, x2p, ones);
      return make_uint2(x1p, x2p);
    };
    uint32_t ready = 0, room = W3_OUT_RING;
    for (uint32_t g = 0; g < max_words; g++) {
      const bool active = has && g < words_l;
      if (__any_sync(FULL, active && ready <= g)) { /* every active lane's group g must be there */
        uint32_t spins = 0;
        for (;;) {
          if (active) ready = w3_lds(prod_l);
          if (__all_sync(FULL, !active || (int32_t)(ready - g) > 0)) break;
          if (++spins > W3_SPIN_LIMIT || w3_lds(abort_a)) { w3_sts(abort_a, 1u); return; }
        }
      }
      if (__any_sync(FULL, active && g >= room)) { /* and room in every active lane's output ring */
        uint32_t spins = 0;
        for (;;) {
          if (active) room = w3_lds(wr_l) + W3_OUT_RING;
          if (__all_sync(FULL, !active || g < room)) break;
          if (++spins > W3_SPIN_LIMIT || w3_lds(abort_a)) { w3_sts(abort_a, 1u); return; }
        }
      }
      const uint32_t ga = ring_l + (g & (W3_ENC_RING - 1u)) * 128u;
      const uint32_t oa = oring_l + (g & (W3_OUT_RING - 1u)) * 256u;
      const uint32_t left = n_l * 8u - g * 32u;
      if (__all_sync(FULL, !active || left >= 32u)) {
        uint32_t v0 = w3_lds(ga), v1 = w3_lds(ga + 4u);
#pragma unroll
        for (uint32_t j = 0; j < 32u; j++) {
          const uint32_t vj = v0;
          v0 = v1;
          if (j + 2u < 32u) v1 = w3_lds(ga + (j + 2u) * 4u);
          const uint2 r = coder_step(vj);
          if (active) w3_sts64(oa + j * 8u, r);
        }
      } else {
        for (uint32_t j = 0; j < 32u; j++) { /* a short last group: bounds that emit nothing past its end */
          if (active) w3_sts64(oa + j * 8u, j < left ? coder_step(w3_lds(ga + j * 4u)) : make_uint2(0u, 0xFFFFFFFFu));
        }
      }
      if (active) {
        w3_sts(cons_l, g + 1u);
        if (g + 1u == words_l) w3_sts(fx2_l, cw.x1 + cw.w);
      }
      __threadfence_block();
      if (active) w3_sts(out_l, g + 1u);
    }
    return; /* an empty block never started its coder: final_x2 is still new_coder's (set above) */
  }

  if (kLaneCoder && ((wid & 3u) == 0u || worker >= 2u * n_teams)) return; /* scheduler 0 belongs to the coder */
  const uint32_t blk = blockIdx.x * n_teams + team;
  if (blk >= n_blocks) return;
  const uint64_t off = (uint64_t)blk * block_size;
  const uint32_t n = (uint32_t)(n_total - off < block_size ? n_total - off : block_size);
  const uint32_t n_words = (n + 3u) >> 2; /* groups of 32 coded bits */
  const uint32_t ring_a = w3_saddr(&tm->ring[0][0]) + lane * 4u;
  const uint32_t prod_a = w3_saddr(&tm->prod_done), cons_a = w3_saddr(&tm->cons_done);
  const uint32_t out_a = w3_saddr(&tm->out_done), wr_a = w3_saddr(&tm->wr_done);

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
    uint32_t room = W3_ENC_RING; /* groups the coder is known to have freed */
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
        uint32_t ctx = (w3_rank_dev(st, key, t) << W3_ALIGN_BITS) | (t & 7u);
        if (t == 0) ctx = W3_CTX0_RANKED;
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
        if (g >= room) {
          if (!w3_wait_gt(cons_a, g - W3_ENC_RING, abort_a)) return;
          room = w3_lds(cons_a) + W3_ENC_RING;
        }
        w3_sts(ring_a + (g & (W3_ENC_RING - 1u)) * 128u, active ? ((p << 16) | bit) : 0u);
        __syncwarp();
        if (lane == 0) {
          __threadfence_block();
          w3_sts(prod_a, g + 1u);
        }
        w0 = w1;
        w1 = w2;
      }
    }
    return;
  }

  if (!kLaneCoder && warp == 0) {
    /* ---- coder warp: the range chain of ac.encode (main.rs:107) and nothing
     * else, carried as (x1, width) with one clz per step (w3_coder_step).  A
     * lone warp pays 2.3-2.7 cycles per INSTRUCTION, on or off the chain
     * (profiles/r01_microbench1_coder_steps.txt), so per step it reads one ring
     * entry and publishes the two bounds as they are after narrowing, nothing
     * more: how many bits that emits and how many it leaves pending is worked
     * out by the writer warp, 32 steps per instruction.  The steady state is 32
     * steps of straight-line code. */
    w3_coder cw = w3_coder_new();
    uint32_t ones;
    /* 0xFFFFFFFF, read back from shared memory: anything ptxas can see through it folds into a
     * constant again, and then spends a move per use to get it into a register */
    w3_sts(w3_saddr(&tm->final_x2), 0xFFFFFFFFu);
    ones = w3_lds(w3_saddr(&tm->final_x2));
    auto coder_step = [&](uint32_t vj) -> uint2 {
      uint32_t flip, x1p, x2p;
      (void)w3_coder_step<false>(cw, vj & 1u, vj & 0xFFFF0000u, flip, x1p, x2p, ones);
      return make_uint2(x1p, x2p);
    };
    uint32_t ready = 0, room = W3_OUT_RING;
#ifdef W3_DEBUG
    long long dbg_wait_model = 0, dbg_wait_writer = 0, dbg_n_model = 0, dbg_n_writer = 0;
    const long long dbg_c0 = clock64();
#endif
    const uint32_t ring_u = w3_saddr(&tm->ring[0][0]);
    const uint32_t oring_u = w3_saddr(&tm->oring[0][0]);
    for (uint32_t g = 0; g < n_words; g++) {
      if (ready <= g) {
#ifdef W3_DEBUG
        const long long dbg_w0 = clock64();
#endif
        if (!w3_wait_gt(prod_a, g, abort_a)) return;
        ready = w3_lds(prod_a);
#ifdef W3_DEBUG
        dbg_wait_model += clock64() - dbg_w0; dbg_n_model++;
#endif
      }
      if (g >= room) {
#ifdef W3_DEBUG
        const long long dbg_w0 = clock64();
#endif
        if (!w3_wait_gt(wr_a, g - W3_OUT_RING, abort_a)) return;
        room = w3_lds(wr_a) + W3_OUT_RING;
#ifdef W3_DEBUG
        dbg_wait_writer += clock64() - dbg_w0; dbg_n_writer++;
#endif
      }
      const uint32_t ga = ring_u + (g & (W3_ENC_RING - 1u)) * 128u;
      const uint32_t oa = oring_u + (g & (W3_OUT_RING - 1u)) * 256u;
      const uint32_t left = n * 8u - g * 32u;
      if (left >= 32u) {
        uint32_t v0 = w3_lds(ga), v1 = w3_lds(ga + 4u);
#pragma unroll
        for (uint32_t j = 0; j < 32u; j++) {
          const uint32_t vj = v0;
          v0 = v1;
          if (j + 2u < 32u) v1 = w3_lds(ga + (j + 2u) * 4u);
          w3_sts64(oa + j * 8u, coder_step(vj));
        }
      } else {
        for (uint32_t j = 0; j < 32u; j++) /* past the end: bounds that emit nothing */
          w3_sts64(oa + j * 8u, j < left ? coder_step(w3_lds(ga + j * 4u)) : make_uint2(0u, 0xFFFFFFFFu));
      }
      w3_sts(cons_a, g + 1u);
      if (g + 1u == n_words) w3_sts(w3_saddr(&tm->final_x2), cw.x1 + cw.w);
#ifndef W3_UNSAFE_NO_CODER_FENCE /* experiment only */
      __threadfence_block();
#endif
      w3_sts(out_a, g + 1u);
    }
    if (n_words == 0) w3_sts(w3_saddr(&tm->final_x2), cw.x1 + cw.w);
#ifdef W3_DEBUG
    if (blk == 0 && lane == 0 && n_words) printf("[w3] coder: %.1f cycles/bit; waits for the model warp %lld (%.1f cycles/bit), for the writer warp %lld (%.1f cycles/bit)\n", (double)(clock64() - dbg_c0) / (n_words * 32.0), dbg_n_model, (double)dbg_wait_model / (n_words * 32.0), dbg_n_writer, (double)dbg_wait_writer / (n_words * 32.0));
#endif
    return;
  }

  /* ---- writer warp: ACWriter (io.rs:52-101) over the published steps --------
   * Lane i holds step i of a group.  Only steps with s > 0 emit anything (about
   * one in seven on text); the pending-bit count each of them meets is a
   * difference of prefix sums of k, so the warp walks just those steps. */
  uint32_t *dst = reinterpret_cast<uint32_t *>(scratch + (uint64_t)blk * stride);
  const uint32_t cap_words = (uint32_t)(stride >> 2);
  uint32_t ovf = 0;
  bool dead = false;
  auto store = [&](uint32_t idx, uint32_t v) {
    if (lane == 0) {
      if (idx < cap_words) dst[idx] = w3_bswap(v);
      else ovf = 1;
    }
  };
  w3_writer wr;
  w3_writer_init(wr);
  const uint32_t oring_l = w3_saddr(&tm->oring[0][0]) + lane * 8u;
  uint32_t have = 0;
  for (uint32_t g = 0; g < n_words; g++) {
    if (have <= g) {
      if (!w3_wait_gt(out_a, g, abort_a)) { dead = true; break; }
      have = w3_lds(out_a);
    }
    const uint2 e = w3_lds64(oring_l + (g & (W3_OUT_RING - 1u)) * 256u);
    __syncwarp();
    if (lane == 0) w3_sts(wr_a, g + 1u); /* entries are in registers now */
    /* e = this step's bounds after narrowing: s = their common prefix (emitted
     * plainly), k = the E3 shifts that follow (arithmetic_coder.rs:51-62) */
    uint32_t s_l, k_l;
    {
      uint32_t b1 = e.x, b2 = e.y;
      int si, ki;
      w3_ac_renorm(b1, b2, si, ki);
      s_l = (uint32_t)si;
      k_l = (uint32_t)ki;
    }
    /* inclusive prefix sum of k over the group */
    uint32_t pk = k_l;
#pragma unroll
    for (uint32_t d = 1; d < 32u; d <<= 1) {
      const uint32_t up = __shfl_up_sync(FULL, pk, d);
      if (lane >= d) pk += up;
    }
    const uint32_t total_k = __shfl_sync(FULL, pk, 31);
    uint32_t hm = __ballot_sync(FULL, s_l != 0u);
#ifdef W3_UNSAFE_NO_WRITER /* experiment only: is the writer warp the encoder's limit? */
    hm = 0;
#endif
    /* pending bits before the group's first emitting step come from wr.rev;
     * `base` = exclusive prefix of k at the previous emitting step */
    uint32_t base = 0;
    bool first = true;
    while (hm) {
      const uint32_t i = (uint32_t)__ffs((int)hm) - 1u;
      hm &= hm - 1u;
      const uint32_t pre_i = __shfl_sync(FULL, e.x, i);
      const uint32_t s_i = __shfl_sync(FULL, s_l, i);
      const uint32_t pk_i = __shfl_sync(FULL, pk, i);   /* inclusive */
      const uint32_t k_i = __shfl_sync(FULL, k_l, i);
      const uint32_t excl = pk_i - k_i;
      /* rev met by step i: what was pending after the previous emitting step
       * (its own k onwards), or the carried-in count plus everything so far */
      wr.rev = first ? wr.rev + excl : excl - base;
      w3_writer_step(wr, pre_i, (int)s_i, 0, store); /* leaves wr.rev == 0 */
      base = excl;
      first = false;
    }
    wr.rev = first ? wr.rev + total_k : total_k - base;
  }
  /* ac.flush — main.rs:111; the coder stored x2 before publishing its last group
   * (an empty block never started the coder: x2 is still new_coder's) */
  uint32_t x2 = 0xFFFFFFFFu;
  if (n_words) x2 = w3_lds(w3_saddr(&tm->final_x2));
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
    if (dead || w3_lds(abort_a)) atomicOr(err, W3_ERR_ABORT);
  }
}

/* ------------------------------------------------------------------------ */
/* decode                                                                   */

#define W3_SRING 128u /* stream words staged in shared memory for the decoder */
#define W3_BIT_BARS W3_CAND_RING /* one per candidate slot, at the slot's stride: the decoder derives both addresses from one offset */
#define W3_BAR_STRIDE (W3_NGUESS * 4u)
#define W3_NGUESS (1u << W3_SPEC_D)

#define W3_THI_ROW 264u
/* the speculators' copy of the threshold table: finer first level (W3_PB2) */
struct w3_dec_tab {
  uint64_t thr[8 * 256 + 8];          /* full thresholds, for the general search */
  uint32_t thi[8 * W3_THI_ROW];       /* their high words, rows padded with ~0: fixed lookups read W3_FIX_THR ahead */
  uint16_t cell2[8 * W3_NCELL2];
  uint16_t table[8];
  uint32_t max_bits;
  uint32_t pad[3];
};
static_assert(sizeof(w3_dec_tab) % 16 == 0, "smem table must keep 16-byte alignment");

struct w3_dec_team {
  uint32_t probs[W3_NCTX + 8];              /* Counter::p << 16 per slot; + read padding (fixed lookups read ahead) */
  uint32_t counts[W3_NCTX];                 /* c0 | c1 << 16 per slot; slot = alignment * 256 + rank (w3_dec_slot) */
  uint32_t cand[W3_CAND_RING][W3_NGUESS];   /* (Counter::p << 16) | tag for every guess of the last 6 bits;
                                               tag = (pos + 1) & 0xFFFF marks the entry as position pos's */
  uint2 sring[W3_SRING];                    /* entry i = big-endian stream words (i, i+1) */
  unsigned long long bars[W3_BIT_BARS][W3_BAR_STRIDE / 8u]; /* [s][0]: mbarrier of the decoded bits k with (k + 1) mod 16 = s —
                                               the decoder arrives, speculators sleep on it */
  uint2 pub;                                /* x = last 32 decoded bits, newest highest; y = bits decoded (one STS.64) */
  uint32_t abort_flag;
  uint32_t fed;                             /* stream words staged so far (feeder -> decoder) */
  uint32_t cons;                            /* decoder's current stream word (decoder -> feeder); ~0 = done */
  uint32_t ones_word;                       /* 0xFFFFFFFF, read back into a register */
  uint32_t done[8];                         /* per alignment a: positions = a mod 8 up to done[a] have their
                                               predecessor's counter update in place and may be predicted */
#ifdef W3_DEBUG
  uint32_t dbg_stamp[64];                   /* clock at which the decoder published bit count t (t & 63) */
#endif
};

/* warp 0 decodes and has scheduler 0 to itself (the other warps = 0 mod 4 exit
 * at once); the warps that are not a multiple of 4, in order: the speculators,
 * then the feeder */
#define W3_DEC_ROLE(warp) ((warp) - 1u - ((warp) >> 2))
#define W3_DEC_TEAM_WARPS (W3_SPEC_WARPS + 2u + W3_SPEC_WARPS / 3u)
#define W3_DEC_TEAM_THREADS (32u * W3_DEC_TEAM_WARPS)
static_assert(W3_DEC_ROLE(W3_DEC_TEAM_WARPS - 1u) == W3_SPEC_WARPS && (W3_DEC_TEAM_WARPS - 1u) % 4u != 0u,
              "the last warp must be the feeder");

/* counter slot of context (rank << 3) | alignment: alignment-major, so that the
 * 32 guesses of a warp (same alignment, scattered ranks) spread over the banks */
__device__ __forceinline__ uint32_t w3_dec_slot(uint32_t ctx) { return ((ctx & 7u) << 8) | (ctx >> 3); }

/* One guess's fixed-latency lookup, written so that two of them share a basic
 * block and interleave.  `sw` is the swizzled cell index, `khi` the key's high
 * word.  The cell entry gives the rank range [base, base + cnt]; the high words
 * of the next W3_FIX_THR thresholds AND the probability of every rank in the
 * range are loaded at once (the counter slots depend only on base), so the
 * chain is LDS -> LDS -> compare -> select.  Thresholds ascend globally, so
 * those past the cell compare greater by themselves.
 * bad = more than W3_FIX_THR thresholds in the cell, or a tie on the high
 * word: the caller falls back to the general (64-bit) search. */
struct w3_fix { uint32_t r, P, e; bool bad; };
/* the tables are constant once staged: plain (schedulable) loads by 32-bit
 * shared address — a generic pointer costs an S2R + LEA per use on this path */
__device__ __forceinline__ uint32_t w3_ldc16(uint32_t a) {
  uint16_t v;
  asm("ld.shared.u16 %0, [%1];" : "=h"(v) : "r"(a));
  return v;
}
__device__ __forceinline__ uint32_t w3_ldc32(uint32_t a) {
  uint32_t v;
  asm("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a));
  return v;
}
/* Nearly every cell of the fine first level holds no threshold at all, a few
 * hundred hold one, a handful up to six (w3_build_cell2) — but text keeps
 * hitting exactly those, so the fixed-latency lookup has to cover them.  It
 * issues all W3_FIX_THR threshold loads and W3_FIX_THR + 1 probability loads at
 * once, each predicated on the cell's count: a lane whose cell is empty loads
 * one probability and nothing else.  Fewer wavefronts matter: the speculators
 * kept the shared-memory pipe 44% busy, 39% of that bank-conflict replays
 * (profiles/r01_ncu_decode_team_v8.txt), which is latency for the decoder. */
#ifndef W3_DEC_FIX
#define W3_DEC_FIX 6 /* thresholds the speculators' fixed-latency lookup covers (<= W3_FIX_THR: the table rows' padding) */
#endif
__device__ __forceinline__ w3_fix w3_rank_fixed_sw(uint32_t khi, uint32_t a, uint32_t sw,
                                                   uint32_t thi_a, uint32_t cell2_a, uint32_t probs_a) {
  const uint32_t e = w3_ldc16(cell2_a + (a * W3_NCELL2 + sw) * 2u);
#ifdef W3_UNSAFE_NO_THR /* experiment only: how much do the speculators' loads cost the decoder? */
  const uint32_t base = e & 0xFFu, cnt = 0;
#else
  const uint32_t base = e & 0xFFu, cnt = e >> 8;
#endif
  const uint32_t ta = thi_a + (a * W3_THI_ROW + base) * 4u;
  /* slot of context (rank << 3) | a is a * 256 + rank (w3_dec_slot) */
  const uint32_t pa = probs_a + (a * 256u + base) * 4u;
  /* A probability word carries its own rank in the low byte (W3_PROB_RANK), so
   * selecting the probability selects the rank: no count of the compares.
   * Loads and compares beyond the cell's count are predicated off; what they
   * would have produced is never looked at. */
  uint32_t Pc[W3_DEC_FIX + 1];
  bool gt[W3_DEC_FIX];
  Pc[0] = w3_lds(pa);
  bool bad = cnt > (uint32_t)W3_DEC_FIX;
#pragma unroll
  for (int i = 0; i < W3_DEC_FIX; i++) {
    gt[i] = false;
#ifdef W3_LOADS_PREDICATED /* fewer wavefronts (shared-memory pipe 44 % -> 33 % busy), six more instructions: 121 instead of 116.5 cycles per bit */
    if (cnt > (uint32_t)i) {
#else
    { /* thresholds past the cell's are greater by themselves: rows are padded with ~0 */
#endif
      const uint32_t bi = w3_ldc32(ta + (uint32_t)i * 4u);
      Pc[i + 1] = w3_lds(pa + (uint32_t)(i + 1) * 4u);
#ifndef W3_P_AFTER_COMPARE
      /* an opaque use: left alone the compiler predicates this load on the compare
       * below (it is only needed if that holds), which puts a third shared-memory
       * round trip on the path to the publishing store */
      asm volatile("" : "+r"(Pc[i + 1]));
#endif
      gt[i] = khi > bi;
      bad |= khi == bi;
    }
  }
  /* thresholds ascend: the compares read 1..1 0..0, the answer is Pc[their
   * count] — picked by a tree, not a chain */
#if W3_DEC_FIX == 6
  const uint32_t Pa = gt[1] ? Pc[2] : (gt[0] ? Pc[1] : Pc[0]);
  const uint32_t Pb = gt[3] ? Pc[4] : Pc[3];
  const uint32_t Pd = gt[5] ? Pc[6] : Pc[5];
  const uint32_t P = gt[4] ? Pd : (gt[2] ? Pb : Pa);
#elif W3_DEC_FIX == 4
  const uint32_t Pa = gt[1] ? Pc[2] : (gt[0] ? Pc[1] : Pc[0]);
  const uint32_t Pb = gt[3] ? Pc[4] : Pc[3];
  const uint32_t P = gt[2] ? Pb : Pa;
#elif W3_DEC_FIX == 3
  const uint32_t Pa = gt[0] ? Pc[1] : Pc[0];
  const uint32_t Pb = gt[2] ? Pc[3] : Pc[2];
  const uint32_t P = gt[1] ? Pb : Pa;
#else
#error "write the select tree for this W3_DEC_FIX"
#endif
  w3_fix f;
  f.r = P & 0xFFu;
  f.P = P & 0xFFFF0000u;
  f.e = e;
  f.bad = bad;
  return f;
}

/* speculator's view of the decoded prefix.  pv = the decoder's published word
 * as last read: x = the newest 32 decoded bits (newest highest — the high word
 * of the hash key, as w3_key_of_history builds it), y = how many are decoded.
 * lo = the 32 bits before those as of `known` bits, i.e. the key's low word:
 * only the rare general search needs it, so it is kept up off the critical path. */
struct w3_known {
  uint2 pv;
  uint32_t lo, hi, known;
};
/* bring (hi, lo) up to pv; at most 32 new bits */
__device__ __forceinline__ void w3_known_absorb(w3_known &kn) {
  kn.lo = __funnelshift_rc(kn.lo, kn.hi, kn.pv.y - kn.known);
  kn.hi = kn.pv.x;
  kn.known = kn.pv.y;
}
/* Sleep until the decoder has decoded `need` bits, then read what it published.
 * The decoder arrives on an mbarrier per bit (bars[(bit + 1) & 15], one arrival per
 * phase), a speculator waits for the phase of bit need-1 with try_wait: the
 * hardware parks the warp instead of having it spin.  That matters more than
 * the wait's own latency: spinning speculators took issue slots and
 * shared-memory wavefronts from the working ones and from the decoder (16
 * spinning warps doubled every lookup's duration; nanosleep in a spin loop wakes
 * too late).  A speculator never waits for a bit 16 or more behind the decoder:
 * it waits for bit pos-6 while the decoder cannot pass bit pos-1 without its
 * candidates, and for bit pos while the decoder cannot pass bit pos+15 without
 * the candidates this warp computes next (W3_SPEC_WARPS <= 16) — so the phase
 * parity (bit index / 16, mod 2) is unambiguous.
 * Returns 0, or bit 31 set if the wait gave up (the caller's `late`). */
#define W3_WAIT_LIMIT (1u << 20)
__device__ __forceinline__ uint32_t w3_bits_wait(w3_known &kn, uint32_t need, uint32_t bars_a, uint32_t pub_a) {
  uint32_t got = 1u;
  if (need) {
    const uint32_t k = need - 1u;
    asm volatile(
        "{\n"
        ".reg .pred p, q;\n"
        ".reg .u32 cnt;\n"
        "mov.u32 cnt, 0;\n"
        "W3_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "@p bra.uni W3_DONE;\n"
        "add.u32 cnt, cnt, 1;\n"
        "setp.lt.u32 q, cnt, %3;\n"
        "@q bra.uni W3_WAIT;\n"
        "W3_DONE:\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(got)
        : "r"(bars_a + ((k + 1u) & (W3_BIT_BARS - 1u)) * W3_BAR_STRIDE), "r"((k / W3_BIT_BARS) & 1u), "n"(W3_WAIT_LIMIT)
        : "memory");
  }
  kn.pv = w3_lds64(pub_a);
  return got ? 0u : 0x80000000u;
}
/* what the decoder publishes as its bit count when it gives up: releases every wait */
#define W3_PUB_ABORT 0x7FFFFF00u

__global__ void __launch_bounds__(W3_DEC_TEAM_THREADS)
w3_decode_blocks_team(const uint8_t *__restrict__ comp, const uint64_t *__restrict__ offsets,
                      uint32_t n_blocks, uint32_t block_size, uint64_t n_total,
                      uint8_t *__restrict__ out, uint32_t *__restrict__ err,
                      const w3_hashtab *__restrict__ tab) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  w3_dec_tab *st = reinterpret_cast<w3_dec_tab *>(smem_raw);
  w3_dec_team *tm = reinterpret_cast<w3_dec_team *>(smem_raw + sizeof(w3_dec_tab));
  const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31u;

  {
    const uint4 *gt = reinterpret_cast<const uint4 *>(&tab->thr[0][0]);
    uint4 *dt = reinterpret_cast<uint4 *>(st->thr);
    for (uint32_t i = threadIdx.x; i < 8u * 256u * 8u / 16u; i += blockDim.x) dt[i] = gt[i];
    if (threadIdx.x < 8) st->thr[8 * 256 + threadIdx.x] = ~(uint64_t)0;
    for (uint32_t i = threadIdx.x; i < 8u * W3_THI_ROW; i += blockDim.x) {
      const uint32_t row = i / W3_THI_ROW, col = i % W3_THI_ROW;
      st->thi[i] = col < 255u ? (uint32_t)(__ldg(&tab->thr[row][col]) >> 32) : 0xFFFFFFFFu;
    }
    /* cell2 goes in swizzled: see w3_cell2_swizzle (bank conflicts) */
    const uint16_t *gc = &tab->cell2[0][0];
    for (uint32_t i = threadIdx.x; i < 8u * W3_NCELL2; i += blockDim.x)
      st->cell2[(i & ~(W3_NCELL2 - 1u)) | w3_cell2_swizzle(i & (W3_NCELL2 - 1u))] = __ldg(gc + i);
    if (threadIdx.x < 8) st->table[threadIdx.x] = tab->table[threadIdx.x];
    if (threadIdx.x == 0) st->max_bits = tab->max_bits;
  }
  for (uint32_t i = threadIdx.x; i < W3_NCTX + 8u; i += blockDim.x) {
    tm->probs[i] = 0x80000000u | (i & 0xFFu); /* Counter::new: p = 32768; low byte = the slot's rank (W3_PROB_RANK) */
    if (i < W3_NCTX) tm->counts[i] = 0u;
  }
  for (uint32_t i = threadIdx.x; i < W3_CAND_RING * W3_NGUESS; i += blockDim.x) (&tm->cand[0][0])[i] = 0u;
  if (threadIdx.x < 8) tm->done[threadIdx.x] = threadIdx.x;
  if (threadIdx.x < W3_BIT_BARS)
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(w3_saddr(&tm->bars[threadIdx.x][0])) : "memory");
  if (threadIdx.x == 0) { tm->ones_word = 0xFFFFFFFFu; tm->pub = make_uint2(0u, 0u); tm->abort_flag = 0; tm->fed = 0; tm->cons = 0; }
  __syncthreads();

  const uint32_t blk = blockIdx.x;
  /* scheduler 0 (warps 0, 4, 8, ...) belongs to the decoder */
  if (blk >= n_blocks || (warp != 0 && (warp & 3u) == 0u)) return;
  const uint64_t off = (uint64_t)blk * block_size;
  const uint32_t n = (uint32_t)(n_total - off < block_size ? n_total - off : block_size);
  const uint32_t nbits = n * 8u;
  /* One opaque base register: left to itself the compiler re-derives every
   * shared address from SR_CgaCtaId (an S2R, ~30 cycles) inside the loops. */
  uint32_t tm_a = w3_saddr(tm);
  asm volatile("mov.u32 %0, %0;" : "+r"(tm_a));
#define W3_TM_A(field) (tm_a + (uint32_t)offsetof(w3_dec_team, field))
  const uint32_t cand_a = W3_TM_A(cand);
  const uint32_t pub_a = W3_TM_A(pub), abort_a = W3_TM_A(abort_flag), bars_a = W3_TM_A(bars);
  const uint32_t probs_a = W3_TM_A(probs), counts_a = W3_TM_A(counts);
  const uint32_t sring_a = W3_TM_A(sring);
  const uint32_t fed_a = W3_TM_A(fed), cons_a = W3_TM_A(cons);

  if (warp != 0 && W3_DEC_ROLE(warp) == W3_SPEC_WARPS) {
    /* ---- feeder: stages the block's stream in shared memory, as big-endian
     * words, zero past its end (ACReader — io.rs:18-48), so that the decoder
     * reads stream bits with one LDS and no refill branch ------------------- */
    const uint8_t *sp = comp + offsets[blk];
    const int64_t slen = (int64_t)(offsets[blk + 1] - offsets[blk]);
    const uint32_t sh = (uint32_t)(reinterpret_cast<uintptr_t>(sp) & 3u);
    const uint32_t *ab = reinterpret_cast<const uint32_t *>(sp - sh);
    const uint32_t sel = ((sh + 3u) | ((sh + 2u) << 4) | ((sh + 1u) << 8) | (sh << 12));
    uint32_t fed = 0, spins = 0, seen = 0;
    for (;;) {
      const uint32_t c = w3_lds(cons_a);
      if (c == 0xFFFFFFFFu || w3_lds(abort_a)) break;
      if ((int32_t)(fed - c) > (int32_t)(W3_SRING - 40u)) { /* ring full */
        /* A full ring is no fault: highly compressible data decodes tens of millions of bits per
         * 1024 stream bits (p = 65535/65536 costs 2.2e-5 bits).  Only a decoder that has stopped
         * decoding is one, so the watchdog counts polls without progress in its published bit
         * count, not polls of the ring. */
        const uint32_t bits = w3_lds(pub_a + 4u);
        spins = bits == seen ? spins + 1u : 0u;
        seen = bits;
        if (spins > W3_SPIN_LIMIT) {
#ifdef W3_DEBUG
          if (lane == 0) printf("[w3] feeder timeout fed %u cons %u\n", fed, c);
#endif
          w3_sts(abort_a, 1u);
          break;
        }
        continue;
      }
      spins = 0;
      const uint32_t i = fed + lane;
      const int64_t rem = slen - (int64_t)i * 4;
      uint32_t v = 0;
      if (rem > 0) {
        v = __byte_perm(__ldg(ab + i), __ldg(ab + i + 1u), sel);
        if (rem < 4) v &= 0xFFFFFFFFu << (8 * (4 - (int)rem));
      }
      w3_sts(sring_a + (i & (W3_SRING - 1u)) * 8u, v);
      if (i) w3_sts(sring_a + ((i - 1u) & (W3_SRING - 1u)) * 8u + 4u, v);
      __syncwarp();
      fed += 32u;
      if (lane == 0) {
        __threadfence_block();
        w3_sts(fed_a, fed);
      }
    }
    return;
  }

  if (warp != 0) {
    /* ---- speculator j: positions j, j + W3_SPEC_WARPS, … -------------------- */
    const uint32_t j = W3_DEC_ROLE(warp);
    const uint32_t done_a = W3_TM_A(done);
    /* Guess index g (W3_SPEC_D bits) of bits pos-D..pos-1: bit 0 = the newest
     * (pos-1), so the two values of the newest bit are neighbours (the decoder
     * fetches them as one pair); the bits above it are the older ones in key
     * order (pos-2 highest) — exactly the top bits of the decoder's history
     * register, no reversal anywhere.  This lane owns guesses lane + 32 i. */
    uint32_t ghi[W3_GPL]; /* the guess as the top bits of the key's high word */
#pragma unroll
    for (uint32_t i = 0; i < W3_GPL; i++) {
      const uint32_t g = lane + 32u * i;
      ghi[i] = (((g & 1u) << (W3_SPEC_D - 1u)) | (g >> 1)) << (32u - W3_SPEC_D);
    }
    /* general search, for the rare guess the fixed-latency lookup declines */
    auto rank_slow = [&](uint64_t key, uint32_t pos, uint32_t e) -> uint32_t {
      uint32_t r = e & 0xFFu;
      const uint32_t cnt = e >> 8;
      const uint64_t *t = st->thr + (pos & 7u) * 256u + r;
      bool eq = false;
      for (uint32_t i = 0; i < cnt; i++) { r += key >= t[i] ? 1u : 0u; eq |= key == t[i]; }
      if (eq) r = w3_rank_of_hash(w3_achash_exact(w3_key_of_history(key), pos, st->table, (int)st->max_bits).hash);
      return r;
    };
    const uint32_t thi_a = tm_a - (uint32_t)sizeof(w3_dec_tab) + (uint32_t)offsetof(w3_dec_tab, thi);
    const uint32_t cell2_a = tm_a - (uint32_t)sizeof(w3_dec_tab) + (uint32_t)offsetof(w3_dec_tab, cell2);
    w3_known kn;
    kn.pv = make_uint2(0u, 0u);
    kn.lo = kn.hi = kn.known = 0;
#ifdef W3_DEBUG
    long long dbg_wait = 0, dbg_work = 0, dbg_n = 0, dbg_look = 0, dbg_pub = 0, dbg_fall = 0;
    long long dbg_l1 = 0, dbg_l2 = 0, dbg_over450 = 0, dbg_over550 = 0, dbg_latepoll = 0;
    const uint32_t dbg_stamp_a = W3_TM_A(dbg_stamp);
#endif
    /* Per position: predict (wait for bit pos-D-1, look up, publish), prepare
     * the counter's two possible updates, commit one when bit pos arrives.  The
     * body is straight-line code up to the publishing store: waits are uniform
     * PTX loops whose outcome is checked once per position (`late`). */
    for (uint32_t pos = j; pos < nbits; pos += W3_SPEC_WARPS) {
      const uint32_t a = pos & 7u;
      /* ---- predict, once pos-D bits are known ------------------------------- */
      const uint32_t need = pos > W3_SPEC_D ? pos - W3_SPEC_D : 0u;
#ifdef W3_DEBUG
      const long long dbg_t0 = clock64();
#endif
      uint32_t late = w3_bits_wait(kn, need, bars_a, pub_a);
      late |= kn.pv.y - need;
#ifdef W3_DEBUG
      const long long dbg_t1 = clock64();
      const uint32_t dbg_st = w3_lds(dbg_stamp_a + (need & 63u) * 4u);
      if (kn.pv.y > need) dbg_latepoll++;
#endif
      /* The counters of alignment a are adapted by whoever predicted the previous
       * position of that alignment — with 16 warps, another warp.  Its update is
       * in place long before it is needed here; `done` proves it (read BEFORE the
       * probabilities; written after them, behind a fence).  The check rides along
       * with the lookup, off its critical path: "not yet" is one more reason for
       * the fallback below. */
#ifdef W3_UNSAFE_NO_DONE /* experiment only: bounds what the done check costs */
      const bool notyet = false;
#else
      const bool notyet = (int32_t)(w3_lds(done_a + a * 4u) - pos) < 0;
#endif
      /* The fixed-latency lookup needs only the key's high word, which comes
       * straight from the published history: drop what is known beyond bit
       * pos-D-1 (for pos < D that is everything), make room for the guess. */
      const uint32_t extra = kn.pv.y + W3_SPEC_D - pos;
      const uint32_t kbhi = (kn.pv.x << (extra & 31u)) >> W3_SPEC_D;
      w3_fix f[W3_GPL];
      bool any_bad = pos == 0u || notyet;
#pragma unroll
      for (uint32_t i = 0; i < W3_GPL; i++) {
        const uint32_t khi = kbhi | ghi[i];
        f[i] = w3_rank_fixed_sw(khi, a, w3_cell2_swizzle(khi >> (32u - W3_PB2)), thi_a, cell2_a, probs_a);
        any_bad |= f[i].bad;
      }
      /* publish first: the tag in the low half makes each entry self-validating
       * (neither a flag nor a fence is needed), and a lane whose lookup declined
       * stores a tag that cannot match, so the decoder simply sees "not yet" */
      const uint32_t tag = (pos + 1u) & 0xFFFFu;
      const uint32_t slot_a = cand_a + (pos & (W3_CAND_RING - 1u)) * (W3_NGUESS * 4u) + lane * 4u;
#pragma unroll
      for (uint32_t i = 0; i < W3_GPL; i++)
        w3_sts(slot_a + i * 128u, f[i].P | (tag ^ (f[i].bad || pos == 0u || notyet ? 1u : 0u)));
#ifdef W3_DEBUG
      const long long dbg_t2 = clock64();
      if (__any_sync(FULL, any_bad)) dbg_fall++;
      if (pos > 64u) {
        const uint32_t l1 = (uint32_t)dbg_t1 - dbg_st, l2 = (uint32_t)dbg_t2 - dbg_st;
        dbg_l1 += l1; dbg_l2 += l2; dbg_over450 += l2 > 450u; dbg_over550 += l2 > 550u;
      }
#endif
      w3_known_absorb(kn);
      if (any_bad) { /* rare; the general search, and the entry published again */
        if (notyet) {
          uint32_t spins = 0;
          while ((int32_t)(w3_lds(done_a + a * 4u) - pos) < 0 && ++spins < W3_WAIT_LIMIT) {}
          if (spins >= W3_WAIT_LIMIT) late |= 0x80000000u;
        }
        /* key without the guessed bits: 64-bit history shifted as kbhi was */
        const uint64_t kb = ((((uint64_t)kn.hi << 32) | kn.lo) << (extra & 63u)) >> W3_SPEC_D;
#pragma unroll
        for (uint32_t i = 0; i < W3_GPL; i++) {
          if (f[i].bad) f[i].r = rank_slow(kb | ((uint64_t)ghi[i] << 32), pos, f[i].e);
          if (pos == 0u) f[i].r = W3_CTX0_RANKED >> W3_ALIGN_BITS;
          f[i].P = w3_lds(probs_a + w3_dec_slot((f[i].r << W3_ALIGN_BITS) | a) * 4u) & 0xFFFF0000u;
          w3_sts(slot_a + i * 128u, f[i].P | tag);
        }
      }
      /* ---- prepare the adaptation (main.rs:136 -> ordern_entropy.rs:32-34) --
       * per lane, for its own guess of the context: the counter and probability
       * after a 0 and after a 1.  The warp has nothing else to do until bit pos
       * is decoded. */
      uint32_t prep_sq[W3_GPL], prep_nv0[W3_GPL], prep_nv1[W3_GPL], prep_np0[W3_GPL], prep_np1[W3_GPL];
#pragma unroll
      for (uint32_t i = 0; i < W3_GPL; i++) {
        prep_sq[i] = w3_dec_slot((f[i].r << W3_ALIGN_BITS) | a) * 4u;
        const uint32_t old = w3_lds(counts_a + prep_sq[i]);
        prep_nv0[i] = w3_counter_update(old, 0u);
        prep_nv1[i] = w3_counter_update(old, 1u);
        /* both outcomes share the new total: one reciprocal (w3_counter_p_both) */
        uint32_t p0, p1;
        w3_counter_p_both(old, p0, p1);
        prep_np0[i] = p0 << 16;
        prep_np1[i] = p1 << 16;
      }
#ifdef W3_DEBUG
      dbg_wait += dbg_t1 - dbg_t0; dbg_work += clock64() - dbg_t1; dbg_n++; dbg_look += dbg_t2 - dbg_t1; dbg_pub += clock64() - dbg_t2;
#endif
      /* ---- commit: the guess that came true and the bit are known once bit pos
       * is decoded */
      late |= w3_bits_wait(kn, pos + 1u, bars_a, pub_a);
      late |= kn.pv.y - pos - 1u;
      {
        const uint32_t hb = kn.pv.x << ((kn.pv.y - pos - 1u) & 31u); /* bit pos on top */
        const bool bq = (int32_t)hb < 0;
        /* the one lane whose guess matches the history stores what it prepared — no
         * shuffles: this path is what the next position of the alignment, predicted
         * two decode steps after bit pos, waits for (it went from 56 % to 7 % of that
         * alignment's predicts finding `done` not yet set) */
#pragma unroll
        for (uint32_t i = 0; i < W3_GPL; i++) {
          if ((((hb << 1) ^ ghi[i]) >> (32u - W3_SPEC_D)) == 0u) {
            w3_sts(counts_a + prep_sq[i], bq ? prep_nv1[i] : prep_nv0[i]);
            w3_sts(probs_a + prep_sq[i], (bq ? prep_np1[i] : prep_np0[i]) | ((prep_sq[i] >> 2) & 0xFFu)); /* + the slot's rank (W3_PROB_RANK) */
          }
        }
        __threadfence_block(); /* the counter before the flag (without it: 1 cycle per bit faster) */
        w3_sts(done_a + a * 4u, pos + 8u);
      }
      /* a wait that gave up, or the decoder's abort mark (bit 30 of a difference) */
      if (__builtin_expect((late >> 30) != 0u, 0)) {
#ifdef W3_DEBUG
        if (lane == 0) printf("[w3] speculator warp %u gives up at pos %u (pub %u)\n", warp, pos, kn.pv.y);
#endif
        w3_sts(abort_a, 1u);
        return;
      }
    }
#ifdef W3_DEBUG
    if (blk == 0 && lane == 0) printf("[w3] B warp %u: from the decoder's publish: detect %.1f, published %.1f cycles; >450: %.3f, >550: %.3f; polled late %.3f\n", warp, (double)dbg_l1 / dbg_n, (double)dbg_l2 / dbg_n, (double)dbg_over450 / dbg_n, (double)dbg_over550 / dbg_n, (double)dbg_latepoll / dbg_n);
    if (blk == 0 && lane == 0) printf("[w3] B warp %u: positions %lld, avg wait-for-bit %.1f, avg spec+publish %.1f cycles (lookups %.1f, slot+publish %.1f, warp-level fallbacks %lld)\n", warp, dbg_n, (double)dbg_wait / dbg_n, (double)dbg_work / dbg_n, (double)dbg_look / dbg_n, (double)dbg_pub / dbg_n, dbg_fall);
#endif
    return;
  }

  /* ---- decoder: decode steps only (main.rs:131-140) -------------------------
   * The steady state is straight-line code with NO conditional branch per bit:
   * a branch (even a warp-uniform vote + predicated call, placed in the shadow
   * of the multiply) costs this lone warp about 40 cycles, against 88 for the
   * whole step (profiles/r01_microbench5_decoder_alone.txt).  So the loop runs
   * ATTEMPTS, 16 per trip.  An attempt decodes bit t iff the split point the
   * previous attempt computed was built from published candidates (`ok`);
   * otherwise the narrowing is predicated off, the range (already normalised)
   * renormalises to itself with a zero shift, nothing advances, and the attempt
   * amounts to re-fetching the candidates.  Either way it ends by fetching the
   * candidate pair of the next bit to decode and noting whether its tag was
   * valid.  Waiting for the speculators is thus "attempts that do nothing". */
  uint8_t *dst = out + off;
  uint32_t x1 = 0, x2 = 0xFFFFFFFFu, x = 0;
  uint32_t bp = 32;   /* stream bits consumed (x holds the first 32) */
  uint32_t rhist = 0; /* decoded bits, newest highest: what the speculators hash, the
                         candidate index (its top bits) and, bit-reversed, the output */
  uint32_t rlo = 0;   /* the 32 bits before those (output words complete mid-trip) */
  bool dead = false;
  uint32_t xmid = 0; /* split point of bit t minus x1 (lerp, arithmetic_coder.rs:81); valid iff miss == 0 */
  uint32_t miss = 1; /* tag mismatch bits of the candidate xmid was built from: the attempt decodes iff 0 */
  uint32_t t = 0;

  /* Every instruction of this lone warp costs 2-3 cycles, on or off the dependent
   * chain (profiles/r01_microbench1_coder_steps_v2.txt), so the attempt is written
   * for instruction count: the split point is kept as an offset from x1 (no
   * 64-bit addend for the multiply), the candidate address of a failed attempt
   * is simply the previous one, the bit's mbarrier lives at the stride of the
   * candidate slots so one offset serves both, x2's fill bits come through a
   * funnel shift from a register of ones that ptxas cannot fold. */
  const uint32_t not_lane0 = lane == 0u ? 0u : 1u; /* one arrival per bit: lane 0's */
  const uint32_t ones = w3_lds(W3_TM_A(ones_word));
  uint32_t pair_a = cand_a; /* position 0, no history */
  auto attempt = [&]() {
    /* ok: the attempt decodes bit t, and next needs the candidates of bit t+1
     * (slot (t+1) mod 16, indexed by bits t-1.., chosen by bit t); not ok: it
     * needs those of bit t again — the pair it fetched last time */
    const uint32_t so = ((t + 1u) * W3_BAR_STRIDE) & ((W3_CAND_RING - 1u) * W3_BAR_STRIDE);
    const uint32_t na = cand_a + so + (rhist >> (33u - W3_SPEC_D)) * 8u;
    const bool ok = miss == 0u;
    pair_a = ok ? na : pair_a;
#ifdef W3_FETCH_EARLY
    const uint2 pair_n = w3_lds64(pair_a);
#endif
    const uint2 sw = w3_lds64(sring_a + ((bp >> 5) & (W3_SRING - 1u)) * 8u);
    /* x <= x1 + m (arithmetic_coder.rs:81-82); x >= x1 always */
    const uint32_t bit = (x - x1) <= xmid ? 1u : 0u;
    if (ok && bit) x2 = x1 + xmid;
    if (ok && !bit) x1 = x1 + xmid + 1u;
    if (ok) {
      rlo = __funnelshift_r(rlo, rhist, 1);
      rhist = __funnelshift_r(rhist, bit, 1);
      t += 1u;
    }
    w3_sts64(pub_a, make_uint2(rhist, t));
#if !defined(W3_FETCH_EARLY) && defined(W3_FETCH_BEFORE_ARRIVE)
    const uint2 pair_n = w3_lds64(pair_a);
#endif
    /* wake whoever waits for the bit just decoded (release: after the word above) */
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.eq.u32 p, %1, 0;\n"
        "@p mbarrier.arrive.shared::cta.b64 _, [%0];\n"
        "}\n" ::"r"(bars_a + so), "r"(miss | not_lane0)
        : "memory");
#ifdef W3_DEBUG
    w3_sts(W3_TM_A(dbg_stamp) + (t & 63u) * 4u, (uint32_t)clock64());
#endif
    /* The candidates of the next bit are fetched as late as the renormalisation's
     * latency hides: the speculators publish just in time (their latency is about
     * five attempts), so an earlier fetch — at the attempt's start, or ahead of
     * the arrival, which delays it by a dozen cycles in the memory pipe's queue —
     * shortens the attempt by a dozen cycles and makes more attempts fail.
     * Measured every time the speculators got faster: late has stayed ahead
     * (108.6 against 109.6 cycles per bit at the last count,
     * profiles/r01_decode_team_experiments.txt). */
#if !defined(W3_FETCH_EARLY) && !defined(W3_FETCH_BEFORE_ARRIVE)
    const uint2 pair_n = w3_lds64(pair_a);
#endif
    /* x <<= nsh with stream bits shifted in; of the E3 flips only the last one
     * survives the shifts (arithmetic_coder.rs:98-102): flip iff k > 0 */
    const uint32_t fresh = __funnelshift_l(sw.y, sw.x, bp);
    int s, k;
    const int nsh = w3_ac_renorm_ones(x1, x2, s, k, ones);
    x = __funnelshift_l(fresh, x, (uint32_t)nsh) ^ ((0u - (uint32_t)k) & 0x80000000u);
    bp += (uint32_t)nsh;
    /* the newest decoded bit picks the entry; its own tag says whether it is
     * the published candidate of bit t (a speculator lane whose fixed lookup
     * declined publishes a non-matching tag first, the real entry later) */
    const uint32_t sel = ((int32_t)rhist < 0) ? pair_n.y : pair_n.x;
    xmid = __umulhi(x2 - x1, sel & 0xFFFF0000u); /* the split point minus x1 */
    miss = (sel ^ (t + 1u)) & 0xFFFFu;
  };
  /* make sure the staged stream reaches `words` beyond the current position */
  auto need_stream = [&](uint32_t words) -> bool {
    const uint32_t cur = bp >> 5;
    w3_sts(cons_a, cur);
    if ((int32_t)(w3_lds(fed_a) - (cur + words)) >= 0) return true;
    return w3_wait_gt(fed_a, cur + words - 1u, abort_a);
  };

  if (nbits) {
    if (!need_stream(3u)) dead = true;
    x = w3_lds(sring_a);
  }
#if defined(W3_DEBUG) || defined(W3_STATS) /* W3_STATS: the decoder's own counts only, no probes in the loops */
  long long dbg_a0 = clock64(), dbg_trips = 0;
#endif
  uint32_t stalled = 0;
  while (t < nbits && !dead) {
    /* 4 trips (64 attempts, at most 62 stream words) per stream-ring check; past
     * the block's last bit no candidates are ever published, so trailing attempts
     * of the last trip do nothing */
    if (!need_stream(66u)) { dead = true; break; }
    const uint32_t t_in = t;
#pragma unroll 1
    for (uint32_t trip = 0; trip < 64u / W3_ATTEMPTS_PER_TRIP; trip++) {
      const uint32_t t0 = t;
#pragma unroll
      for (uint32_t i = 0; i < W3_ATTEMPTS_PER_TRIP; i++) attempt();
      if ((t ^ t0) & 32u) { /* a word completed t - (t & ~31) < 16 bits ago */
        const uint32_t w = __funnelshift_l(rlo, rhist, t & 31u);
        if (lane == 0) *reinterpret_cast<uint32_t *>(dst + (t >> 5) * 4u - 4u) = w3_bswap(__brev(w));
      }
#if defined(W3_DEBUG) || defined(W3_STATS)
      dbg_trips++;
#endif
    }
    if (w3_lds(abort_a)) dead = true;
    stalled = t == t_in ? stalled + 1u : 0u;
    if (stalled > W3_SPIN_LIMIT / 64u) { w3_sts(abort_a, 1u); dead = true; }
  }
#if defined(W3_DEBUG) || defined(W3_STATS)
  if (blk == 0 && lane == 0) printf("[w3] A: %u bits, %.1f cycles/bit overall; %.1f cycles/attempt, %.3f attempts/bit\n", nbits, (double)(clock64() - dbg_a0) / nbits, (double)(clock64() - dbg_a0) / (16.0 * dbg_trips), 16.0 * dbg_trips / nbits);
#endif
  w3_sts(cons_a, 0xFFFFFFFFu); /* releases the feeder */
  if (dead) { /* giving up: release the speculators too */
    w3_sts64(pub_a, make_uint2(0u, W3_PUB_ABORT));
    __threadfence_block();
    /* they sleep on mbarriers no decoded bit will complete any more: one arrival on
     * each ends the phase they wait for; they wake, read the mark above and leave */
    if (lane == 0)
      for (uint32_t sl = 0; sl < W3_BIT_BARS; sl++)
        asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bars_a + sl * W3_BAR_STRIDE) : "memory");
  }
  if (lane == 0) {
    const uint32_t rem = n & 3u; /* bytes after the last whole word */
    const uint32_t hist = __brev(rhist); /* newest bit lowest */
    for (uint32_t b = 0; b < rem; b++) dst[(n & ~3u) + b] = (uint8_t)(hist >> (8u * (rem - 1u - b)));
    if (dead || w3_lds(abort_a)) atomicOr(err, W3_ERR_ABORT);
  }
}
