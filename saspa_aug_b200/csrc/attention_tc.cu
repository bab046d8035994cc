// tcgen05 / TMEM flash attention for sm_100a: O = softmax(Q K^T * scale) V per (batch, head).
//
// Serves the UNet / ControlNet transformer blocks of the denoising loop (self-attention over 4096 / 1024 /
// 256 / 64 image tokens with head_dim 40 / 80 / 160, cross-attention over 77 text tokens) and the CLIP
// towers (head_dim 64, causal).  Replaces F.scaled_dot_product_attention under diffusers'
// AttnProcessor2_0 (reference call site: pipe(**pipe_args), run_aug/run_aug.py:278).
//
// One CTA owns 256 queries of one (batch, head) as two 128-row tiles and sweeps the keys in tiles of
// BKV (128 for head_dim <= 64, else 64).  Twelve warps:
//   warp 0     TMA producer: Q once, then K_j / V_j into a STAGES-deep ring of 128B-swizzled smem tiles
//              (64-column panels straight out of the fused [b, t, 3c] QKV buffer -- nothing is repacked in HBM).
//   warp 1, 11 MMA issuers of query tile 0 / 1 (warp 1 also owns the 512 TMEM columns); one elected lane issues every tcgen05.mma:
//                S_i  = Q_i K_j^T          (SS: both operands K-major in smem)           -> TMEM
//                O_i += P_i V_j            (TS: P_i read from TMEM, V_j MN-major in smem) -> TMEM
//              issue order per tile and key tile j:  QK(i,j+1) then PV(i,j), so the next S is ready before the softmax warps
//              finish the current one; the two tiles run on independent barriers (see the issuer code).
//   warps 2-5  softmax of tile 0, warps 6-9 softmax of tile 1: one thread per query row (no shuffles):
//              tcgen05.ld the S row, release S, running max with lazy rescale (O is only rescaled in TMEM
//              when the max grows by more than 2^8), p = ex2(s*c - m), bf16 P packed two per column back
//              into TMEM with tcgen05.st; finally O / l -> global.
//              (An FMA-pipe exponential -- Cody-Waite split + cubic, rel. error 7.5e-5 -- is kept as a tuning option,
//              ACfg::POLY_EVERY: once the MMA issue and tile phasing were fixed it measured slower than the MUFU for every share.)
//   warp 10    (head_dim % 16 == 8 only) patches a column of ones into the zero-cost pad of each landed V tile, so
//              the P V MMA also accumulates the softmax denominator sum_j P_ij in O[:, D] -- no per-element adds.
// Padding is free: head_dim 40 runs as K = 48 (the Q pad chunk is zeroed in smem; K's pad columns then
// multiply zeros) and PV as N = 48 (the extra accumulator columns are never stored).
#include "tc_ptx.cuh"
#include "../../include/saspa_b200.h"

namespace {
using namespace tcx;

constexpr int ATC_THREADS = 384;
constexpr int QROWS = 128;

template <int D>
struct ACfg {
  static constexpr int KS = (D + 15) / 16;    // k-steps of Q K^T
  static constexpr int NPAN = (D + 63) / 64;  // 64-column (128 B) panels of the head dim
  static constexpr int ON = KS * 16;          // N of the P V MMA = O columns in TMEM
  static constexpr int BKV = D <= 64 ? 128 : 64;
  static constexpr int STAGES = D <= 128 ? 3 : 2;
  static constexpr int QPAN_BYTES = QROWS * 128;
  static constexpr int KPAN_BYTES = BKV * 128;
  static constexpr int Q_BYTES = 2 * NPAN * QPAN_BYTES;
  static constexpr int KV_STAGE_BYTES = NPAN * KPAN_BYTES;
  static constexpr int SMEM = Q_BYTES + 2 * STAGES * KV_STAGE_BYTES + 1024 /*align*/ + 256 /*barriers*/;
  static constexpr int P_OFF = 2 * BKV;  // TMEM columns: S_i at i*BKV, P_i at P_OFF + i*BKV/2, O_i at O_OFF + i*ON
  static constexpr int O_OFF = 3 * BKV;
  static constexpr bool ONES = (D % 16) != 0;  // a free pad column exists: row sums come out of the P V MMA
  // every POLY_EVERY-th exponential on the FMA pipe (ex2_poly), 0 = all on the MUFU.  Re-measured after the issue / phase fixes
  // (tools/attn_debug.py, -DSASPA_ATTN_POLY_SWEEP): d = 40 1.43 / 1.50 / 1.53 / 1.78 ms and d = 64 1.57 / 1.74 / 1.76 / 2.08 ms for
  // 0 / 6 / 4 / 2 -- the softmax warps are issue-bound, not MUFU-bound, so the emulation only adds instructions.
  static constexpr int POLY_EVERY = 0;
  static_assert(O_OFF + 2 * ON <= 512, "TMEM budget");
  static_assert(SMEM <= 227 * 1024, "smem budget");
};

struct AttnParams {
  __nv_bfloat16* o;
  int ldo;
  int heads, tq, tkv;
  float scale_log2;
  int causal;
  unsigned long long* trace;  // phase timers of one CTA (tools/attn_debug.py --trace); nullptr in production
  int debug;  // timing experiments only (tools/attn_debug.py): 1 = no exponentials, 2 = no P V MMAs, 4 = no Q K^T MMAs
};

template <int D, bool TRACE, int PE = ACfg<D>::POLY_EVERY>
__global__ void __launch_bounds__(ATC_THREADS, 1)
    attn_tc_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK, const __grid_constant__ CUtensorMap tmV,
                   const AttnParams p) {
  using C = ACfg<D>;
  constexpr int BKV = C::BKV, STAGES = C::STAGES, ON = C::ON, KS = C::KS, NPAN = C::NPAN;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
  uint8_t* sQ = smem;
  uint8_t* sK = sQ + C::Q_BYTES;
  uint8_t* sV = sK + STAGES * C::KV_STAGE_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sV + STAGES * C::KV_STAGE_BYTES);
  uint64_t* q_full = bars;
  uint64_t* q_ready = bars + 1;
  uint64_t* k_full = bars + 2;
  uint64_t* k_empty = k_full + STAGES;
  uint64_t* v_full = k_empty + STAGES;
  uint64_t* v_empty = v_full + STAGES;
  uint64_t* v_ready = v_empty + STAGES;  // V tile patched with the ones column (ONES only)
  uint64_t* s_full = v_ready + STAGES;   // [2]
  uint64_t* s_free = s_full + 2;
  uint64_t* p_full = s_free + 2;
  uint64_t* o_done = p_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(o_done + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int bh = blockIdx.y, b = bh / p.heads, h = bh % p.heads;
  const int q0 = blockIdx.x * (2 * QROWS);
  int n_tiles = (p.tkv + BKV - 1) / BKV;
  if (p.causal) n_tiles = min(n_tiles, (min(q0 + 2 * QROWS, p.tq) + BKV - 1) / BKV);
  constexpr bool PAD_Q = (D % 16) != 0;  // head_dim % 16 == 8: zero the pad chunk of Q in smem

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
    mbar_init(q_full, 1);
    mbar_init(q_ready, 256);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&k_full[s], 1);
      mbar_init(&k_empty[s], 2);  // both MMA issuers
      mbar_init(&v_full[s], 1);
      mbar_init(&v_empty[s], 2);
      mbar_init(&v_ready[s], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&s_full[i], 1);
      mbar_init(&s_free[i], 4);
      mbar_init(&p_full[i], 4);
      mbar_init(&o_done[i], 1);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      const int col0 = h * D;
      mbar_expect_tx(q_full, C::Q_BYTES);
      for (int i = 0; i < 2; ++i)
        for (int pn = 0; pn < NPAN; ++pn) tma_load_3d(&tmQ, sQ + (i * NPAN + pn) * C::QPAN_BYTES, q_full, col0 + 64 * pn, q0 + i * QROWS, b);
      for (int j = 0; j < n_tiles; ++j) {
        const int s = j % STAGES;
        const uint32_t par = ((j / STAGES) & 1) ^ 1;
        mbar_wait(&k_empty[s], par);
        mbar_expect_tx(&k_full[s], C::KV_STAGE_BYTES);
        for (int pn = 0; pn < NPAN; ++pn)
          tma_load_3d(&tmK, sK + s * C::KV_STAGE_BYTES + pn * C::KPAN_BYTES, &k_full[s], col0 + 64 * pn, j * BKV, b);
        mbar_wait(&v_empty[s], par);
        mbar_expect_tx(&v_full[s], C::KV_STAGE_BYTES);
        for (int pn = 0; pn < NPAN; ++pn)
          tma_load_3d(&tmV, sV + s * C::KV_STAGE_BYTES + pn * C::KPAN_BYTES, &v_full[s], col0 + 64 * pn, j * BKV, b);
      }
    }
  } else if (warp == 1 || warp == 11) {
    // ===================== MMA issuers: warp 1 serves query tile 0, warp 11 query tile 1 =====================
    // One issuer per tile, each on its own barriers (Q K^T of key tile j+1 as soon as the tile's softmax warps hold S_j in
    // registers, P V of key tile j as soon as P_j is in TMEM): the two tiles drift into opposite phases -- one on the MUFU, the
    // other on TMEM / ALU -- instead of contending for the same pipe in lockstep, and a tile's next S is never held back by
    // the other tile's softmax.  The hardware interleaves the two instruction streams in the tensor pipe; a K / V ring stage
    // is released when BOTH issuers have committed it (mbarrier count 2).
    // Each WHOLE warp runs its loop converged and one elected lane issues every tcgen05 instruction: under a divergent
    // `if (lane == 0)` ptxas wraps each UTCHMMA in an ELECT / R2UR / branch loop (~87 clk per MMA, measured with the phase
    // timers: issuing the 16 P V MMAs of a key tile took 1400 clk and bounded the d = 40 kernel).
    {
      const int it = warp == 1 ? 0 : 1;
      constexpr uint32_t idesc_qk = make_idesc(QROWS, BKV, 0);
      constexpr uint32_t idesc_pv = make_idesc(QROWS, ON, 1);
      auto issue_qk = [&](int i, int s) {
        const uint32_t d_tmem = tmem_base + i * BKV;
#pragma unroll
        for (int ks = 0; ks < KS; ++ks) {
          if (p.debug & 4) break;
          const int pn = ks >> 2, kk = ks & 3;
          const uint64_t a = make_smem_desc(smem_u32(sQ + (i * NPAN + pn) * C::QPAN_BYTES) + kk * 32);
          const uint64_t bd = make_smem_desc(smem_u32(sK + s * C::KV_STAGE_BYTES + pn * C::KPAN_BYTES) + kk * 32);
          if (elect_one()) tc_mma_bf16(d_tmem, a, bd, idesc_qk, ks != 0 ? 1u : 0u);
        }
        if (elect_one()) tc_commit(&s_full[i]);
      };
      const bool tr = TRACE && p.trace != nullptr && blockIdx.x == 1 && blockIdx.y == 3 && it == 0;
      long long tacc[8] = {0, 0, 0, 0, 0, 0, 0, 0}, tt0 = TRACE ? clock64() : 0;
#define ATC_TRACE(n)                    \
  if constexpr (TRACE) {                \
    if (tr) {                           \
      const long long tt1 = clock64();  \
      tacc[n] += tt1 - tt0;             \
      tt0 = tt1;                        \
    }                                   \
  }
      mbar_wait(PAD_Q ? q_ready : q_full, 0);
      tc_fence_after();
      mbar_wait(&k_full[0], 0);
      tc_fence_after();
      issue_qk(it, 0);
      if (elect_one()) tc_commit(&k_empty[0]);
      ATC_TRACE(0)
      for (int j = 0; j < n_tiles; ++j) {
        const int s = j % STAGES;
        if (j + 1 < n_tiles) {
          const int s1 = (j + 1) % STAGES;
          mbar_wait(&k_full[s1], ((j + 1) / STAGES) & 1);
          ATC_TRACE(1)
          mbar_wait(&s_free[it], j & 1);  // the tile's softmax warps hold S_j in registers
          tc_fence_after();
          ATC_TRACE(2)
          issue_qk(it, s1);
          if (elect_one()) tc_commit(&k_empty[s1]);
          ATC_TRACE(4)
        }
        mbar_wait(C::ONES ? &v_ready[s] : &v_full[s], (j / STAGES) & 1);
        ATC_TRACE(5)
        mbar_wait(&p_full[it], j & 1);
        tc_fence_after();
        ATC_TRACE(6)
        if (!(p.debug & 2)) {
#pragma unroll
          for (int ks = 0; ks < BKV / 16; ++ks) {
            const uint64_t bd = make_smem_desc_mn(smem_u32(sV + s * C::KV_STAGE_BYTES) + ks * 2048, C::KPAN_BYTES);
            if (elect_one())
              tc_mma_bf16_ts(tmem_base + C::O_OFF + it * ON, tmem_base + C::P_OFF + it * (BKV / 2) + ks * 8, bd, idesc_pv, (j > 0 || ks != 0) ? 1u : 0u);
          }
        }
        if (elect_one()) tc_commit(&o_done[it]);
        if (elect_one()) tc_commit(&v_empty[s]);
        ATC_TRACE(7)
      }
      if constexpr (TRACE) {
        if (tr && lane == 0)
          for (int n = 0; n < 8; ++n) p.trace[n] = (unsigned long long)tacc[n];
      }
    }
  } else if (warp == 10) {
    // ===================== V patcher: ones into pad column D of every landed V tile =====================
    if (C::ONES) {
      constexpr int ch = D / 8, pn = ch / 8, lc = ch % 8;  // 16-byte chunk holding column D
      for (int j = 0; j < n_tiles; ++j) {
        const int s = j % STAGES;
        mbar_wait(&v_full[s], (j / STAGES) & 1);
        uint8_t* tile = sV + s * C::KV_STAGE_BYTES + pn * C::KPAN_BYTES;
        for (int row = lane; row < BKV; row += 32) {
          // the whole chunk [D, D + 8) is pad: ones in column D, zeros after it (keeps the unused accumulator columns finite)
          *reinterpret_cast<uint4*>(tile + row * 128 + ((lc ^ (row & 7)) << 4)) = make_uint4(0x00003F80u, 0u, 0u, 0u);
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(&v_ready[s]);
      }
    }
  } else {
    // ===================== softmax / correction / epilogue (warps 2..9) =====================
    const int i = (warp - 2) >> 2;  // query tile of this warpgroup
    const int quad = warp & 3;      // TMEM lane quadrant this warp may access
    const int row = quad * 32 + lane;
    const int qi = q0 + i * QROWS + row;
    const uint32_t lane_base = tmem_base + ((uint32_t)(quad * 32) << 16);
    const uint32_t t_s = lane_base + i * BKV;
    const uint32_t t_p = lane_base + C::P_OFF + i * (BKV / 2);
    const uint32_t t_o = lane_base + C::O_OFF + i * ON;

    if (PAD_Q) {
      mbar_wait(q_full, 0);
      constexpr int ch = D / 8, pn = ch / 8, lc = ch % 8;
      uint8_t* dst = sQ + (i * NPAN + pn) * C::QPAN_BYTES + row * 128 + ((lc ^ (row & 7)) << 4);
      *reinterpret_cast<uint4*>(dst) = make_uint4(0, 0, 0, 0);
      fence_proxy_async_smem();
      mbar_arrive(q_ready);
    }

    const bool tr = TRACE && p.trace != nullptr && blockIdx.x == 1 && blockIdx.y == 3 && lane == 0 && (warp == 2 || warp == 6);
    long long tacc[8] = {0, 0, 0, 0, 0, 0, 0, 0}, tt0 = TRACE ? clock64() : 0;
    float m_used = -INFINITY;  // exponent reference (log2 domain); lags the true running max by at most 8
    float lsum[4] = {0.0f, 0.0f, 0.0f, 0.0f};  // unused when the P V MMA accumulates the row sum (C::ONES)
    for (int j = 0; j < n_tiles; ++j) {
      mbar_wait(&s_full[i], j & 1);
      tc_fence_after();
      ATC_TRACE(0)
      uint32_t su[BKV];
#pragma unroll
      for (int c = 0; c < BKV; c += 32) tc_ld32p(t_s + c, &su[c]);
      tc_wait_ld();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&s_free[i]);
      ATC_TRACE(1)

      const int kv0 = j * BKV;
      if (kv0 + BKV > p.tkv) {
#pragma unroll
        for (int c = 0; c < BKV; ++c)
          if (kv0 + c >= p.tkv) su[c] = 0xff800000u;
      }
      if (p.causal && kv0 + BKV - 1 > qi) {
#pragma unroll
        for (int c = 0; c < BKV; ++c)
          if (kv0 + c > qi) su[c] = 0xff800000u;
      }
      // row maximum: eight independent chains of three-input maxima (FMNMX3), 16 columns per round
      float mx[8];
#pragma unroll
      for (int c = 0; c < 8; ++c) mx[c] = fmaxf(__uint_as_float(su[c]), __uint_as_float(su[c + 8]));
#pragma unroll
      for (int c = 16; c < BKV; c += 16) {
#pragma unroll
        for (int e = 0; e < 8; ++e) mx[e] = fmax3(mx[e], __uint_as_float(su[c + e]), __uint_as_float(su[c + 8 + e]));
      }
      const float m_tile = fmax3(fmax3(mx[0], mx[1], mx[2]), fmax3(mx[3], mx[4], mx[5]), fmaxf(mx[6], mx[7])) * p.scale_log2;

      bool p_free = (j == 0);
      if (j == 0) {
        m_used = (m_tile == -INFINITY) ? 0.0f : m_tile;
      } else {
        const bool need = m_tile > m_used + 8.0f;
        if (__any_sync(0xffffffffu, need)) {
          // rare: rescale this tile's O accumulator in TMEM (after P V of the previous key tile has landed)
          mbar_wait(&o_done[i], (j - 1) & 1);
          tc_fence_after();
          p_free = true;
          const float f = need ? ex2_approx(m_used - m_tile) : 1.0f;
          if (need) {
            m_used = m_tile;
            if constexpr (!C::ONES) {
#pragma unroll
              for (int a = 0; a < 4; ++a) lsum[a] *= f;
            }
          }
#pragma unroll 1
          for (int c = 0; c < ON; c += 16) {
            uint32_t o[16];
            tc_ld16(t_o + c, o);
            tc_wait_ld();
#pragma unroll
            for (int e = 0; e < 16; ++e) o[e] = __float_as_uint(__uint_as_float(o[e]) * f);
            tc_st16(t_o + c, o);
          }
          tc_wait_st();
        }
      }

      // One-time phase offset: tile 1 enters its first exponential phase only when tile 0 has left its own, so the two tiles
      // start out in opposite phases (one on the MUFU, the other on TMEM / ALU) instead of halving each other's MUFU rate in
      // lockstep; nothing re-synchronises them afterwards (measured: -14 % on d = 40; a strict per-tile ping-pong was slower).
      if (j == 0 && i == 1) asm volatile("bar.sync 2, 256;" ::: "memory");
      ATC_TRACE(2)
      const float neg_m = -m_used;
      uint32_t pk[BKV / 2];
      if (p.debug & 1) {  // timing experiment: no exponentials at all
#pragma unroll
        for (int c = 0; c < BKV; c += 2)
          pk[c >> 1] = pack_bf16(fmaf(__uint_as_float(su[c]), p.scale_log2, neg_m), fmaf(__uint_as_float(su[c + 1]), p.scale_log2, neg_m));
      } else {
#pragma unroll
        for (int c = 0; c < BKV; c += 2) {
          const float x0 = fmaf(__uint_as_float(su[c]), p.scale_log2, neg_m);
          const float x1 = fmaf(__uint_as_float(su[c + 1]), p.scale_log2, neg_m);
          const float p0 = (PE > 0 && (c % (PE > 0 ? PE : 1)) == PE - 1) ? ex2_poly(x0) : ex2_approx(x0);
          const float p1 = (PE > 0 && ((c + 1) % (PE > 0 ? PE : 1)) == PE - 1) ? ex2_poly(x1) : ex2_approx(x1);
          if constexpr (!C::ONES) lsum[(c >> 1) & 3] += p0 + p1;
          pk[c >> 1] = pack_bf16(p0, p1);
        }
      }
      if (j == 0 && i == 0) asm volatile("bar.arrive 2, 256;" ::: "memory");
      ATC_TRACE(3)
      if (!p_free) {  // P V of the previous key tile must have consumed P_i
        mbar_wait(&o_done[i], (j - 1) & 1);
        tc_fence_after();
      }
      ATC_TRACE(4)
#pragma unroll
      for (int c = 0; c < BKV / 2; c += 32) tc_st32(t_p + c, &pk[c]);
      tc_wait_st();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&p_full[i]);
      ATC_TRACE(5)
    }
    if constexpr (TRACE) {
      if (tr)
        for (int n = 0; n < 8; ++n) p.trace[8 + 8 * i + n] = (unsigned long long)tacc[n];
    }

    // ---- epilogue: O / l -> global ----
    mbar_wait(&o_done[i], (n_tiles - 1) & 1);
    tc_fence_after();
    float l = (lsum[0] + lsum[1]) + (lsum[2] + lsum[3]);
    if constexpr (C::ONES) {
      uint32_t lu;
      tc_ld1(t_o + D, lu);  // column D of O = sum_j P_ij (ones column of V)
      tc_wait_ld();
      l = __uint_as_float(lu);
    }
    const float inv = l > 0.0f ? 1.0f / l : 0.0f;
    __nv_bfloat16* og = p.o + ((long long)b * p.tq + qi) * p.ldo + (long long)h * D;
    const bool valid = qi < p.tq;
#pragma unroll 1
    for (int c = 0; c < ON; c += 16) {
      uint32_t o[16];
      tc_ld16(t_o + c, o);
      tc_wait_ld();
      if (valid) {
#pragma unroll
        for (int g = 0; g < 2; ++g) {
          if (c + g * 8 < D) {
            uint4 u;
            u.x = pack_bf16(__uint_as_float(o[g * 8 + 0]) * inv, __uint_as_float(o[g * 8 + 1]) * inv);
            u.y = pack_bf16(__uint_as_float(o[g * 8 + 2]) * inv, __uint_as_float(o[g * 8 + 3]) * inv);
            u.z = pack_bf16(__uint_as_float(o[g * 8 + 4]) * inv, __uint_as_float(o[g * 8 + 5]) * inv);
            u.w = pack_bf16(__uint_as_float(o[g * 8 + 6]) * inv, __uint_as_float(o[g * 8 + 7]) * inv);
            *reinterpret_cast<uint4*>(og + c + g * 8) = u;
          }
        }
      }
      __syncwarp();
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
  }
}

// ------------------------------------------------------------------------------------------------
// Host side
// ------------------------------------------------------------------------------------------------
typedef CUresult (*PFN_tmapEncodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                        const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                        CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

PFN_tmapEncodeTiled encode_fn() {
  static PFN_tmapEncodeTiled fn = nullptr;
  if (!fn) {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess && qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_tmapEncodeTiled>(ptr);
  }
  return fn;
}

// bf16 [batch, rows, cols] with row stride ld (elements) and dense batch stride rows*ld; box = [1, box_rows, 64], 128B swizzle.
int encode_rows3d(CUtensorMap* tm, const void* base, int cols, int rows, int batch, long long ld, int box_rows) {
  PFN_tmapEncodeTiled enc = encode_fn();
  if (!enc) {
    saspa_set_error("cuTensorMapEncodeTiled entry point unavailable (no CUDA driver?)");
    return SASPA_ERR_DRIVER;
  }
  cuuint64_t gdim[3] = {(cuuint64_t)cols, (cuuint64_t)rows, (cuuint64_t)batch};
  cuuint64_t gstride[2] = {(cuuint64_t)ld * 2, (cuuint64_t)ld * 2 * rows};
  cuuint32_t box[3] = {64u, (cuuint32_t)box_rows, 1u};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    saspa_set_error("cuTensorMapEncodeTiled(attention cols=%d rows=%d batch=%d ld=%lld box_rows=%d) failed: %d", cols, rows, batch, ld, box_rows, (int)r);
    return SASPA_ERR_DRIVER;
  }
  return SASPA_OK;
}

int g_attn_debug = 0;

unsigned long long* g_attn_trace = nullptr;

int g_attn_poly = -1;  // tuning hook: -1 = ACfg<D>::POLY_EVERY, else every g_attn_poly-th exponential on the FMA pipe (0 = none)

template <int D, bool TRACE, int PE = ACfg<D>::POLY_EVERY>
int launch_tc_impl(const void* q, int ldq, const void* k, int ldk, const void* v, int ldv, void* o, int ldo, int batch, int heads, int tq, int tkv,
              float scale, int causal, cudaStream_t stream) {
  using C = ACfg<D>;
  static bool configured = false;
  if (!configured) {
    SASPA_CUDA(cudaFuncSetAttribute((attn_tc_kernel<D, TRACE, PE>), cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM));
    configured = true;
  }
  CUtensorMap tmQ, tmK, tmV;
  int rc;
  if ((rc = encode_rows3d(&tmQ, q, heads * D, tq, batch, ldq, QROWS))) return rc;
  if ((rc = encode_rows3d(&tmK, k, heads * D, tkv, batch, ldk, C::BKV))) return rc;
  if ((rc = encode_rows3d(&tmV, v, heads * D, tkv, batch, ldv, C::BKV))) return rc;
  AttnParams p;
  p.o = static_cast<__nv_bfloat16*>(o);
  p.ldo = ldo;
  p.heads = heads;
  p.tq = tq;
  p.tkv = tkv;
  p.scale_log2 = scale * 1.4426950408889634f;
  p.causal = causal;
  p.debug = g_attn_debug;
  p.trace = g_attn_trace;
  dim3 grid(ceil_div(tq, 2 * QROWS), batch * heads);
  attn_tc_kernel<D, TRACE, PE><<<grid, ATC_THREADS, C::SMEM, stream>>>(tmQ, tmK, tmV, p);
  SASPA_LAUNCH_CHECK();
  return SASPA_OK;
}

template <int D>
int launch_tc(const void* q, int ldq, const void* k, int ldk, const void* v, int ldv, void* o, int ldo, int batch, int heads, int tq, int tkv,
              float scale, int causal, cudaStream_t stream) {
  if (g_attn_trace != nullptr && (D == 40 || D == 128))  // the phase-timer build exists for the two profiled head dims only
    return launch_tc_impl<(D == 40 || D == 128) ? D : 40, true>(q, ldq, k, ldk, v, ldv, o, ldo, batch, heads, tq, tkv, scale, causal, stream);
#ifdef SASPA_ATTN_POLY_SWEEP  // tuning build only (tools/attn_debug.py --poly): the FMA-pipe share of the exponentials, d = 40 / 64
  if constexpr (D == 40 || D == 64) {
    switch (g_attn_poly) {
      case 0: return launch_tc_impl<D, false, 0>(q, ldq, k, ldk, v, ldv, o, ldo, batch, heads, tq, tkv, scale, causal, stream);
      case 2: return launch_tc_impl<D, false, 2>(q, ldq, k, ldk, v, ldv, o, ldo, batch, heads, tq, tkv, scale, causal, stream);
      case 3: return launch_tc_impl<D, false, 3>(q, ldq, k, ldk, v, ldv, o, ldo, batch, heads, tq, tkv, scale, causal, stream);
      case 4: return launch_tc_impl<D, false, 4>(q, ldq, k, ldk, v, ldv, o, ldo, batch, heads, tq, tkv, scale, causal, stream);
      case 6: return launch_tc_impl<D, false, 6>(q, ldq, k, ldk, v, ldv, o, ldo, batch, heads, tq, tkv, scale, causal, stream);
      case 8: return launch_tc_impl<D, false, 8>(q, ldq, k, ldk, v, ldv, o, ldo, batch, heads, tq, tkv, scale, causal, stream);
      default: break;
    }
  }
#endif
  return launch_tc_impl<D, false>(q, ldq, k, ldk, v, ldv, o, ldo, batch, heads, tq, tkv, scale, causal, stream);
}

}  // namespace

// Returns SASPA_ERR_UNSUPPORTED (without setting an error) when the shape has no tcgen05 instantiation; the
// caller (saspa_attention_bf16) then uses the mma.sync kernel.
int saspa_attention_tc(const void* q, int ldq, const void* k, int ldk, const void* v, int ldv, void* o, int ldo, int batch, int heads, int tq,
                       int tkv, int d, float scale, int causal, cudaStream_t stream) {
  if ((ldo % 8) != 0 || (reinterpret_cast<uintptr_t>(o) & 15) != 0 || scale <= 0.0f) return SASPA_ERR_UNSUPPORTED;
  switch (d) {
    case 40: return launch_tc<40>(q, ldq, k, ldk, v, ldv, o, ldo, batch, heads, tq, tkv, scale, causal, stream);
    case 64: return launch_tc<64>(q, ldq, k, ldk, v, ldv, o, ldo, batch, heads, tq, tkv, scale, causal, stream);
    case 80: return launch_tc<80>(q, ldq, k, ldk, v, ldv, o, ldo, batch, heads, tq, tkv, scale, causal, stream);
    case 128: return launch_tc<128>(q, ldq, k, ldk, v, ldv, o, ldo, batch, heads, tq, tkv, scale, causal, stream);
    case 160: return launch_tc<160>(q, ldq, k, ldk, v, ldv, o, ldo, batch, heads, tq, tkv, scale, causal, stream);
    default: return SASPA_ERR_UNSUPPORTED;
  }
}

// Timing-experiment hook (see AttnParams::debug); results are WRONG while it is non-zero.  Not part of the product API.
// Phase-timer hook: a device buffer of 24 u64 receives the clock64 sums of CTA (1, 3): [0..7] MMA issuer (prologue, wait K, wait
// S-free tile 0 / 1, QK issue, wait V, wait P, PV issue), [8..15] / [16..23] one softmax thread of tile 0 / 1 (wait S, TMEM load,
// max + rescale, exp + pack, wait P V, TMEM store).  nullptr disables it.  Not part of the product API.
extern "C" void saspa_attention_poly(int every) { g_attn_poly = every; }

extern "C" void saspa_attention_trace(void* dev_buf) { g_attn_trace = static_cast<unsigned long long*>(dev_buf); }

extern "C" int saspa_attention_debug(int flags) {
  const int prev = g_attn_debug;
  g_attn_debug = flags;
  return prev;
}
