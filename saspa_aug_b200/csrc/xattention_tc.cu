// tcgen05 / TMEM cross-attention over a SHORT key sequence (the 77 text tokens), persistent, K/V resident in shared memory.
//
// Serves attn2 of every BasicTransformerBlock of the UNet / ControlNet (diffusers models/attention.py; reference call
// pipe(**pipe_args), run_aug/run_aug.py:278): O = softmax(Q K^T * scale) V with Tq = 4096 / 1024 / 256 image tokens per (batch, head)
// and Tkv <= 128 keys.  The whole key sequence is ONE tile, so there is no online-softmax recurrence; the op is bound by streaming
// Q in and O out of HBM (2 x Tq x d x 2 B per head), and the kernel is organised around that:
//
//   * persistent CTAs (one per SM) take RUNS of R consecutive 128-query tiles of one (batch, head), round-robin over the order
//     [batch][query chunk][head]: at any moment neighbouring CTAs work on the heads of the SAME token rows, so every 640-byte token row
//     of Q is fetched from HBM once and its other heads hit L2 (with head-major ranges per CTA the same lines were fetched once per
//     head: ncu showed 663 MB read for 168 MB of Q).  K / V of the run's head are loaded once per run into a 2-deep ring and stay
//     resident while its query tiles stream through a Q ring (TMA, 128B swizzle, 64-column panels straight out of the projection
//     outputs -- nothing is repacked);
//   * warp 1 issues S = Q K^T (SS, N = keys rounded up to 16) and O = P V (TS: P read from TMEM, V MN-major in smem) for two tiles in
//     flight: S / P / O of tile parity i live in their own TMEM columns, P aliases the columns of the S it was computed from;
//   * one softmax warpgroup (warps 2-5, one thread per query row) does EVERY tile: tcgen05.ld S -> mask -> max -> ex2 -> bf16 P back
//     to TMEM; the 80 exponentials per row make it the busiest role (640 MUFU clocks per tile and SM), so nothing else runs in it;
//   * one epilogue warpgroup (warps 6-9) drains O once P V has landed: O / l -> bf16 -> a dense shared-memory tile -> ONE TMA store per
//     tile (full 32-byte sectors; a thread-per-row store would write half-used sectors of an HBM-bound kernel's larger stream); the row
//     sums l travel through shared memory.  S_i recycles as soon as P V has consumed P_i, O_i as soon as the epilogue has read it, so the
//     softmax of tile k + 1 overlaps the P V and the epilogue of tile k;
//   * head_dim 40 runs as K = 48: the pad chunk [40, 48) is zeroed in the resident K tile (warp 10), so whatever the Q panel holds there
//     (the next head's values) multiplies zeros; P V runs with N = 48 and the extra accumulator columns are never stored.
#include "tc_ptx.cuh"
#include "../../include/saspa_b200.h"

namespace {
using namespace tcx;

constexpr int XT_THREADS = 352;  // warp 0 TMA, warp 1 MMA, warps 2-5 softmax, warps 6-9 epilogue, warp 10 K patcher
constexpr int XQ = 128;          // query rows per tile

template <int D, int NS>
struct XCfg {
  static constexpr int KS = (D + 15) / 16;      // k-steps of Q K^T
  static constexpr int NPAN = (D + 63) / 64;    // 64-column panels of the head dim
  static constexpr int ON = KS * 16;            // N of P V = O columns in TMEM
  static constexpr bool PAD = (D % 16) != 0;
  static constexpr int QS = NPAN == 1 ? 4 : 2;  // Q ring depth
  static constexpr int KVS = (NS > 96 && NPAN == 2) ? 1 : 2;
  static constexpr int QPAN_BYTES = XQ * 128;
  static constexpr int KPAN_BYTES = NS * 128;
  static constexpr int Q_STAGE = NPAN * QPAN_BYTES;
  static constexpr int KV_STAGE = 2 * NPAN * KPAN_BYTES;  // K panels then V panels
  static constexpr int O_TILE = XQ * D * 2;               // dense [128][D] bf16 staging tile per softmax group
  static constexpr int O_TILE_AL = (O_TILE + 1023) / 1024 * 1024;
  static constexpr int SMEM = QS * Q_STAGE + KVS * KV_STAGE + 2 * O_TILE_AL + 1024 /*align*/ + 256 /*barriers*/ + 4 * XQ * 4 /*row sums*/;
  static constexpr int S_STRIDE = NS <= 96 ? 96 : 128;  // TMEM columns: S_i (and P_i) at i * S_STRIDE, O_i at 2 * S_STRIDE + i * ON
  static constexpr int O_OFF = 2 * S_STRIDE;
  static_assert(O_OFF + 2 * ON <= 512, "TMEM budget");
  static_assert(SMEM <= 227 * 1024, "smem budget");
  static_assert(NS % 16 == 0 && NS <= 128, "keys are processed as one tile of at most 128");
};

struct XParams {
  int heads, tq, tkv;
  int tiles_per_bh;    // 128-query tiles per (batch, head)
  int run;             // R: consecutive tiles of one head per run (divides tiles_per_bh)
  int chunks_per_bh;   // tiles_per_bh / R
  int total_runs;      // batch * chunks_per_bh * heads, ordered [batch][chunk][head]
  float scale_log2;
};

__device__ __forceinline__ void x_tma_store_3d(const CUtensorMap* tm, const void* src, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"(reinterpret_cast<uint64_t>(tm)),
               "r"(smem_u32(src)), "r"(c0), "r"(c1), "r"(c2)
               : "memory");
}

template <int D, int NS>
__global__ void __launch_bounds__(XT_THREADS, 1)
    xattn_tc_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK, const __grid_constant__ CUtensorMap tmV,
                    const __grid_constant__ CUtensorMap tmO, const XParams p) {
  using C = XCfg<D, NS>;
  constexpr int QS = C::QS, KVS = C::KVS, NPAN = C::NPAN, KS = C::KS, ON = C::ON;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
  uint8_t* sQ = smem;
  uint8_t* sKV = sQ + QS * C::Q_STAGE;
  uint8_t* sO = sKV + KVS * C::KV_STAGE;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sO + 2 * C::O_TILE_AL);
  uint64_t* q_full = bars;             // [QS]
  uint64_t* q_empty = q_full + 4;      // [QS]
  uint64_t* kv_full = q_empty + 4;     // [KVS]
  uint64_t* kv_empty = kv_full + 2;    // [KVS]
  uint64_t* k_ready = kv_empty + 2;    // [KVS] (PAD only: pad chunk of K zeroed)
  uint64_t* s_full = k_ready + 2;      // [2]
  uint64_t* p_full = s_full + 2;       // [2]
  uint64_t* o_full = p_full + 2;       // [2]
  uint64_t* t_free = o_full + 2;       // [2] O_i drained by the epilogue
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(t_free + 2);
  uint64_t* l_full = bars + 24;                     // [4] row sums of tile k published in slot k & 3
  float* sL = reinterpret_cast<float*>(bars + 32);  // [4][128] softmax denominators of the tiles in flight

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // local tile k of this CTA = tile (k % R) of its run number k / R; run g = (k / R) * gridDim.x + blockIdx.x
  const int my_runs = (p.total_runs - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
  const int n_my = my_runs * p.run;
  auto decode = [&](int k, int& b, int& h, int& qt) {
    const int lr = k / p.run, r = k - lr * p.run;
    const int g = lr * (int)gridDim.x + (int)blockIdx.x;
    h = g % p.heads;
    const int t = g / p.heads;
    const int qc = t % p.chunks_per_bh;
    b = t / p.chunks_per_bh;
    qt = qc * p.run + r;
    return r == 0;  // first tile of a run: a new head's K / V
  };

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
    tma_prefetch_desc(&tmO);
    for (int s = 0; s < QS; ++s) {
      mbar_init(&q_full[s], 1);
      mbar_init(&q_empty[s], 1);
    }
    for (int s = 0; s < KVS; ++s) {
      mbar_init(&kv_full[s], 1);
      mbar_init(&kv_empty[s], 1);
      mbar_init(&k_ready[s], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&s_full[i], 1);
      mbar_init(&p_full[i], 4);
      mbar_init(&o_full[i], 1);
      mbar_init(&t_free[i], 4);
    }
    for (int i = 0; i < 4; ++i) mbar_init(&l_full[i], 4);
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
    // ===================== TMA producer (converged warp, one elected lane per instruction) =====================
    int kvn = -1;
    for (int k = 0; k < n_my; ++k) {
      int b, h, qt;
      const bool new_head = decode(k, b, h, qt);
      const int col0 = h * D;
      if (new_head) {
        ++kvn;
        const int st = kvn % KVS;
        mbar_wait(&kv_empty[st], ((kvn / KVS) & 1) ^ 1);
        uint8_t* dst = sKV + st * C::KV_STAGE;
        if (elect_one()) {
          mbar_expect_tx(&kv_full[st], C::KV_STAGE);
          for (int pn = 0; pn < NPAN; ++pn) {
            tma_load_3d(&tmK, dst + pn * C::KPAN_BYTES, &kv_full[st], col0 + 64 * pn, 0, b);
            tma_load_3d(&tmV, dst + (NPAN + pn) * C::KPAN_BYTES, &kv_full[st], col0 + 64 * pn, 0, b);
          }
        }
      }
      const int s = k % QS;
      mbar_wait(&q_empty[s], ((k / QS) & 1) ^ 1);
      if (elect_one()) {
        mbar_expect_tx(&q_full[s], C::Q_STAGE);
        for (int pn = 0; pn < NPAN; ++pn) tma_load_3d(&tmQ, sQ + s * C::Q_STAGE + pn * C::QPAN_BYTES, &q_full[s], col0 + 64 * pn, qt * XQ, b);
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    constexpr uint32_t idesc_qk = make_idesc(XQ, NS, 0);
    constexpr uint32_t idesc_pv = make_idesc(XQ, ON, 1);
    int kvn = -1;
    int st_of[2] = {0, 0};  // K/V stage of the tile in flight on parity i
    auto issue_pv = [&](int k) {
      const int i = k & 1;
      mbar_wait(&p_full[i], (k >> 1) & 1);
      mbar_wait(&t_free[i], ((k >> 1) & 1) ^ 1);  // the epilogue has drained O_i of tile k - 2
      tc_fence_after();
      const uint32_t v_base = smem_u32(sKV + st_of[i] * C::KV_STAGE + NPAN * C::KPAN_BYTES);
#pragma unroll
      for (int ks = 0; ks < NS / 16; ++ks) {
        const uint64_t bd = make_smem_desc_mn(v_base + ks * 2048, C::KPAN_BYTES);
        if (elect_one()) tc_mma_bf16_ts(tmem_base + C::O_OFF + i * ON, tmem_base + i * C::S_STRIDE + ks * 8, bd, idesc_pv, ks != 0 ? 1u : 0u);
      }
      if (elect_one()) tc_commit(&o_full[i]);
    };
    for (int k = 0; k < n_my; ++k) {
      const int i = k & 1;
      bool pv_issued = false;
      int release = -1;  // K / V stage whose last reader is tile k - 1's P V
      if (k % p.run == 0) {
        // new run = new head.  The previous head's stage may be refilled once every MMA that reads it has completed; its last reader is
        // tile k - 1's P V.  With a two-deep ring the new head already sits in the other stage, so that P V keeps its usual place (after
        // this tile's Q K^T) and the stage is released right behind it; with a one-deep ring the producer cannot load the new head
        // before the release, so the P V has to go first.
        if (k > 0) {
          release = st_of[(k - 1) & 1];
          if (KVS == 1) {
            issue_pv(k - 1);
            pv_issued = true;
            if (elect_one()) tc_commit(&kv_empty[release]);
            release = -1;
          }
        }
        ++kvn;
        mbar_wait(C::PAD ? &k_ready[kvn % KVS] : &kv_full[kvn % KVS], (kvn / KVS) & 1);
        tc_fence_after();
      }
      st_of[i] = kvn % KVS;
      const int s = k % QS;
      mbar_wait(&q_full[s], (k / QS) & 1);
      mbar_wait(&o_full[i], ((k >> 1) & 1) ^ 1);  // P V of tile k - 2 has consumed P_i, which aliases S_i
      tc_fence_after();
      const uint32_t k_base = smem_u32(sKV + st_of[i] * C::KV_STAGE);
#pragma unroll
      for (int ks = 0; ks < KS; ++ks) {
        const int pn = ks >> 2, kk = ks & 3;
        const uint64_t a = make_smem_desc(smem_u32(sQ + s * C::Q_STAGE + pn * C::QPAN_BYTES) + kk * 32);
        const uint64_t bd = make_smem_desc(k_base + pn * C::KPAN_BYTES + kk * 32);
        if (elect_one()) tc_mma_bf16(tmem_base + i * C::S_STRIDE, a, bd, idesc_qk, ks != 0 ? 1u : 0u);
      }
      if (elect_one()) tc_commit(&s_full[i]);
      if (elect_one()) tc_commit(&q_empty[s]);
      if (k > 0 && !pv_issued) issue_pv(k - 1);  // Q K^T of tile k is already queued in front of it: S_k is ready when the softmax needs it
      if (release >= 0 && elect_one()) tc_commit(&kv_empty[release]);
    }
    if (n_my > 0) issue_pv(n_my - 1);
  } else if (warp == 10) {
    // ===================== K patcher: zero the pad chunk [D, D + 8) of every landed K tile (head_dim % 16 == 8) =====================
    if (C::PAD) {
      constexpr int ch = D / 8, pn = ch / 8, lc = ch % 8;
      int kvn = -1;
      for (int k = 0; k < n_my; k += p.run) {
        ++kvn;
        const int st = kvn % KVS;
        mbar_wait(&kv_full[st], (kvn / KVS) & 1);
        uint8_t* tile = sKV + st * C::KV_STAGE + pn * C::KPAN_BYTES;
        for (int row = lane; row < NS; row += 32) *reinterpret_cast<uint4*>(tile + row * 128 + ((lc ^ (row & 7)) << 4)) = make_uint4(0u, 0u, 0u, 0u);
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(&k_ready[st]);
      }
    }
  } else if (warp < 6) {
    // ===================== softmax warpgroup (warps 2-5): every tile =====================
    const int quad = warp & 3;  // TMEM lane quadrant this warp may access
    const int row = quad * 32 + lane;
    const uint32_t lane_base = tmem_base + ((uint32_t)(quad * 32) << 16);
    for (int k = 0; k < n_my; ++k) {
      const int i = k & 1;
      const uint32_t t_s = lane_base + i * C::S_STRIDE;
      mbar_wait(&s_full[i], (k >> 1) & 1);
      tc_fence_after();
      uint32_t su[NS];
#pragma unroll
      for (int c = 0; c < NS; c += 16) tc_ld16(t_s + c, &su[c]);
      tc_wait_ld();
      // columns past the key count hold Q . 0 = 0: mask them out (77 keys in an 80-wide tile: only the last chunk can be affected)
      if (p.tkv > NS - 16) {
#pragma unroll
        for (int c = NS - 16; c < NS; ++c)
          if (c >= p.tkv) su[c] = 0xff800000u;
      } else {
#pragma unroll
        for (int c = 0; c < NS; ++c)
          if (c >= p.tkv) su[c] = 0xff800000u;
      }
      float mx[8];
#pragma unroll
      for (int c = 0; c < 8; ++c) mx[c] = fmaxf(__uint_as_float(su[c]), __uint_as_float(su[c + 8]));
#pragma unroll
      for (int c = 16; c < NS; c += 16) {
#pragma unroll
        for (int e = 0; e < 8; ++e) mx[e] = fmax3(mx[e], __uint_as_float(su[c + e]), __uint_as_float(su[c + 8 + e]));
      }
      const float neg_m = -fmax3(fmax3(mx[0], mx[1], mx[2]), fmax3(mx[3], mx[4], mx[5]), fmaxf(mx[6], mx[7])) * p.scale_log2;
      float ls[4] = {0.0f, 0.0f, 0.0f, 0.0f};
      uint32_t pk[NS / 2];
#pragma unroll
      for (int c = 0; c < NS; c += 2) {
        const float p0 = ex2_approx(fmaf(__uint_as_float(su[c]), p.scale_log2, neg_m));
        const float p1 = ex2_approx(fmaf(__uint_as_float(su[c + 1]), p.scale_log2, neg_m));
        ls[(c >> 1) & 3] += p0 + p1;
        pk[c >> 1] = pack_bf16(p0, p1);
      }
      // P overwrites the columns of the S it came from (every thread holds its whole S row in registers by now)
#pragma unroll
      for (int c = 0; c < NS / 2; c += 8) {
        asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%8], {%0, %1, %2, %3, %4, %5, %6, %7};" ::"r"(pk[c]), "r"(pk[c + 1]), "r"(pk[c + 2]),
                     "r"(pk[c + 3]), "r"(pk[c + 4]), "r"(pk[c + 5]), "r"(pk[c + 6]), "r"(pk[c + 7]), "r"(t_s + c)
                     : "memory");
      }
      // the row sum rides to the epilogue warpgroup through shared memory (slot k & 3: tile k + 4 cannot get here before the epilogue of
      // tile k is over -- its Q K^T waits for P V of k + 2, which waits for that epilogue); the arrive on l_full publishes it.  (The
      // epilogue must NOT wait on p_full: that barrier can run a whole phase ahead of a lagging epilogue, whose parity wait would then
      // never succeed.)
      sL[(k & 3) * XQ + row] = (ls[0] + ls[1]) + (ls[2] + ls[3]);
      tc_wait_st();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(&l_full[k & 3]);
        mbar_arrive(&p_full[i]);
      }
    }
  } else {
    // ===================== epilogue warpgroup (warps 6-9): O / l -> bf16 -> dense smem tile -> one TMA store per tile =====================
    const int quad = warp & 3;
    const int row = quad * 32 + lane;
    const uint32_t lane_base = tmem_base + ((uint32_t)(quad * 32) << 16);
    const bool leader = warp == 6 && lane == 0;
    for (int k = 0; k < n_my; ++k) {
      const int i = k & 1;
      int b, h, qt;
      decode(k, b, h, qt);
      const uint32_t t_o = lane_base + C::O_OFF + i * ON;
      uint8_t* stage = sO + i * C::O_TILE_AL;
      if (leader) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");  // the store of tile k - 2 has drained this staging tile
      asm volatile("bar.sync 1, 128;" ::: "memory");
      mbar_wait(&l_full[k & 3], (k >> 2) & 1);  // acquires the row sums
      mbar_wait(&o_full[i], (k >> 1) & 1);
      tc_fence_after();
      const float inv = 1.0f / sL[(k & 3) * XQ + row];  // the row maximum contributes ex2(0) = 1: never zero
      uint8_t* orow = stage + row * (D * 2);
#pragma unroll
      for (int c = 0; c < ON; c += 16) {
        uint32_t o[16];
        tc_ld16(t_o + c, o);
        tc_wait_ld();
#pragma unroll
        for (int g = 0; g < 2; ++g) {
          if (c + g * 8 < D) {
            uint4 u;
            u.x = pack_bf16(__uint_as_float(o[g * 8 + 0]) * inv, __uint_as_float(o[g * 8 + 1]) * inv);
            u.y = pack_bf16(__uint_as_float(o[g * 8 + 2]) * inv, __uint_as_float(o[g * 8 + 3]) * inv);
            u.z = pack_bf16(__uint_as_float(o[g * 8 + 4]) * inv, __uint_as_float(o[g * 8 + 5]) * inv);
            u.w = pack_bf16(__uint_as_float(o[g * 8 + 6]) * inv, __uint_as_float(o[g * 8 + 7]) * inv);
            *reinterpret_cast<uint4*>(orow + (c + g * 8) * 2) = u;
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&t_free[i]);  // O_i may be overwritten by P V of tile k + 2
      fence_proxy_async_smem();
      asm volatile("bar.sync 1, 128;" ::: "memory");
      if (leader) {
        x_tma_store_3d(&tmO, stage, h * D, qt * XQ, b);  // rows past tq are clipped by the tensor map
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      }
    }
    if (leader) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
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

PFN_tmapEncodeTiled x_encode_fn() {
  static PFN_tmapEncodeTiled fn = nullptr;
  if (!fn) {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess && qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_tmapEncodeTiled>(ptr);
  }
  return fn;
}

// bf16 [batch, rows, cols] with row stride ld (elements), dense batch stride rows * ld; box = [1, box_rows, box_cols]
int x_encode_rows3d(CUtensorMap* tm, const void* base, int cols, int rows, int batch, long long ld, int box_rows, int box_cols, CUtensorMapSwizzle swz) {
  PFN_tmapEncodeTiled enc = x_encode_fn();
  if (!enc) {
    saspa_set_error("cuTensorMapEncodeTiled entry point unavailable (no CUDA driver?)");
    return SASPA_ERR_DRIVER;
  }
  cuuint64_t gdim[3] = {(cuuint64_t)cols, (cuuint64_t)rows, (cuuint64_t)batch};
  cuuint64_t gstride[2] = {(cuuint64_t)ld * 2, (cuuint64_t)ld * 2 * rows};
  cuuint32_t box[3] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows, 1u};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, swz,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    saspa_set_error("cuTensorMapEncodeTiled(cross-attention cols=%d rows=%d batch=%d ld=%lld box=%dx%d) failed: %d", cols, rows, batch, ld, box_rows,
                    box_cols, (int)r);
    return SASPA_ERR_DRIVER;
  }
  return SASPA_OK;
}

template <int D, int NS>
int launch_xtc(const void* q, int ldq, const void* k, int ldk, const void* v, int ldv, void* o, int ldo, int batch, int heads, int tq, int tkv,
               float scale, cudaStream_t stream) {
  using C = XCfg<D, NS>;
  static bool configured = false;
  if (!configured) {
    SASPA_CUDA(cudaFuncSetAttribute((xattn_tc_kernel<D, NS>), cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM));
    configured = true;
  }
  CUtensorMap tmQ, tmK, tmV, tmO;
  int rc;
  if ((rc = x_encode_rows3d(&tmQ, q, heads * D, tq, batch, ldq, XQ, 64, CU_TENSOR_MAP_SWIZZLE_128B))) return rc;
  if ((rc = x_encode_rows3d(&tmK, k, heads * D, tkv, batch, ldk, NS, 64, CU_TENSOR_MAP_SWIZZLE_128B))) return rc;
  if ((rc = x_encode_rows3d(&tmV, v, heads * D, tkv, batch, ldv, NS, 64, CU_TENSOR_MAP_SWIZZLE_128B))) return rc;
  if ((rc = x_encode_rows3d(&tmO, o, heads * D, tq, batch, ldo, XQ, D, CU_TENSOR_MAP_SWIZZLE_NONE))) return rc;
  XParams p;
  p.heads = heads;
  p.tq = tq;
  p.tkv = tkv;
  p.tiles_per_bh = ceil_div(tq, XQ);
  p.run = 1;
  for (int r = 4; r > 1; r >>= 1)
    if (p.tiles_per_bh % r == 0) {
      p.run = r;
      break;
    }
  p.chunks_per_bh = p.tiles_per_bh / p.run;
  p.total_runs = batch * p.chunks_per_bh * heads;
  const int sms = saspa_num_sms();
  const int ctas = p.total_runs < sms ? p.total_runs : sms;
  p.scale_log2 = scale * 1.4426950408889634f;
  xattn_tc_kernel<D, NS><<<ctas, XT_THREADS, C::SMEM, stream>>>(tmQ, tmK, tmV, tmO, p);
  SASPA_LAUNCH_CHECK();
  return SASPA_OK;
}

template <int D>
int launch_xtc_d(const void* q, int ldq, const void* k, int ldk, const void* v, int ldv, void* o, int ldo, int batch, int heads, int tq, int tkv,
                 float scale, cudaStream_t stream) {
  if (tkv <= 80) return launch_xtc<D, 80>(q, ldq, k, ldk, v, ldv, o, ldo, batch, heads, tq, tkv, scale, stream);
  return launch_xtc<D, 128>(q, ldq, k, ldk, v, ldv, o, ldo, batch, heads, tq, tkv, scale, stream);
}

}  // namespace

// Returns SASPA_ERR_UNSUPPORTED (without setting an error) when the shape has no instantiation; the caller then falls back.
int saspa_xattention_tc(const void* q, int ldq, const void* k, int ldk, const void* v, int ldv, void* o, int ldo, int batch, int heads, int tq,
                        int tkv, int d, float scale, cudaStream_t stream) {
  if (tkv > 128 || tkv < 1 || scale <= 0.0f || (ldo % 8) != 0 || (reinterpret_cast<uintptr_t>(o) & 15) != 0) return SASPA_ERR_UNSUPPORTED;
  switch (d) {
    case 40: return launch_xtc_d<40>(q, ldq, k, ldk, v, ldv, o, ldo, batch, heads, tq, tkv, scale, stream);
    case 64: return launch_xtc_d<64>(q, ldq, k, ldk, v, ldv, o, ldo, batch, heads, tq, tkv, scale, stream);
    case 80: return launch_xtc_d<80>(q, ldq, k, ldk, v, ldv, o, ldo, batch, heads, tq, tkv, scale, stream);
    case 128: return launch_xtc_d<128>(q, ldq, k, ldk, v, ldv, o, ldo, batch, heads, tq, tkv, scale, stream);
    default: return SASPA_ERR_UNSUPPORTED;
  }
}
