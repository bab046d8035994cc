// tcgen05 / TMEM / TMA GEMM and implicit-GEMM convolution for sm_100a (B200).
//
// One persistent, warp-specialised kernel serves both entry points:
//   saspa_gemm_bf16          D[M,N] = epi(A[M,K] . B[N,K]^T)           (Linear, 1x1 conv, im2col'd conv)
//   saspa_conv2d_igemm_bf16  NHWC stride-1 "same" conv, ksize 1|3, optional 2-source channel concat
// They replace the cuBLAS/cuDNN calls under diffusers' UNet2DConditionModel / ControlNetModel /
// AutoencoderKL forward (reference: pipe(**pipe_args), run_aug/run_aug.py:278) and the filter nets
// (all_utils/utils.py:361 WSDAN_CAL, :152-164 CLIP).
//
// Structure per CTA (384 threads, 1 CTA / SM):
//   warp 0   : TMA producer  - cp.async.bulk.tensor into a ring of 128B-swizzled smem tiles, signalled by mbarrier
//              complete_tx.  GEMM: 2-D boxes.  Conv, per-tap mode: nine shifted 4-D NHWC boxes per 64-channel chunk
//              (TMA's out-of-bounds zero fill IS the conv padding).  Conv, halo mode (3x3): ONE box of (16+2)x(8+2)
//              pixels per chunk; the nine taps are row-shifted UMMA descriptors into it (the 128B swizzle is a
//              function of the absolute smem address, so shifted windows read back what TMA wrote) -- operand
//              traffic out of L2 is what bounds these kernels.
//   warp 1   : allocates 512 TMEM columns; one lane issues tcgen05.mma.kind::f16 (fp32 accumulate in TMEM),
//              tcgen05.commit frees smem stages and publishes the accumulator.  Two accumulator buffers (columns
//              0 / 256) let the epilogue of tile i overlap the main loop of tile i+1.
//   warps 2-9: epilogue, two groups of four warps (one warp per TMEM lane quadrant); group g owns the
//              32-column output panels k = g, g+2, ...:  tcgen05.ld 32x32b.x32 (one output row per thread),
//              fused bias / per-image row bias (time embedding) / activation / GEGLU / alpha / residual.
//              bf16 outputs leave through shared memory: each thread writes its 64-byte row slice into a
//              64B-swizzled [128 x 32] staging panel and arrives on the panel's mbarrier; the residual arrives the
//              same way (TMA load into the staging panel, prefetched two panels ahead, across tile boundaries).
//              fp32 / unaligned outputs take the direct (row-per-thread, 16-byte) global path.  LayerNorm folded into
//              the GEMMs around it: per-row partial (sum, sum of squares) out, mean / rstd applied to the accumulator.
//   warps 10-11: store helpers, one per epilogue group: wait for a panel's four arrivals, issue its TMA store, prefetch
//              the next residual panel into a drained buffer (or signal it free) -- no epilogue warp waits for another.
//
// CTAS = 2 (cta_group::2): two CTAs of a cluster (one TPC) compute a 256 x BN tile.  Each CTA loads its own 128 rows
// of A and HALF of the B tile; the leader CTA issues tcgen05.mma.cta_group::2, which reads B from both CTAs' shared
// memory and writes each CTA's 128 accumulator rows into that CTA's TMEM.  This halves the B bytes every SM pulls
// from L2 and reads from shared memory per MMA -- the single-CTA 128 x 160 tile is shared-memory-bandwidth bound.
// TMA loads of both CTAs complete on the leader's "full" barrier; tcgen05.commit multicasts "empty" / "accumulator
// ready" to both CTAs; the peer's epilogue warps release the accumulator with a remote, RELAXED mbarrier arrive (a release
// at cluster scope costs a memory barrier per arrive and made pairs lose below K = 1024).
#include "tc_ptx.cuh"
#include <atomic>

#include "../../include/saspa_b200.h"
#include "tuning_hooks.h"

namespace {

constexpr int BM = 128;
constexpr int BK = 64;  // 64 bf16 = 128 B = one swizzle atom row
constexpr int EPI_WARPS = 8;
constexpr int HELPER_WARPS = 2;  // one per epilogue group: issues the group's TMA stores and residual prefetches
constexpr int NUM_THREADS = 64 + 32 * EPI_WARPS + 32 * HELPER_WARPS;
constexpr int A_BYTES = BM * BK * 2;
constexpr int TMEM_COLS = 512;
constexpr int ACC_STRIDE = 256;  // column offset between the two accumulator buffers
constexpr int PANEL = 32;        // output columns per staging panel (64 B of bf16 per row)
constexpr int PANEL_BYTES = BM * PANEL * 2;
constexpr int SMEM_LIMIT = 227 * 1024;
// halo mode: one TMA box of (HALO_BH + 2) x (HALO_BW + 2) pixels x 64 channels feeds all nine 3x3 taps
constexpr int HALO_BW = 8, HALO_BH = 16;
constexpr int HALO_ROWS = (HALO_BW + 2) * (HALO_BH + 2);
constexpr int HALO_TX_BYTES = HALO_ROWS * 128;
constexpr int HALO_STAGE_BYTES = (HALO_TX_BYTES + 1023) / 1024 * 1024;
constexpr int HALO_STAGES = 2;

struct GemmParams {
  int M, N, K;
  int num_m_tiles, num_n_tiles;
  int mode;  // 0 = GEMM, 1 = implicit conv (one TMA box per tap), 2 = implicit 3x3 conv from a halo tile
  // conv geometry
  int n_img, H, W, c0, c1, ksize;  // H, W: OUTPUT map
  int stride, pad;                 // per-tap mode: input pixel = output pixel * stride + tap - pad (stride 2: TMA element strides)
  int bw, bh, bn;
  int tiles_x, tiles_y;
  // epilogue
  const float* bias;
  const float* row_bias;
  int rows_per_group;
  int ld_row_bias;
  int act;
  int act_post;
  float alpha;
  const __nv_bfloat16* residual;
  int ld_res;
  float beta;
  int out_fp32;
  void* D;
  int ldd;
  int n_out;      // valid output columns (N, or N/2 for GEGLU)
  int vec_ok;     // 16-byte vector path allowed for residual loads / stores (direct path)
  int tma_store;  // bf16 output (and residual) move through smem staging panels + TMA
  // LayerNorm folded into this GEMM (mode 0): acc = x . W'^T with W' = W * gamma; v = rstd_r * (acc - mean_r * colsum_j) + bias'_j
  const float2* ln_stats;  // [M, ln_slots] partial (sum, sum of squares) of each A row, written by the producer's epilogue
  int ln_slots;
  const float* ln_colsum;  // [N] sum_k W'[j, k] (fp32 sum of the bf16 weights the MMA multiplies)
  float ln_eps, ln_inv_k;
  // per-row partial statistics of THIS GEMM's bf16 output (the next LayerNorm's input): slot = 2 * n_tile + epilogue group
  float2* stats_out;
  int stats_slots;
  int m_rev;  // GEMM mode: row tiles in descending order
};

using namespace tcx;

template <int BN, int CTAS>
struct Cfg {
  static constexpr int B_BYTES = BN * BK * 2 / CTAS;  // per CTA: its share of the B tile
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  // staging panels per epilogue group: 4 (residual prefetch distance 2) unless the operand ring needs the room
  static constexpr int NBUF = (BN > 160 && CTAS == 1) ? 2 : 4;
  static constexpr int PD = NBUF / 2;
  static constexpr int STAGING_BYTES = 2 * NBUF * PANEL_BYTES;
  static constexpr int RING_BUDGET = SMEM_LIMIT - STAGING_BYTES - 1024 /*align slack*/ - 512 /*barriers*/ - 4096 /*bias + LayerNorm column sums*/;
  static constexpr int RING_BYTES = RING_BUDGET / 1024 * 1024;
  static constexpr int STAGES = (RING_BYTES / STAGE_BYTES) > 8 ? 8 : (RING_BYTES / STAGE_BYTES);
  // halo mode (3x3 conv): 2 halo stages of the activation + a ring of weight k-blocks
  static constexpr int HB_STAGES = ((RING_BYTES - HALO_STAGES * HALO_STAGE_BYTES) / B_BYTES) > 8 ? 8 : ((RING_BYTES - HALO_STAGES * HALO_STAGE_BYTES) / B_BYTES);
  static constexpr int SMEM_BYTES = RING_BYTES + STAGING_BYTES + 1024 + 512 + 4096;
  static_assert(STAGES >= 3 && HB_STAGES >= 4, "operand ring too shallow");
  static_assert(BN % PANEL == 0, "BN must be a multiple of the staging panel width");
  static_assert((BN / CTAS) % 8 == 0, "each CTA's B share must be whole 8-row core matrices");
};

__device__ __forceinline__ float apply_act(float v, int act) {
  switch (act) {
    case SASPA_ACT_SILU: return silu_f(v);
    case SASPA_ACT_GELU: return gelu_erf_f(v);
    case SASPA_ACT_RELU: return fmaxf(v, 0.0f);
    case SASPA_ACT_QUICKGELU: return quick_gelu_f(v);
    default: return v;
  }
}

// (sum, sum of squares) of the 32 fp32 values of a panel row, taken BEFORE the bf16 rounding of the store (two instructions per
// element instead of four; the rounding noise averages out over a 320..1280-wide row: |mean error| ~ 2^-9 |x| / sqrt(C)), in four
// independent chains
__device__ __forceinline__ void row_stats_add(const float (&f)[32], float& s, float& q) {
  float s4[4] = {0.f, 0.f, 0.f, 0.f}, q4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int j = 0; j < 32; ++j) {
    s4[j & 3] += f[j];
    q4[j & 3] = fmaf(f[j], f[j], q4[j & 3]);
  }
  s += (s4[0] + s4[1]) + (s4[2] + s4[3]);
  q += (q4[0] + q4[1]) + (q4[2] + q4[3]);
}

__device__ __forceinline__ void tma_store_2d(const CUtensorMap* tm, const void* src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(reinterpret_cast<uint64_t>(tm)),
               "r"(smem_u32(src)), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void tma_store_4d(const CUtensorMap* tm, const void* src, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];" ::"l"(reinterpret_cast<uint64_t>(tm)),
               "r"(smem_u32(src)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_read() {
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}

// ---- operand loads / MMA / commit, one- and two-CTA flavours.  `bar` is a shared::cluster address: the CTA's own
// barrier for CTAS == 1, the leader CTA's barrier (mapa rank 0) for CTAS == 2. ----
template <int CTAS>
__device__ __forceinline__ void op_load_2d(const CUtensorMap* tm, void* dst, uint32_t bar, int c0, int c1) {
  if constexpr (CTAS == 1) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(smem_u32(dst)),
                 "l"(reinterpret_cast<uint64_t>(tm)), "r"(bar), "r"(c0), "r"(c1)
                 : "memory");
  } else {
    asm volatile("cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
                     smem_u32(dst)),
                 "l"(reinterpret_cast<uint64_t>(tm)), "r"(bar), "r"(c0), "r"(c1)
                 : "memory");
  }
}
template <int CTAS>
__device__ __forceinline__ void op_load_4d(const CUtensorMap* tm, void* dst, uint32_t bar, int c0, int c1, int c2, int c3) {
  if constexpr (CTAS == 1) {
    asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(
                     smem_u32(dst)),
                 "l"(reinterpret_cast<uint64_t>(tm)), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
                 : "memory");
  } else {
    asm volatile(
        "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(
            smem_u32(dst)),
        "l"(reinterpret_cast<uint64_t>(tm)), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
        : "memory");
  }
}
template <int CTAS>
__device__ __forceinline__ void op_mma(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  if constexpr (CTAS == 1) {
    tc_mma_bf16(tmem_d, desc_a, desc_b, idesc, accumulate);
  } else {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
        "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
  }
}
// arrives (once the MMAs issued so far have completed) on `bar` of this CTA and, for CTAS == 2, of the peer CTA too
template <int CTAS>
__device__ __forceinline__ void op_commit(uint64_t* bar) {
  if constexpr (CTAS == 1) {
    tc_commit(bar);
  } else {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(bar)),
                 "h"((uint16_t)3)
                 : "memory");
  }
}
__device__ __forceinline__ uint32_t mapa_rank(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
// Accumulator hand-back of a CTA pair ("this warp has read its TMEM rows"): ordered against the tcgen05.ld's by the
// tcgen05.fence::before_thread_sync that precedes it; no generic-memory write has to be published to the MMA issuer, so the arrive is
// RELAXED.  With .release.cluster the compiler put a cluster-scope memory barrier in front of every arrive (MEMBAR stalls: 18 % of all
// warp samples of a K = 320 pair launch, profiles/r2_pair_handback_membar.txt) -- the reason pairs lost to single-CTA tiles below K = 1024.
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}

template <int BN, int CTAS>
__global__ void __launch_bounds__(NUM_THREADS, 1)
    gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA0, const __grid_constant__ CUtensorMap tmA1,
                   const __grid_constant__ CUtensorMap tmB, const __grid_constant__ CUtensorMap tmD,
                   const __grid_constant__ CUtensorMap tmR, const GemmParams p) {
  using C = Cfg<BN, CTAS>;
  constexpr int STAGES = C::STAGES;
  constexpr int NBUF = C::NBUF;
  constexpr int PD = C::PD;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
  uint8_t* sA = smem;
  uint8_t* sB = smem + STAGES * A_BYTES;
  uint8_t* sHalo = smem;                                    // mode 2: HALO_STAGES halo tiles ...
  uint8_t* sBh = smem + HALO_STAGES * HALO_STAGE_BYTES;     // ... then HB_STAGES weight k-blocks
  uint8_t* sStage = smem + C::RING_BYTES;                   // 1024-aligned
  uint64_t* bars = reinterpret_cast<uint64_t*>(sStage + C::STAGING_BYTES);
  uint64_t* full = bars;         // [8]   (CTAS == 2: only the leader's are waited on)
  uint64_t* empty = bars + 8;    // [8]
  uint64_t* afull = bars + 16;   // [HALO_STAGES]
  uint64_t* aempty = bars + 18;  // [HALO_STAGES]
  uint64_t* tfull = bars + 20;
  uint64_t* tempty = bars + 22;  // (CTAS == 2: only the leader's are waited on)
  uint64_t* rfull = bars + 24;   // [2 groups][NBUF]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(rfull + 2 * NBUF);
  uint64_t* pfull = bars + 34;   // [2 groups][4] staging panel written by the group's four epilogue warps (-> store helper)
  uint64_t* pfree = bars + 42;   // [2 groups][4] the store that read a staging panel has drained (store helper -> epilogue warps)
  float* sBias = reinterpret_cast<float*>(bars) + 128;  // [2][256]: the tile's bias slice, staged once per tile (512 B past the barriers)
  float* sColsum = sBias + 512;                         // [2][256]: the tile's slice of the LayerNorm column sums

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint32_t rank = 0;
  if constexpr (CTAS == 2) asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
  const bool is_leader = rank == 0;
  // persistent schedule over (group of CTAS m-tiles, n-tile); both CTAs of a pair walk the same sequence
  const int num_clusters = gridDim.x / CTAS, cid = blockIdx.x / CTAS;
  const int total_tiles = ((p.num_m_tiles + CTAS - 1) / CTAS) * p.num_n_tiles;
  const int cchunks = (p.mode != 0) ? (p.c0 + p.c1 + BK - 1) / BK : 0;
  const int num_kb = (p.mode != 0) ? p.ksize * p.ksize * cchunks : (p.K + BK - 1) / BK;
  // m_rev walks the row tiles from the last to the first: a GEMM that reads what the previous launch wrote front to back then starts on
  // the rows that are still in L2
  const int m_groups = (p.num_m_tiles + CTAS - 1) / CTAS;
  auto tile_m = [&](int tile) {
    const int g = tile / p.num_n_tiles;
    return (p.m_rev ? m_groups - 1 - g : g) * CTAS + (int)rank;
  };

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA0);
    tma_prefetch_desc(&tmB);
    if (p.mode != 0 && p.c1 > 0) tma_prefetch_desc(&tmA1);
    if (p.tma_store) {
      tma_prefetch_desc(&tmD);
      if (p.residual) tma_prefetch_desc(&tmR);
    }
    for (int s = 0; s < 8; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    for (int a = 0; a < HALO_STAGES; ++a) {
      mbar_init(&afull[a], 1);
      mbar_init(&aempty[a], 1);
    }
    for (int i = 0; i < 8; ++i) {
      mbar_init(&pfull[i], 4);
      mbar_init(&pfree[i], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tfull[a], 1);
      mbar_init(&tempty[a], EPI_WARPS * CTAS);
    }
    for (int i = 0; i < 2 * NBUF; ++i) mbar_init(&rfull[i], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    if constexpr (CTAS == 1) {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(TMEM_COLS) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    } else {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(TMEM_COLS) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
  }
  tc_fence_before();
  __syncthreads();
  if constexpr (CTAS == 2) cluster_sync_all();  // the peer's barriers are initialised before anything signals them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // ---- staging-panel bookkeeping shared by the epilogue warps and their store helpers: group g owns panels k = g, g + 2, ... of a tile ----
  const bool geglu = (p.act == SASPA_ACT_GEGLU);
  const int out_bn = geglu ? BN / 2 : BN;
  const int NP = out_bn / PANEL;
  auto panel_valid = [&](int tile, int k) { return (tile % p.num_n_tiles) * out_bn + k * PANEL < p.n_out; };
  auto panel_next = [&](int& tile, int& k, int g) {  // next valid (tile, panel) of group g; tile >= total_tiles at the end
    for (;;) {
      k += 2;
      if (k >= NP) {
        tile += num_clusters;
        k = g;
        if (tile >= total_tiles) return;
      }
      if (k < NP && panel_valid(tile, k)) return;
    }
  };
  auto panel_coords = [&](int tile, int k, int& col, int& c1, int& c2, int& c3) {
    const int m_blk = tile_m(tile), n_blk = tile % p.num_n_tiles;
    col = n_blk * out_bn + k * PANEL;
    if (p.mode == 0) {
      c1 = m_blk * BM;
      c2 = c3 = 0;
    } else {
      int tx = m_blk % p.tiles_x, rr = m_blk / p.tiles_x;
      int ty = rr % p.tiles_y, tn = rr / p.tiles_y;
      c1 = tx * p.bw;
      c2 = ty * p.bh;
      c3 = tn * p.bn;
    }
  };
  auto issue_res_load = [&](int tile, int k, uint8_t* dst, uint64_t* bar) {
    int col, c1, c2, c3;
    panel_coords(tile, k, col, c1, c2, c3);
    mbar_expect_tx(bar, PANEL_BYTES);
    if (p.mode == 0)
      tma_load_2d(&tmR, dst, bar, col, c1);
    else
      tma_load_4d(&tmR, dst, bar, col, c1, c2, c3);
  };

  if (warp == 0) {
    // ===================== TMA producer (each CTA fills its own smem; completion lands on the leader's barrier) =====================
    // (converged warp, one elected lane per instruction -- see the MMA issuer)
    const int b_rows = BN / CTAS;  // this CTA's share of the B tile
    if (p.mode == 2) {
      int sa = 0, sb = 0;
      uint32_t pa = 0, pb = 0;
      const int ctot = p.c0 + p.c1;
      for (int tile = cid; tile < total_tiles; tile += num_clusters) {
        const int m_blk = tile_m(tile), n_blk = tile % p.num_n_tiles;
        const int tx = m_blk % p.tiles_x, r = m_blk / p.tiles_x;
        const int ty = r % p.tiles_y, tn = r / p.tiles_y;
        for (int cc = 0; cc < cchunks; ++cc) {
          const int c = cc * BK;
          mbar_wait(&aempty[sa], pa ^ 1);
          if (is_leader && elect_one()) mbar_expect_tx(&afull[sa], HALO_TX_BYTES * CTAS);
          const uint32_t abar = (CTAS == 2) ? mapa_rank(smem_u32(&afull[sa]), 0) : smem_u32(&afull[sa]);
          if (c < p.c0) {
            if (elect_one()) op_load_4d<CTAS>(&tmA0, sHalo + sa * HALO_STAGE_BYTES, abar, c, tx * HALO_BW - 1, ty * HALO_BH - 1, tn);
          } else {
            if (elect_one()) op_load_4d<CTAS>(&tmA1, sHalo + sa * HALO_STAGE_BYTES, abar, c - p.c0, tx * HALO_BW - 1, ty * HALO_BH - 1, tn);
          }
          if (++sa == HALO_STAGES) {
            sa = 0;
            pa ^= 1;
          }
          for (int tap = 0; tap < 9; ++tap) {
            mbar_wait(&empty[sb], pb ^ 1);
            if (is_leader && elect_one()) mbar_expect_tx(&full[sb], C::B_BYTES * CTAS);
            const uint32_t bbar = (CTAS == 2) ? mapa_rank(smem_u32(&full[sb]), 0) : smem_u32(&full[sb]);
            if (elect_one()) op_load_2d<CTAS>(&tmB, sBh + sb * C::B_BYTES, bbar, tap * ctot + c, n_blk * BN + (int)rank * b_rows);
            if (++sb == C::HB_STAGES) {
              sb = 0;
              pb ^= 1;
            }
          }
        }
      }
    } else {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = cid; tile < total_tiles; tile += num_clusters) {
        const int m_blk = tile_m(tile), n_blk = tile % p.num_n_tiles;
        int x0 = 0, y0 = 0, n0 = 0;
        if (p.mode == 1) {
          int tx = m_blk % p.tiles_x, r = m_blk / p.tiles_x;
          int ty = r % p.tiles_y, tn = r / p.tiles_y;
          x0 = tx * p.bw;
          y0 = ty * p.bh;
          n0 = tn * p.bn;
        }
        const int pad = p.pad;
        x0 *= p.stride;
        y0 *= p.stride;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&empty[stage], phase ^ 1);
          if (is_leader && elect_one()) mbar_expect_tx(&full[stage], C::STAGE_BYTES * CTAS);
          const uint32_t fbar = (CTAS == 2) ? mapa_rank(smem_u32(&full[stage]), 0) : smem_u32(&full[stage]);
          uint8_t* a_dst = sA + stage * A_BYTES;
          uint8_t* b_dst = sB + stage * C::B_BYTES;
          int kB;
          if (p.mode == 0) {
            kB = kb * BK;
            if (elect_one()) op_load_2d<CTAS>(&tmA0, a_dst, fbar, kB, m_blk * BM);
          } else {
            int tap = kb / cchunks, cc = kb - tap * cchunks;
            int ky = tap / p.ksize, kx = tap - ky * p.ksize;
            int c = cc * BK;
            kB = tap * (p.c0 + p.c1) + c;
            if (c < p.c0) {
              if (elect_one()) op_load_4d<CTAS>(&tmA0, a_dst, fbar, c, x0 + kx - pad, y0 + ky - pad, n0);
            } else {
              if (elect_one()) op_load_4d<CTAS>(&tmA1, a_dst, fbar, c - p.c0, x0 + kx - pad, y0 + ky - pad, n0);
            }
          }
          if (elect_one()) op_load_2d<CTAS>(&tmB, b_dst, fbar, kB, n_blk * BN + (int)rank * b_rows);
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (leader CTA only) =====================
    // The whole warp runs the loops converged and ONE ELECTED lane issues each tcgen05 instruction (elect.sync): under a divergent
    // `if (lane == 0)` ptxas serialises every UTCHMMA through an ELECT / R2UR / branch loop (~90 clk per MMA -- more than the
    // 80 clk a 128 x 160 x 16 MMA takes, i.e. the narrower tiles were issue-bound).
    constexpr uint32_t idesc = make_idesc(BM * CTAS, BN);
    if (is_leader && p.mode == 2) {
      int sa = 0, sb = 0;
      uint32_t pa = 0, pb = 0;
      int it = 0;
      for (int tile = cid; tile < total_tiles; tile += num_clusters, ++it) {
        const int acc = it & 1;
        const uint32_t acc_phase = (it >> 1) & 1;
        mbar_wait(&tempty[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * ACC_STRIDE;
        for (int cc = 0; cc < cchunks; ++cc) {
          mbar_wait(&afull[sa], pa);
          tc_fence_after();
          const uint32_t halo = smem_u32(sHalo + sa * HALO_STAGE_BYTES);
#pragma unroll 1
          for (int tap = 0; tap < 9; ++tap) {
            mbar_wait(&full[sb], pb);
            tc_fence_after();
            const int ky = tap / 3, kx = tap - 3 * ky;
            // output pixel (ly, lx) reads halo pixel (ly + ky, lx + kx): a row shift inside the halo tile.  Each
            // 8-row core-matrix group is one image row of the tile (8 px x 128 B contiguous); groups are one halo
            // row ((HALO_BW + 2) x 128 B) apart.
            const uint64_t a_desc = make_smem_desc_sbo(halo + (ky * (HALO_BW + 2) + kx) * 128, (HALO_BW + 2) * 128);
            const uint64_t b_desc = make_smem_desc(smem_u32(sBh + sb * C::B_BYTES));
#pragma unroll
            for (int k = 0; k < BK / 16; ++k)
              if (elect_one()) op_mma<CTAS>(d_tmem, a_desc + 2 * k, b_desc + 2 * k, idesc, (cc | tap | k) != 0 ? 1u : 0u);
            if (elect_one()) op_commit<CTAS>(&empty[sb]);
            if (++sb == C::HB_STAGES) {
              sb = 0;
              pb ^= 1;
            }
          }
          if (elect_one()) op_commit<CTAS>(&aempty[sa]);
          if (++sa == HALO_STAGES) {
            sa = 0;
            pa ^= 1;
          }
        }
        if (elect_one()) op_commit<CTAS>(&tfull[acc]);
      }
    } else if (is_leader) {
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int tile = cid; tile < total_tiles; tile += num_clusters, ++it) {
        const int acc = it & 1;
        const uint32_t acc_phase = (it >> 1) & 1;
        mbar_wait(&tempty[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * ACC_STRIDE;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&full[stage], phase);
          tc_fence_after();
          const uint64_t a_desc = make_smem_desc(smem_u32(sA + stage * A_BYTES));
          const uint64_t b_desc = make_smem_desc(smem_u32(sB + stage * C::B_BYTES));
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) {
            // +32 B along K inside the 128B swizzle atom = +2 in the 16-byte start-address field
            if (elect_one()) op_mma<CTAS>(d_tmem, a_desc + 2 * k, b_desc + 2 * k, idesc, (kb | k) != 0 ? 1u : 0u);
          }
          if (elect_one()) op_commit<CTAS>(&empty[stage]);
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
        if (elect_one()) op_commit<CTAS>(&tfull[acc]);
      }
    }
  } else if (warp >= 2 + EPI_WARPS) {
    // ===================== store helpers (warps 10, 11): one per epilogue group =====================
    // All of a group's TMA traffic is issued here -- the bf16 output panels and the residual prefetch -- so that no epilogue warp ever
    // waits for another one: an epilogue warp fills its 32 rows of a staging panel and arrives on the panel's `pfull` barrier (count 4);
    // this warp waits for the four arrivals, issues the store and recycles the buffer, either by loading the next residual panel into
    // it (its `rfull` barrier then tells the epilogue warps the buffer is theirs again) or by arriving on `pfree` once the store has
    // drained.  (With the group leader issuing, the other three warps of a group sat at a barrier for the leader's bookkeeping on every
    // panel: 28 % of the epilogue warps' time on the K = 320 GEMMs, profiles/r2_smallk_gemm_epilogue.txt.)
    const int grp = warp - (2 + EPI_WARPS);
    if (p.tma_store) {
      const bool tma_res = p.residual != nullptr;
      uint8_t* stg = sStage + grp * NBUF * PANEL_BYTES;
      uint64_t* rf = rfull + grp * NBUF;
      uint64_t* pf = pfull + grp * 4;
      uint64_t* fr = pfree + grp * 4;
      int pf_tile = cid, pf_k = grp - 2, pf_n = 0;  // residual prefetch cursor
      if (tma_res) {
        panel_next(pf_tile, pf_k, grp);
        for (int i = 0; i < PD && pf_tile < total_tiles; ++i) {
          if (lane == 0) issue_res_load(pf_tile, pf_k, stg + (pf_n % NBUF) * PANEL_BYTES, &rf[pf_n % NBUF]);
          ++pf_n;
          panel_next(pf_tile, pf_k, grp);
        }
      }
      int s_tile = cid, s_k = grp - 2;  // store cursor
      panel_next(s_tile, s_k, grp);
      int n = 0, freed = 0;
      constexpr int KEEP = NBUF >= 2 ? NBUF - 2 : 0;  // stores that may still be reading their panel after the wait below
      while (s_tile < total_tiles) {
        const int buf = n % NBUF;
        if (tma_res && pf_tile < total_tiles) {
          // panel n + PD reuses the buffer of store n + PD - NBUF, which must have drained
          if (lane == 0) {
            bulk_wait_read<NBUF - PD - 1>();
            issue_res_load(pf_tile, pf_k, stg + (pf_n % NBUF) * PANEL_BYTES, &rf[pf_n % NBUF]);
          }
          ++pf_n;
          panel_next(pf_tile, pf_k, grp);
        }
        mbar_wait(&pf[buf], (uint32_t)(n / NBUF) & 1u);
        if (lane == 0) {
          int col, c1, c2, c3;
          panel_coords(s_tile, s_k, col, c1, c2, c3);
          if (p.mode == 0)
            tma_store_2d(&tmD, stg + buf * PANEL_BYTES, col, c1);
          else
            tma_store_4d(&tmD, stg + buf * PANEL_BYTES, col, c1, c2, c3);
          bulk_commit();
          if (!tma_res) {
            bulk_wait_read<KEEP>();  // stores <= n - KEEP have drained
            for (; freed <= n - KEEP; ++freed) mbar_arrive(&fr[freed % NBUF]);
          }
        }
        ++n;
        panel_next(s_tile, s_k, grp);
        __syncwarp();
      }
      if (lane == 0) bulk_wait_read<0>();  // staging panels must outlive the stores that read them
    }
  } else {
    // ===================== epilogue (warps 2..9) =====================
    const int grp = (warp - 2) >> 2;  // epilogue group: owns panels k = grp, grp + 2, ...
    const int q = warp & 3;           // TMEM lane quadrant this warp may access
    const int r = q * 32 + lane;      // accumulator row of this thread
    uint8_t* stg = sStage + grp * NBUF * PANEL_BYTES;
    uint64_t* rf = rfull + grp * NBUF;
    uint64_t* pf = pfull + grp * 4;
    uint64_t* fr = pfree + grp * 4;
    const uint32_t row_off = (uint32_t)r * 64u;
    const uint32_t row_swz = (uint32_t)(r >> 1) & 3u;  // SWIZZLE_64B: 16-byte chunk index ^= address bits [7,8]
    const bool use_tma = p.tma_store != 0;
    const bool tma_res = use_tma && p.residual != nullptr;

    int n_seq = 0;  // panels processed by this group so far
    int it = 0;
    int bias_n_blk = -1, bias_buf = 0;
    // Folded LayerNorm with four statistics slots per row (a 320-wide producer): the row's partial sums are fetched ONE TILE AHEAD.  At
    // that width the epilogue is the critical path (K = 320 gives the tensor core 1600 clk of work per tile), the accumulator is ready
    // when a tile starts and the L2 latency of the loads used to be exposed on every tile.
    const bool ln_ahead = p.ln_stats != nullptr && p.ln_slots == 4 && p.mode == 0 && (reinterpret_cast<uintptr_t>(p.ln_stats) & 15) == 0;
    float4 pre0 = make_float4(0.f, 0.f, 0.f, 0.f), pre1 = pre0;
    auto ln_fetch = [&](int tile) {
      const long long row = (long long)tile_m(tile) * BM + r;
      if (row < p.M) {
        const float4* sp = reinterpret_cast<const float4*>(p.ln_stats + (size_t)row * 4);
        pre0 = __ldg(sp);
        pre1 = __ldg(sp + 1);
      }
    };
    if (ln_ahead && cid < total_tiles) ln_fetch(cid);
    for (int tile = cid; tile < total_tiles; tile += num_clusters, ++it) {
      const int acc = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1;
      const int m_blk = tile_m(tile), n_blk = tile % p.num_n_tiles;
      long long pix;
      bool row_ok;
      int group;
      if (p.mode == 0) {
        pix = (long long)m_blk * BM + r;
        row_ok = pix < p.M;
        group = p.row_bias ? (int)(pix / p.rows_per_group) : 0;
      } else {
        int tx = m_blk % p.tiles_x, rr = m_blk / p.tiles_x;
        int ty = rr % p.tiles_y, tn = rr / p.tiles_y;
        int lx = r % p.bw, t2 = r / p.bw;
        int ly = t2 % p.bh, ln = t2 / p.bh;
        int x = tx * p.bw + lx, y = ty * p.bh + ly, n = tn * p.bn + ln;
        row_ok = (x < p.W) && (y < p.H) && (n < p.n_img);
        pix = ((long long)n * p.H + y) * p.W + x;
        group = row_ok ? n : 0;
      }
      const int col_base_in = n_blk * BN;       // column in the (interleaved) weight / bias space
      const int col_base_out = n_blk * out_bn;  // column in the output
      // the bias slice of this tile goes through shared memory: its global-load latency is paid once per tile, off the
      // accumulator-drain path (it used to stall every panel).  Two buffers: nobody runs more than a tile ahead.
      // It is reloaded only when the CTA moves to another N tile (with 148 CTAs and an even number of N tiles: never after the
      // first tile), so the barrier below is off the steady-state path.  Safe with two buffers: a warp that is still on the previous
      // tile reads the other buffer, and nobody is more than one reload behind (the barrier).
      if ((p.bias || p.ln_colsum) && n_blk != bias_n_blk) {
        bias_n_blk = n_blk;
        bias_buf ^= 1;
        const int t = (int)threadIdx.x - 64;
        if (t < BN) {
          if (p.bias) sBias[bias_buf * 256 + t] = (col_base_in + t < p.N) ? __ldg(p.bias + col_base_in + t) : 0.0f;
          if (p.ln_colsum) sColsum[bias_buf * 256 + t] = (col_base_in + t < p.N) ? __ldg(p.ln_colsum + col_base_in + t) : 0.0f;
        }
        asm volatile("bar.sync 3, 256;" ::: "memory");
      }
      const float* sb = sBias + bias_buf * 256;
      const float* sc = sColsum + bias_buf * 256;
      // folded LayerNorm: this row's mean / rstd from the producer's partials, summed in slot order (fixed => batch invariant);
      // the loads are in flight while the accumulator is still being produced
      float ln_mean = 0.0f, ln_rstd = 1.0f;
      if (p.ln_stats && row_ok) {
        float a = 0.0f, b = 0.0f;
        if (ln_ahead) {  // fetched while the previous tile drained; same summation order as the loop below
          a = ((pre0.x + pre0.z) + pre1.x) + pre1.z;
          b = ((pre0.y + pre0.w) + pre1.y) + pre1.w;
        } else {
          const float2* sp = p.ln_stats + (size_t)pix * p.ln_slots;
          for (int s2 = 0; s2 < p.ln_slots; ++s2) {
            const float2 v2 = __ldg(sp + s2);
            a += v2.x;
            b += v2.y;
          }
        }
        ln_mean = a * p.ln_inv_k;
        ln_rstd = rsqrtf(fmaxf(b * p.ln_inv_k - ln_mean * ln_mean, 0.0f) + p.ln_eps);
      }
      if (ln_ahead && tile + num_clusters < total_tiles) ln_fetch(tile + num_clusters);
      float st_sum = 0.0f, st_sq = 0.0f;  // row statistics of this tile's bf16 output (p.stats_out)
      mbar_wait(&tfull[acc], acc_phase);
      tc_fence_after();
      const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16) + acc * ACC_STRIDE;
#pragma unroll 1
      for (int k = grp; k < NP; k += 2) {
        const int c0 = k * PANEL;
        if (col_base_out + c0 >= p.n_out) break;  // group-uniform
        __syncwarp();  // tcgen05.ld is .sync.aligned
        uint32_t v[32];
        tc_ld32(t_row + c0, v);
        uint32_t gt[32];
        if (geglu) tc_ld32(t_row + c0 + BN / 2, gt);
        tc_wait_ld();
        float f[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(v[j]);
        if (p.ln_stats) {
          // v = rstd * (acc - mean * colsum) + bias = fma(rstd, acc, fma(-rstd * mean, colsum, bias)): two FFMA per element, the
          // second one takes the place of the plain bias add
          const float nm = -ln_rstd * ln_mean;
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            const float4 cs = *reinterpret_cast<const float4*>(sc + c0 + j);
            float4 b = make_float4(0.f, 0.f, 0.f, 0.f);
            if (p.bias) b = *reinterpret_cast<const float4*>(sb + c0 + j);
            f[j] = fmaf(ln_rstd, f[j], fmaf(nm, cs.x, b.x)); f[j + 1] = fmaf(ln_rstd, f[j + 1], fmaf(nm, cs.y, b.y));
            f[j + 2] = fmaf(ln_rstd, f[j + 2], fmaf(nm, cs.z, b.z)); f[j + 3] = fmaf(ln_rstd, f[j + 3], fmaf(nm, cs.w, b.w));
          }
          if (geglu) {  // gate columns: statistics applied here, their bias is added with the GELU below
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
              const float4 cs = *reinterpret_cast<const float4*>(sc + BN / 2 + c0 + j);
              gt[j] = __float_as_uint(fmaf(ln_rstd, __uint_as_float(gt[j]), nm * cs.x));
              gt[j + 1] = __float_as_uint(fmaf(ln_rstd, __uint_as_float(gt[j + 1]), nm * cs.y));
              gt[j + 2] = __float_as_uint(fmaf(ln_rstd, __uint_as_float(gt[j + 2]), nm * cs.z));
              gt[j + 3] = __float_as_uint(fmaf(ln_rstd, __uint_as_float(gt[j + 3]), nm * cs.w));
            }
          }
        } else if (p.bias) {
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            const float4 b = *reinterpret_cast<const float4*>(sb + c0 + j);
            f[j] += b.x; f[j + 1] += b.y; f[j + 2] += b.z; f[j + 3] += b.w;
          }
        }
        if (geglu) {
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            float4 b = make_float4(0.f, 0.f, 0.f, 0.f);
            if (p.bias) b = *reinterpret_cast<const float4*>(sb + BN / 2 + c0 + j);
            f[j] *= gelu_erf_poly_f(__uint_as_float(gt[j]) + b.x);
            f[j + 1] *= gelu_erf_poly_f(__uint_as_float(gt[j + 1]) + b.y);
            f[j + 2] *= gelu_erf_poly_f(__uint_as_float(gt[j + 2]) + b.z);
            f[j + 3] *= gelu_erf_poly_f(__uint_as_float(gt[j + 3]) + b.w);
          }
        } else {
          if (p.row_bias && row_ok) {
            const float* rb = p.row_bias + (size_t)group * p.ld_row_bias;
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
              int n = col_base_in + c0 + j;
              if (n + 3 < p.N) {
                float4 b = __ldg(reinterpret_cast<const float4*>(rb + n));
                f[j] += b.x; f[j + 1] += b.y; f[j + 2] += b.z; f[j + 3] += b.w;
              } else {
                for (int e = 0; e < 4; ++e)
                  if (n + e < p.N) f[j + e] += __ldg(rb + n + e);
              }
            }
          }
          if (p.act != SASPA_ACT_NONE && !p.act_post) {
#pragma unroll
            for (int j = 0; j < 32; ++j) f[j] = apply_act(f[j], p.act);
          }
        }
        if (p.alpha != 1.0f) {
#pragma unroll
          for (int j = 0; j < 32; ++j) f[j] *= p.alpha;
        }
        const int n0 = col_base_out + c0;

        if (use_tma) {
          // ---------- staged path: residual in / bf16 out through a swizzled smem panel + TMA ----------
          const int buf = n_seq % NBUF;
          uint8_t* pbuf = stg + buf * PANEL_BYTES + row_off;
          if (tma_res) {
            mbar_wait(&rf[buf], (uint32_t)(n_seq / NBUF) & 1u);  // the residual panel has landed (=> the buffer's last store drained)
#pragma unroll
            for (int c = 0; c < 4; ++c) {
              const uint4 u = *reinterpret_cast<const uint4*>(pbuf + ((c ^ row_swz) << 4));
              const int j = c * 8;
              f[j + 0] += p.beta * bf16_lo(u.x); f[j + 1] += p.beta * bf16_hi(u.x);
              f[j + 2] += p.beta * bf16_lo(u.y); f[j + 3] += p.beta * bf16_hi(u.y);
              f[j + 4] += p.beta * bf16_lo(u.z); f[j + 5] += p.beta * bf16_hi(u.z);
              f[j + 6] += p.beta * bf16_lo(u.w); f[j + 7] += p.beta * bf16_hi(u.w);
            }
          } else if (n_seq >= NBUF) {
            mbar_wait(&fr[buf], (uint32_t)(n_seq / NBUF - 1) & 1u);  // the store that last read this buffer has drained
          }
          if (p.act_post) {
#pragma unroll
            for (int j = 0; j < 32; ++j) f[j] = apply_act(f[j], p.act);
          }
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            const int j = c * 8;
            uint4 u;
            u.x = pack_bf16(f[j], f[j + 1]);
            u.y = pack_bf16(f[j + 2], f[j + 3]);
            u.z = pack_bf16(f[j + 4], f[j + 5]);
            u.w = pack_bf16(f[j + 6], f[j + 7]);
            *reinterpret_cast<uint4*>(pbuf + ((c ^ row_swz) << 4)) = u;
          }
          if (p.stats_out) row_stats_add(f, st_sum, st_sq);
          fence_proxy_async_smem();  // generic-proxy writes -> visible to the TMA store the helper warp issues
          __syncwarp();
          if (lane == 0) mbar_arrive(&pf[buf]);
          ++n_seq;
          continue;
        }

        // ---------- direct path (fp32 / unaligned outputs) ----------
        if (row_ok) {
          if (p.residual) {
            const __nv_bfloat16* rp = p.residual + (size_t)pix * p.ld_res + n0;
            if (p.vec_ok && n0 + 31 < p.n_out) {
#pragma unroll
              for (int j = 0; j < 32; j += 8) {
                uint4 u = __ldg(reinterpret_cast<const uint4*>(rp + j));
                f[j + 0] += p.beta * bf16_lo(u.x); f[j + 1] += p.beta * bf16_hi(u.x);
                f[j + 2] += p.beta * bf16_lo(u.y); f[j + 3] += p.beta * bf16_hi(u.y);
                f[j + 4] += p.beta * bf16_lo(u.z); f[j + 5] += p.beta * bf16_hi(u.z);
                f[j + 6] += p.beta * bf16_lo(u.w); f[j + 7] += p.beta * bf16_hi(u.w);
              }
            } else {
              for (int j = 0; j < 32; ++j)
                if (n0 + j < p.n_out) f[j] += p.beta * __bfloat162float(rp[j]);
            }
          }
          if (p.act_post) {
#pragma unroll
            for (int j = 0; j < 32; ++j) f[j] = apply_act(f[j], p.act);
          }
          if (p.out_fp32) {
            float* op = reinterpret_cast<float*>(p.D) + (size_t)pix * p.ldd + n0;
            if (p.vec_ok && n0 + 31 < p.n_out) {
#pragma unroll
              for (int j = 0; j < 32; j += 4) *reinterpret_cast<float4*>(op + j) = make_float4(f[j], f[j + 1], f[j + 2], f[j + 3]);
            } else {
              for (int j = 0; j < 32; ++j)
                if (n0 + j < p.n_out) op[j] = f[j];
            }
          } else {
            __nv_bfloat16* op = reinterpret_cast<__nv_bfloat16*>(p.D) + (size_t)pix * p.ldd + n0;
            if (p.vec_ok && n0 + 31 < p.n_out) {
#pragma unroll
              for (int j = 0; j < 32; j += 8) {
                uint4 u;
                u.x = pack_bf16(f[j], f[j + 1]);
                u.y = pack_bf16(f[j + 2], f[j + 3]);
                u.z = pack_bf16(f[j + 4], f[j + 5]);
                u.w = pack_bf16(f[j + 6], f[j + 7]);
                *reinterpret_cast<uint4*>(op + j) = u;
              }
              if (p.stats_out) row_stats_add(f, st_sum, st_sq);
            } else {
              for (int j = 0; j < 32; ++j)
                if (n0 + j < p.n_out) op[j] = __float2bfloat16(f[j]);
            }
          }
        }  // row_ok
      }
      if (p.stats_out && row_ok) p.stats_out[(size_t)pix * p.stats_slots + 2 * n_blk + grp] = make_float2(st_sum, st_sq);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if constexpr (CTAS == 1) {
          mbar_arrive(&tempty[acc]);
        } else {
          mbar_arrive_cluster(mapa_rank(smem_u32(&tempty[acc]), 0));  // the MMA issuer lives in the leader CTA
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if constexpr (CTAS == 2) cluster_sync_all();  // neither CTA may exit while the pair's MMAs / commits can still touch it
  if (warp == 1) {
    tc_fence_after();
    if constexpr (CTAS == 1) {
      asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
    } else {
      asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Host side
// ------------------------------------------------------------------------------------------------
typedef CUresult (*PFN_tmapEncodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                        const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                        CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

PFN_tmapEncodeTiled get_encode_fn() {
  static PFN_tmapEncodeTiled fn = nullptr;
  if (!fn) {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_tmapEncodeTiled>(ptr);
  }
  return fn;
}

// rank-2 bf16 tensor [rows, cols] with row stride ld (elements); box = [box_rows, 64 cols], 128B swizzle.
int encode_2d(CUtensorMap* tm, const void* base, long long rows, long long cols, long long ld, int box_rows, int box_cols = BK,
              CUtensorMapSwizzle swz = CU_TENSOR_MAP_SWIZZLE_128B) {
  PFN_tmapEncodeTiled enc = get_encode_fn();
  if (!enc) {
    saspa_set_error("cuTensorMapEncodeTiled entry point unavailable (no CUDA driver?)");
    return SASPA_ERR_DRIVER;
  }
  cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t gstride[1] = {(cuuint64_t)ld * 2};
  cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), gdim, gstride, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    saspa_set_error("cuTensorMapEncodeTiled(2d rows=%lld cols=%lld ld=%lld box_rows=%d) failed: %d", rows, cols, ld, box_rows, (int)r);
    return SASPA_ERR_DRIVER;
  }
  return SASPA_OK;
}

// rank-4 NHWC bf16 activation [n, h, w, c] with pixel stride ld (elements); box = [bn, bh, bw, 64].
// estride > 1: the box walks the map with that element stride in w and h (a strided conv's A operand), i.e. it still delivers
// bn x bh x bw pixels but spans bh*estride x bw*estride input pixels.
int encode_nhwc(CUtensorMap* tm, const void* base, int n, int h, int w, int c, long long ld, int bn, int bh, int bw, int box_c = BK,
                CUtensorMapSwizzle swz = CU_TENSOR_MAP_SWIZZLE_128B, int estride = 1) {
  PFN_tmapEncodeTiled enc = get_encode_fn();
  if (!enc) {
    saspa_set_error("cuTensorMapEncodeTiled entry point unavailable (no CUDA driver?)");
    return SASPA_ERR_DRIVER;
  }
  cuuint64_t gdim[4] = {(cuuint64_t)c, (cuuint64_t)w, (cuuint64_t)h, (cuuint64_t)n};
  cuuint64_t gstride[3] = {(cuuint64_t)ld * 2, (cuuint64_t)ld * 2 * w, (cuuint64_t)ld * 2 * w * h};
  cuuint32_t box[4] = {(cuuint32_t)box_c, (cuuint32_t)(bw * estride), (cuuint32_t)(bh * estride), (cuuint32_t)bn};
  cuuint32_t estr[4] = {1, (cuuint32_t)estride, (cuuint32_t)estride, 1};
  CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), gdim, gstride, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    saspa_set_error("cuTensorMapEncodeTiled(nhwc n=%d h=%d w=%d c=%d ld=%lld box=%d,%d,%d) failed: %d", n, h, w, c, ld, bn, bh, bw, (int)r);
    return SASPA_ERR_DRIVER;
  }
  return SASPA_OK;
}

template <int BN, int CTAS>
int launch(const CUtensorMap& a0, const CUtensorMap& a1, const CUtensorMap& b, const CUtensorMap& d, const CUtensorMap& r, const GemmParams& p,
           cudaStream_t stream) {
  using C = Cfg<BN, CTAS>;
  static bool configured = false;
  if (!configured) {
    SASPA_CUDA(cudaFuncSetAttribute(gemm_tc_kernel<BN, CTAS>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES));
    configured = true;
  }
  const int total = ceil_div(p.num_m_tiles, CTAS) * p.num_n_tiles;  // tiles of CTAS x 128 rows
  const int max_clusters = saspa_num_sms() / CTAS;
  const int clusters = total < max_clusters ? total : max_clusters;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(clusters * CTAS, 1, 1);
  cfg.blockDim = dim3(NUM_THREADS, 1, 1);
  cfg.dynamicSmemBytes = C::SMEM_BYTES;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CTAS;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  SASPA_CUDA(cudaLaunchKernelEx(&cfg, gemm_tc_kernel<BN, CTAS>, a0, a1, b, d, r, p));
  return SASPA_OK;
}

// Tile width along N.  SD channel counts are multiples of 320 -> 160 tiles them exactly; the VAE /
// ResNets are powers of two; GEGLU needs value|gate halves in one tile (256 = 128 + 128).
std::atomic<int> g_force_bn{0};  // tuning hook (saspa_gemm_force_bn): 0 = heuristic

// Tile width along N.
//  * main-loop-bound problems (convs, K >= 2048): measured per-tile time relative to BN = 256 is 0.49 / 0.63 / 0.69 for
//    BN = 64 / 128 / 160 (profiles/r1_bn_sweep_y.txt, after the elected-issue fix; 0.72 / 0.72 / 0.80 while the narrow tiles were
//    issue-bound), so cost = waves x per-tile time decides -- the widest tile unless it
//    pads columns or leaves SMs idle in the last wave of a small-M problem;
//  * short-K GEMMs are bound by the epilogue / stores, where padded columns cost real time: exact tilings first
//    (SD channel counts are multiples of 320 -> 160; powers of two -> 256 / 128).
int pick_bn(int N, int act, int num_m_tiles, bool mainloop_bound, bool fixed_by_n = false) {
  if (act == SASPA_ACT_GEGLU) return 256;  // value | gate halves in one tile
  if (const int forced = g_force_bn.load()) return forced;
  // a GEMM that emits row statistics partitions each row's sums by N tile: the tile width must then be a function of N alone,
  // or a row's LayerNorm statistics would depend on how many other rows share the launch
  if (fixed_by_n) mainloop_bound = false;
  if (N <= 32) return 32;
  if (N <= 64) return 64;
  if (mainloop_bound) {
    const int cands[4] = {256, 160, 128, 64};
    const double t_rel[4] = {1.0, 0.69, 0.66, 0.49};  // 128: 0.63 on a compute-bound GEMM, worse on per-tap convolutions (operand bytes per flop)
    const int sms = saspa_num_sms();
    int best = 256;
    double best_cost = 1e30;
    for (int i = 0; i < 4; ++i) {
      const long long tiles = (long long)num_m_tiles * ceil_div(N, cands[i]);
      const double waves = tiles <= 4LL * sms ? (double)ceil_div_ll(tiles, sms) : (double)tiles / sms;
      const double cost = waves * t_rel[i];
      if (cost < best_cost * 0.999) {
        best_cost = cost;
        best = cands[i];
      }
    }
    return best;
  }
  if (N % 160 == 0 && N % 256 != 0) {
    // 160 tiles N exactly, but the wider tile moves fewer operand bytes per flop (these launches run at the L2 -> SM bandwidth:
    // (BM + BN) * 128 B per 2 * BM * BN * 64 flop): once the padding of the last 256-wide tile is under 1/12 of N the wide tile wins
    // -- N = 960 / 1920 (the fused q|k|v projections), measured 282 -> 222 us and 190 -> 157 us (profiles/r2_gemm_shape_sweep.txt)
    const int padded = ceil_div(N, 256) * 256 - N;
    return padded * 12 <= N ? 256 : 160;
  }
  if (N <= 128) return 128;
  if (N % 256 == 0) return 256;
  if (N % 128 == 0) return 128;
  int best = 256, best_waste = (ceil_div(N, 256) * 256 - N);
  const int cands[3] = {160, 128, 64};
  for (int i = 0; i < 3; ++i) {
    int waste = ceil_div(N, cands[i]) * cands[i] - N;
    if (waste < best_waste) {
      best = cands[i];
      best_waste = waste;
    }
  }
  return best;
}

std::atomic<int> g_reverse_m{0};   // tuning hook (saspa_gemm_reverse_m): 1 = GEMM row tiles in descending order
std::atomic<int> g_force_ctas{0};  // tuning hook (saspa_gemm_force_ctas): 0 = heuristic, 1 / 2 = CTAs per tile

// Two-CTA tiles (cta_group::2, 256 x BN): measured +14% at BN = 256 and +9% at BN = 160 on the plain GEMM main loop, a loss
// for narrower tiles and for the halo conv (profiles/r1_bn_sweep_y.txt), so only the wide long-K GEMM tiles pair up.
int pick_ctas(int num_m_tiles, int bn, int mode, int act, int K, bool has_residual) {
  if (const int forced = g_force_ctas.load()) return num_m_tiles >= 2 || forced == 1 ? forced : 1;
  // CTA pairs halve the B bytes each CTA pulls from L2.  They pay from K = 1024 up, and from K = 640 for the residual-carrying output
  // projections (64.5 vs 72.2 us at 65536 x 640 x 640); the other short-K launches (LayerNorm-folded, K = 320) are level with
  // single-CTA tiles since the hand-back lost its memory barrier (profiles/r2_pair_handback_membar.txt) and stay single
  const int k_min = has_residual ? 640 : 1024;
  if (bn >= 160 && mode == 0 && num_m_tiles >= 2) {
    // the GEGLU feed-forward GEMMs: 5 / 11 / 8 % faster as pairs at K = 640 / 1280 (16 x 16 and 8 x 8 levels), level at K = 320
    // (profiles/r2_pair_handback_membar.txt; they were excluded while the hand-back carried its barrier: 325 -> 381 us then)
    if (act == SASPA_ACT_GEGLU) return 2;
    if (K >= k_min) return 2;
  }
  // halo convolutions for which the wave model already chose the 256-wide tile (the 16 x 16 level): the pair tile is 3-5 % faster
  // (295 vs 311 us at 64 x 16 x 16, 1280 -> 1280; profiles/r2_conv_pair_sweep.txt); narrower tiles and per-tap convolutions are not
  return (mode == 2 && bn == 256 && num_m_tiles >= 2) ? 2 : 1;
}

template <int CTAS>
int dispatch_bn(int bn, const CUtensorMap& a0, const CUtensorMap& a1, const CUtensorMap& b, const CUtensorMap& d, const CUtensorMap& r,
                const GemmParams& p, cudaStream_t stream) {
  switch (bn) {
    case 32: return launch<32, CTAS>(a0, a1, b, d, r, p, stream);
    case 64: return launch<64, CTAS>(a0, a1, b, d, r, p, stream);
    case 128: return launch<128, CTAS>(a0, a1, b, d, r, p, stream);
    case 160: return launch<160, CTAS>(a0, a1, b, d, r, p, stream);
    case 256: return launch<256, CTAS>(a0, a1, b, d, r, p, stream);
  }
  saspa_set_error("internal: no kernel for BN=%d", bn);
  return SASPA_ERR_UNSUPPORTED;
}

int dispatch(int bn, int ctas, const CUtensorMap& a0, const CUtensorMap& a1, const CUtensorMap& b, const CUtensorMap& d, const CUtensorMap& r,
             const GemmParams& p, cudaStream_t stream) {
  return ctas == 2 ? dispatch_bn<2>(bn, a0, a1, b, d, r, p, stream) : dispatch_bn<1>(bn, a0, a1, b, d, r, p, stream);
}

std::atomic<int> g_conv_impl{0};  // 0 auto, 1 per-tap boxes only, 2 halo only (tests / A-B timing)

int fill_epilogue(GemmParams& p, const saspa_epilogue* ep, int N, void* D, int ldd) {
  static const saspa_epilogue kDefault = {nullptr, nullptr, 1, 0, SASPA_ACT_NONE, 1.0f, nullptr, 0, 0.0f, 0, 0, nullptr, 0, nullptr, 0, nullptr, 0.0f};
  if (!ep) ep = &kDefault;
  p.bias = ep->bias;
  p.row_bias = ep->row_bias;
  p.rows_per_group = ep->rows_per_group > 0 ? ep->rows_per_group : 1;
  p.ld_row_bias = ep->ld_row_bias > 0 ? ep->ld_row_bias : N;
  p.act = ep->act;
  p.act_post = (ep->act_after_residual && ep->act != SASPA_ACT_GEGLU) ? 1 : 0;
  p.alpha = ep->alpha;
  p.residual = static_cast<const __nv_bfloat16*>(ep->residual);
  p.ld_res = ep->ld_res;
  p.beta = ep->beta;
  p.out_fp32 = ep->out_fp32;
  p.D = D;
  p.ldd = ldd;
  p.n_out = (ep->act == SASPA_ACT_GEGLU) ? N / 2 : N;
  p.stats_out = static_cast<float2*>(ep->row_stats_out);
  p.stats_slots = ep->row_stats_slots;
  p.ln_stats = static_cast<const float2*>(ep->ln_stats);
  p.ln_slots = ep->ln_slots;
  p.ln_colsum = ep->ln_colsum;
  p.ln_eps = ep->ln_eps;
  SASPA_CHECK_ARG(!ep->row_stats_out || (!ep->out_fp32 && ep->act != SASPA_ACT_GEGLU && N % 32 == 0 && (reinterpret_cast<uintptr_t>(ep->row_stats_out) & 7) == 0),
                  "epilogue: row_stats_out needs a bf16, non-GEGLU output with N %% 32 == 0 (N=%d)", N);
  SASPA_CHECK_ARG(!ep->ln_stats || (ep->ln_colsum && ep->ln_slots > 0 && ep->ln_slots <= 64 && (reinterpret_cast<uintptr_t>(ep->ln_colsum) & 15) == 0 &&
                                    (reinterpret_cast<uintptr_t>(ep->ln_stats) & 7) == 0),
                  "epilogue: ln_stats needs ln_colsum (16-byte aligned) and 1 <= ln_slots <= 64 (got %d)", ep->ln_slots);
  SASPA_CHECK_ARG(ep->act >= SASPA_ACT_NONE && ep->act <= SASPA_ACT_GEGLU, "epilogue: unknown activation %d", ep->act);
  SASPA_CHECK_ARG(!(ep->act == SASPA_ACT_GEGLU && (N % 256 != 0)), "GEGLU epilogue needs N %% 256 == 0 (tile-interleaved weights), got %d", N);
  SASPA_CHECK_ARG(!(ep->act == SASPA_ACT_GEGLU && ep->row_bias), "GEGLU epilogue does not take a row_bias");
  const int esz = ep->out_fp32 ? 4 : 2;
  bool vec = ((reinterpret_cast<uintptr_t>(D) & 15) == 0) && (((size_t)ldd * esz) % 16 == 0);
  if (ep->residual) vec = vec && ((reinterpret_cast<uintptr_t>(ep->residual) & 15) == 0) && (ep->ld_res % 8 == 0);
  if (ep->bias) vec = vec && ((reinterpret_cast<uintptr_t>(ep->bias) & 15) == 0);
  if (ep->row_bias) vec = vec && ((reinterpret_cast<uintptr_t>(ep->row_bias) & 15) == 0) && (N % 4 == 0) && (p.ld_row_bias % 4 == 0);
  p.vec_ok = vec ? 1 : 0;
  // staged TMA epilogue: bf16 output whose rows (and the residual's) are 16-byte aligned / strided
  p.tma_store = (!ep->out_fp32 && ((reinterpret_cast<uintptr_t>(D) & 15) == 0) && (ldd % 8 == 0) &&
                 (!ep->residual || (((reinterpret_cast<uintptr_t>(ep->residual) & 15) == 0) && (ep->ld_res % 8 == 0))))
                    ? 1
                    : 0;
  // the float4 bias loads assume 16-byte aligned bias pointers; fall back is per-element only when n+3 >= N
  SASPA_CHECK_ARG(!ep->bias || (reinterpret_cast<uintptr_t>(ep->bias) & 15) == 0, "epilogue: bias must be 16-byte aligned");
  SASPA_CHECK_ARG(!ep->row_bias || ((reinterpret_cast<uintptr_t>(ep->row_bias) & 15) == 0 && N % 4 == 0 && p.ld_row_bias % 4 == 0),
                  "epilogue: row_bias must be 16-byte aligned with N %% 4 == 0 and ld_row_bias %% 4 == 0");
  return SASPA_OK;
}

}  // namespace

extern "C" int saspa_gemm_bf16(const void* A, int lda, const void* B, int ldb, void* D, int ldd, int M, int N, int K,
                               const saspa_epilogue* ep, cudaStream_t stream) {
  SASPA_CHECK_ARG(M >= 0 && N >= 0 && K >= 0, "saspa_gemm_bf16: negative dims");
  if (M == 0 || N == 0) return SASPA_OK;
  SASPA_CHECK_ARG(A && B && D, "saspa_gemm_bf16: null pointer");
  SASPA_CHECK_ARG(K > 0, "saspa_gemm_bf16: K must be positive");
  SASPA_CHECK_ARG(lda % 8 == 0 && ldb % 8 == 0 && lda >= K && ldb >= K, "saspa_gemm_bf16: lda/ldb must be >= K and multiples of 8 (lda=%d ldb=%d K=%d)", lda, ldb, K);
  SASPA_CHECK_ARG(((reinterpret_cast<uintptr_t>(A) | reinterpret_cast<uintptr_t>(B)) & 15) == 0, "saspa_gemm_bf16: A/B must be 16-byte aligned");
  GemmParams p = {};
  int rc = fill_epilogue(p, ep, N, D, ldd);
  if (rc) return rc;
  p.M = M;
  p.N = N;
  p.K = K;
  p.mode = 0;
  p.m_rev = g_reverse_m.load();
  p.num_m_tiles = ceil_div(M, BM);
  const int bn = pick_bn(N, p.act, p.num_m_tiles, K >= 2048, p.stats_out != nullptr);
  p.num_n_tiles = ceil_div(N, bn);
  p.ln_inv_k = 1.0f / (float)K;
  SASPA_CHECK_ARG(!p.stats_out || p.stats_slots == 2 * p.num_n_tiles, "saspa_gemm_bf16: row_stats_slots must be saspa_gemm_row_stats_slots(N) = %d, got %d",
                  2 * p.num_n_tiles, p.stats_slots);
  CUtensorMap tmA, tmB;
  if ((rc = encode_2d(&tmA, A, M, K, lda, BM))) return rc;
  const int ctas = pick_ctas(p.num_m_tiles, bn, 0, p.act, K, p.residual != nullptr);
  if ((rc = encode_2d(&tmB, B, N, K, ldb, bn / ctas))) return rc;  // each CTA of a pair loads its share of the B tile
  CUtensorMap tmD = tmA, tmR = tmA;
  if (p.tma_store) {
    if ((rc = encode_2d(&tmD, D, M, p.n_out, ldd, BM, PANEL, CU_TENSOR_MAP_SWIZZLE_64B))) return rc;
    if (p.residual && (rc = encode_2d(&tmR, p.residual, M, p.n_out, p.ld_res, BM, PANEL, CU_TENSOR_MAP_SWIZZLE_64B))) return rc;
  }
  return dispatch(bn, ctas, tmA, tmA, tmB, tmD, tmR, p, stream);
}

// Slots per row of the partial row statistics a GEMM with N output columns writes (2 epilogue groups x N tiles; the tile width is
// then a function of N alone, see pick_bn).
extern "C" int saspa_gemm_row_stats_slots(int N) {
  if (N <= 0) return 0;
  return 2 * ceil_div(N, pick_bn(N, SASPA_ACT_NONE, 1, false, true));
}

extern "C" int saspa_conv2d_igemm_strided_bf16(const void* x0, int ldx0, int c0, const void* x1, int ldx1, int c1, int n, int ih, int iw,
                                               const void* weight, int ksize, int stride, int pad, int h, int w, void* out, int ldo, int cout,
                                               const saspa_epilogue* ep, cudaStream_t stream) {
  SASPA_CHECK_ARG(n >= 0 && h >= 0 && w >= 0 && ih >= 0 && iw >= 0 && cout >= 0, "saspa_conv2d_igemm_bf16: negative dims");
  if (n == 0 || h == 0 || w == 0 || cout == 0) return SASPA_OK;
  SASPA_CHECK_ARG(x0 && weight && out, "saspa_conv2d_igemm_bf16: null pointer");
  SASPA_CHECK_ARG(ksize == 1 || ksize == 3, "saspa_conv2d_igemm_bf16: ksize must be 1 or 3, got %d", ksize);
  SASPA_CHECK_ARG(stride == 1 || stride == 2, "saspa_conv2d_igemm_bf16: stride must be 1 or 2, got %d", stride);
  SASPA_CHECK_ARG(pad >= 0 && pad <= ksize / 2, "saspa_conv2d_igemm_bf16: top/left padding must be in [0, ksize/2], got %d", pad);
  SASPA_CHECK_ARG((long long)(h - 1) * stride + ksize - pad <= ih + ksize / 2 + 1 && (long long)(w - 1) * stride + ksize - pad <= iw + ksize / 2 + 1,
                  "saspa_conv2d_igemm_bf16: output map %dx%d reaches beyond the zero-padded %dx%d input", h, w, ih, iw);
  if (!x1) c1 = 0;
  SASPA_CHECK_ARG(c0 > 0 && c0 % 8 == 0 && c1 % 8 == 0, "saspa_conv2d_igemm_bf16: channel counts must be multiples of 8 (c0=%d c1=%d)", c0, c1);
  SASPA_CHECK_ARG(c1 == 0 || c0 % BK == 0, "saspa_conv2d_igemm_bf16: c0 must be a multiple of 64 when a second source is given (c0=%d)", c0);
  SASPA_CHECK_ARG(ldx0 % 8 == 0 && ldx0 >= c0 && (c1 == 0 || (ldx1 % 8 == 0 && ldx1 >= c1)), "saspa_conv2d_igemm_bf16: bad pixel strides");
  SASPA_CHECK_ARG((reinterpret_cast<uintptr_t>(x0) & 15) == 0 && (reinterpret_cast<uintptr_t>(weight) & 15) == 0 &&
                      (c1 == 0 || (reinterpret_cast<uintptr_t>(x1) & 15) == 0),
                  "saspa_conv2d_igemm_bf16: 16-byte alignment required");
  const bool same = stride == 1 && pad == ksize / 2 && ih == h && iw == w;
  GemmParams p = {};
  int rc = fill_epilogue(p, ep, cout, out, ldo);
  if (rc) return rc;
  SASPA_CHECK_ARG(!p.stats_out && !p.ln_stats, "saspa_conv2d_igemm_bf16: row statistics / folded LayerNorm are GEMM-only epilogue options");
  // M tile = bw x bh x bnimg = 128 output pixels; minimise padded work, tie-break towards square tiles.
  int best_bw = 0, best_bh = 0, best_bi = 0;
  long long best_cost = -1;
  for (int bw = 128; bw >= 1; bw >>= 1) {
    int bh = 128 / bw;
    int bi = 1;
    // shrink bh to the image height (power of two) and spill into the batch dimension
    while (bh > 1 && bh / 2 >= h) {
      bh >>= 1;
      bi <<= 1;
    }
    if (bw > 1 && bw / 2 >= w) continue;  // a narrower box covers the row just as well
    if (bw * stride > 256 || bh * stride > 256) continue;  // TMA box extent limit
    long long cost = (long long)ceil_div(w, bw) * bw * ceil_div(h, bh) * bh * ceil_div(n, bi) * bi;
    long long halo = (long long)(bw + 2) * (bh + 2);  // L2 traffic proxy
    long long score = cost * 1024 + halo;
    if (best_cost < 0 || score < best_cost) {
      best_cost = score;
      best_bw = bw;
      best_bh = bh;
      best_bi = bi;
    }
  }
  // 3x3 on maps of at least one 8 x 16 tile: halo mode (the activation is fetched once per 64-channel chunk
  // instead of once per tap; operand traffic out of L2 is what bounds this kernel)
  const bool halo = same && ksize == 3 && h >= HALO_BH && w >= HALO_BW && g_conv_impl != 1;
  if (g_conv_impl == 2 && !halo) {
    saspa_set_error("saspa_conv2d_igemm_bf16: halo mode forced but the shape is not eligible (ksize=%d h=%d w=%d)", ksize, h, w);
    return SASPA_ERR_UNSUPPORTED;
  }
  if (halo) {
    best_bw = HALO_BW;
    best_bh = HALO_BH;
    best_bi = 1;
  }
  p.mode = halo ? 2 : 1;
  p.n_img = n;
  p.H = h;
  p.W = w;
  p.c0 = c0;
  p.c1 = c1;
  p.ksize = ksize;
  p.stride = stride;
  p.pad = pad;
  p.bw = best_bw;
  p.bh = best_bh;
  p.bn = best_bi;
  p.tiles_x = ceil_div(w, p.bw);
  p.tiles_y = ceil_div(h, p.bh);
  const int tiles_n = ceil_div(n, p.bn);
  p.num_m_tiles = p.tiles_x * p.tiles_y * tiles_n;
  const int bn_tile = pick_bn(cout, p.act, p.num_m_tiles, ksize == 3 || c0 + c1 >= 2048);
  p.num_n_tiles = ceil_div(cout, bn_tile);
  p.N = cout;
  p.K = ksize * ksize * (c0 + c1);
  p.M = n * h * w;
  CUtensorMap tmA0, tmA1, tmB;
  const int abh = halo ? p.bh + 2 : p.bh, abw = halo ? p.bw + 2 : p.bw;  // activation box (with the 1-px halo in mode 2)
  if ((rc = encode_nhwc(&tmA0, x0, n, ih, iw, c0, ldx0, p.bn, abh, abw, BK, CU_TENSOR_MAP_SWIZZLE_128B, stride))) return rc;
  if (c1 > 0) {
    if ((rc = encode_nhwc(&tmA1, x1, n, ih, iw, c1, ldx1, p.bn, abh, abw, BK, CU_TENSOR_MAP_SWIZZLE_128B, stride))) return rc;
  } else {
    tmA1 = tmA0;
  }
  const int ctas = pick_ctas(p.num_m_tiles, bn_tile, p.mode, p.act, p.K, p.residual != nullptr);
  if ((rc = encode_2d(&tmB, weight, cout, p.K, p.K, bn_tile / ctas))) return rc;
  CUtensorMap tmD = tmA0, tmR = tmA0;
  if (p.tma_store) {
    if ((rc = encode_nhwc(&tmD, out, n, h, w, p.n_out, ldo, p.bn, p.bh, p.bw, PANEL, CU_TENSOR_MAP_SWIZZLE_64B))) return rc;
    if (p.residual && (rc = encode_nhwc(&tmR, p.residual, n, h, w, p.n_out, p.ld_res, p.bn, p.bh, p.bw, PANEL, CU_TENSOR_MAP_SWIZZLE_64B)))
      return rc;
  }
  return dispatch(bn_tile, ctas, tmA0, tmA1, tmB, tmD, tmR, p, stream);
}

extern "C" int saspa_conv2d_igemm_bf16(const void* x0, int ldx0, int c0, const void* x1, int ldx1, int c1, int n, int h, int w,
                                       const void* weight, int ksize, void* out, int ldo, int cout, const saspa_epilogue* ep,
                                       cudaStream_t stream) {
  return saspa_conv2d_igemm_strided_bf16(x0, ldx0, c0, x1, ldx1, c1, n, h, w, weight, ksize, 1, ksize / 2, h, w, out, ldo, cout, ep, stream);
}

extern "C" int saspa_conv_impl(int impl) {
  const int prev = g_conv_impl.load();
  if (impl >= 0 && impl <= 2) g_conv_impl = impl;
  return prev;
}

// Tuning hook: force the N tile width of the non-GEGLU kernels (0 restores the heuristic).  Not part of the product API.
extern "C" int saspa_gemm_force_bn(int bn) {
  const int prev = g_force_bn.load();
  if (bn == 0 || bn == 32 || bn == 64 || bn == 128 || bn == 160 || bn == 256) g_force_bn = bn;
  return prev;
}

// Tuning hook: force one- or two-CTA tiles (0 restores the heuristic).  Not part of the product API.
extern "C" int saspa_gemm_reverse_m(int on) {
  const int prev = g_reverse_m.load();
  if (on == 0 || on == 1) g_reverse_m = on;
  return prev;
}

extern "C" int saspa_gemm_force_ctas(int ctas) {
  const int prev = g_force_ctas.load();
  if (ctas >= 0 && ctas <= 2) g_force_ctas = ctas;
  return prev;
}
