// Fused softmax(Q K^T * scale) V for the UNet / ControlNet transformer blocks (self-attention over
// 4096/1024/256/64 image tokens with head_dim 40/80/160, cross-attention over 77 text tokens) and the
// CLIP / Q-Former towers (head_dim 64).  Replaces F.scaled_dot_product_attention under diffusers'
// AttnProcessor2_0 (pipe(**pipe_args), run_aug/run_aug.py:278).
//
// Round-1 implementation: flash-style single pass, 64 queries x 64 keys per CTA iteration, 4 warps
// x 16 query rows, bf16 mma.sync.m16n8k16 with fp32 accumulators and online softmax (exp2 domain),
// K/V tiles double-buffered with cp.async, head_dim zero-padded to a multiple of 16 in shared
// memory only (never in HBM), rows padded by 16 B so ldmatrix is bank-conflict free.
// (A tcgen05/TMEM version with S/P resident in TMEM is the planned replacement; see DESIGN.md.)
//
// Also here: row softmax + 2-D transpose used by the d = 512 single-head VAE mid-block attention,
// which runs as GEMM -> softmax -> GEMM on the tcgen05 GEMM kernel.
#include "common.cuh"
#include <atomic>

#include "../../include/saspa_b200.h"
#include "tuning_hooks.h"

namespace {

constexpr int BQ = 64, BKV = 64, ATT_THREADS = 128;

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem, bool pred) {
  uint32_t s = (uint32_t)__cvta_generic_to_shared(smem);
  int sz = pred ? 16 : 0;  // src-size 0 -> zero fill
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(s), "l"(gmem), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void ldmatrix_x4(uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3, const void* p) {
  uint32_t s = (uint32_t)__cvta_generic_to_shared(p);
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(s));
}
__device__ __forceinline__ void ldmatrix_x4_trans(uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3, const void* p) {
  uint32_t s = (uint32_t)__cvta_generic_to_shared(p);
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(s));
}
__device__ __forceinline__ void mma_bf16_16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// Stage a [rows x d] bf16 tile (row stride ld elements in gmem) into smem rows of DP*2+16 bytes,
// zero-filling rows >= valid_rows and columns in [d, DP).
__device__ __forceinline__ float ex2_approx_f(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

template <int DP>
__device__ __forceinline__ void load_tile(__nv_bfloat16* s, const __nv_bfloat16* g, long long ld, int valid_rows, int d, int tid) {
  constexpr int ROWB = DP * 2 + 16;
  constexpr int CH = DP / 8;  // 16-byte chunks per row
  for (int i = tid; i < 64 * CH; i += ATT_THREADS) {
    int r = i / CH, c = i % CH;
    bool ok = (r < valid_rows) && (c * 8 < d);
    const __nv_bfloat16* src = ok ? g + (long long)r * ld + c * 8 : g;
    cp_async16(reinterpret_cast<uint8_t*>(s) + r * ROWB + c * 16, src, ok);
  }
}

template <int DP>
__global__ void __launch_bounds__(ATT_THREADS) flash_attn_kernel(const __nv_bfloat16* __restrict__ q, int ldq, const __nv_bfloat16* __restrict__ k,
                                                                  int ldk, const __nv_bfloat16* __restrict__ v, int ldv,
                                                                  __nv_bfloat16* __restrict__ o, int ldo, int heads, int tq, int tkv, int d,
                                                                  float scale_log2, int causal) {
  constexpr int ROWB = DP * 2 + 16;
  constexpr int KS = DP / 16;  // k-steps over the head dim
  constexpr int NT = DP / 8;   // n-tiles of the output
  extern __shared__ __align__(16) uint8_t att_smem[];
  __nv_bfloat16* sQ = reinterpret_cast<__nv_bfloat16*>(att_smem);
  uint8_t* sK = att_smem + 64 * ROWB;      // 2 buffers
  uint8_t* sV = sK + 2 * 64 * ROWB;        // 2 buffers

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int bh = blockIdx.y, b = bh / heads, h = bh % heads;
  const int q0 = blockIdx.x * BQ;
  const __nv_bfloat16* qg = q + ((long long)b * tq + q0) * ldq + (long long)h * d;
  const __nv_bfloat16* kg = k + ((long long)b * tkv) * ldk + (long long)h * d;
  const __nv_bfloat16* vg = v + ((long long)b * tkv) * ldv + (long long)h * d;

  load_tile<DP>(sQ, qg, ldq, min(BQ, tq - q0), d, tid);
  load_tile<DP>(reinterpret_cast<__nv_bfloat16*>(sK), kg, ldk, min(BKV, tkv), d, tid);
  load_tile<DP>(reinterpret_cast<__nv_bfloat16*>(sV), vg, ldv, min(BKV, tkv), d, tid);
  cp_async_commit();

  int n_iter = (tkv + BKV - 1) / BKV;
  if (causal) n_iter = min(n_iter, (min(q0 + BQ, tq) + BKV - 1) / BKV);  // keys beyond the last query row are masked
  uint32_t qf[KS][4];
  float oacc[NT][4];
#pragma unroll
  for (int i = 0; i < NT; ++i) oacc[i][0] = oacc[i][1] = oacc[i][2] = oacc[i][3] = 0.0f;
  float m_run[2] = {-INFINITY, -INFINITY}, l_run[2] = {0.0f, 0.0f};

  for (int it = 0; it < n_iter; ++it) {
    const int buf = it & 1;
    if (it + 1 < n_iter) {
      const int kv0 = (it + 1) * BKV;
      load_tile<DP>(reinterpret_cast<__nv_bfloat16*>(sK + (buf ^ 1) * 64 * ROWB), kg + (long long)kv0 * ldk, ldk, min(BKV, tkv - kv0), d, tid);
      load_tile<DP>(reinterpret_cast<__nv_bfloat16*>(sV + (buf ^ 1) * 64 * ROWB), vg + (long long)kv0 * ldv, ldv, min(BKV, tkv - kv0), d, tid);
      cp_async_commit();
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();

    if (it == 0) {
      // Q fragments for this warp's 16 rows stay in registers for the whole KV sweep.
#pragma unroll
      for (int ks = 0; ks < KS; ++ks) {
        int row = warp * 16 + (lane & 7) + ((lane >> 3) & 1) * 8;
        int col = ks * 16 + (lane >> 4) * 8;
        ldmatrix_x4(qf[ks][0], qf[ks][1], qf[ks][2], qf[ks][3], reinterpret_cast<uint8_t*>(sQ) + row * ROWB + col * 2);
      }
    }

    const uint8_t* kb = sK + buf * 64 * ROWB;
    const uint8_t* vb = sV + buf * 64 * ROWB;

    // ---- S = Q K^T (16 x 64 per warp) ----
    float sacc[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i) sacc[i][0] = sacc[i][1] = sacc[i][2] = sacc[i][3] = 0.0f;
#pragma unroll
    for (int ks = 0; ks < KS; ++ks) {
#pragma unroll
      for (int np = 0; np < 4; ++np) {  // pairs of 8-key n-tiles
        uint32_t b0, b1, b2, b3;
        int key = np * 16 + (lane & 7) + (lane >> 4) * 8;
        int col = ks * 16 + ((lane >> 3) & 1) * 8;
        ldmatrix_x4(b0, b1, b2, b3, kb + key * ROWB + col * 2);
        mma_bf16_16816(sacc[np * 2], qf[ks], b0, b1);
        mma_bf16_16816(sacc[np * 2 + 1], qf[ks], b2, b3);
      }
    }

    // ---- online softmax (rows g = lane/4 and g+8; this thread owns cols (lane%4)*2 + {0,1} of each n-tile) ----
    const int kv0 = it * BKV;
    float mx[2] = {-INFINITY, -INFINITY};
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        int key = kv0 + nt * 8 + (lane & 3) * 2 + (e & 1);
        bool ok = key < tkv;
        if (causal) ok = ok && (key <= q0 + warp * 16 + (lane >> 2) + (e >> 1) * 8);
        float sv = ok ? sacc[nt][e] * scale_log2 : -INFINITY;
        sacc[nt][e] = sv;
        mx[e >> 1] = fmaxf(mx[e >> 1], sv);
      }
    }
    float corr[2];
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 1));
      mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 2));
      float m_new = fmaxf(m_run[r], mx[r]);
      corr[r] = (m_run[r] == -INFINITY) ? 0.0f : exp2f(m_run[r] - m_new);
      m_run[r] = m_new;
      if (m_new == -INFINITY) m_new = 0.0f;  // fully masked so far (padding rows): keep exp2 arguments finite
    }
    float rs[2] = {0.0f, 0.0f};
    uint32_t pf[4][4];  // P as A-fragments for 4 k-steps of 16 keys
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      const float mr0 = (m_run[0] == -INFINITY) ? 0.0f : m_run[0], mr1 = (m_run[1] == -INFINITY) ? 0.0f : m_run[1];
      float p0 = exp2f(sacc[nt][0] - mr0);
      float p1 = exp2f(sacc[nt][1] - mr0);
      float p2 = exp2f(sacc[nt][2] - mr1);
      float p3 = exp2f(sacc[nt][3] - mr1);
      rs[0] += p0 + p1;
      rs[1] += p2 + p3;
      pf[nt >> 1][(nt & 1) * 2 + 0] = pack_bf16(p0, p1);
      pf[nt >> 1][(nt & 1) * 2 + 1] = pack_bf16(p2, p3);
    }
#pragma unroll
    for (int r = 0; r < 2; ++r) l_run[r] = l_run[r] * corr[r] + rs[r];
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {
      oacc[nt][0] *= corr[0];
      oacc[nt][1] *= corr[0];
      oacc[nt][2] *= corr[1];
      oacc[nt][3] *= corr[1];
    }

    // ---- O += P V ----
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {  // 16 keys per step
#pragma unroll
      for (int np = 0; np < NT / 2; ++np) {  // pairs of 8-wide d tiles
        uint32_t b0, b1, b2, b3;
        int key = ks * 16 + (lane & 7) + ((lane >> 3) & 1) * 8;
        int col = np * 16 + (lane >> 4) * 8;
        ldmatrix_x4_trans(b0, b1, b2, b3, vb + key * ROWB + col * 2);
        mma_bf16_16816(oacc[np * 2], pf[ks], b0, b1);
        mma_bf16_16816(oacc[np * 2 + 1], pf[ks], b2, b3);
      }
    }
    __syncthreads();  // everyone done with buf before the next iteration's prefetch overwrites it
  }

  // ---- finalize ----
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    l_run[r] += __shfl_xor_sync(0xffffffffu, l_run[r], 1);
    l_run[r] += __shfl_xor_sync(0xffffffffu, l_run[r], 2);
  }
  const float inv0 = l_run[0] > 0.0f ? 1.0f / l_run[0] : 0.0f;
  const float inv1 = l_run[1] > 0.0f ? 1.0f / l_run[1] : 0.0f;
  const int r0 = q0 + warp * 16 + (lane >> 2);
  __nv_bfloat16* og = o + ((long long)b * tq) * ldo + (long long)h * d;
#pragma unroll
  for (int nt = 0; nt < NT; ++nt) {
    int col = nt * 8 + (lane & 3) * 2;
    if (col < d) {
      if (r0 < tq) *reinterpret_cast<uint32_t*>(og + (long long)r0 * ldo + col) = pack_bf16(oacc[nt][0] * inv0, oacc[nt][1] * inv0);
      if (r0 + 8 < tq) *reinterpret_cast<uint32_t*>(og + (long long)(r0 + 8) * ldo + col) = pack_bf16(oacc[nt][2] * inv1, oacc[nt][3] * inv1);
    }
  }
}

template <int DP>
int launch_flash(const void* q, int ldq, const void* k, int ldk, const void* v, int ldv, void* o, int ldo, int batch, int heads, int tq,
                 int tkv, int d, float scale, int causal, cudaStream_t stream) {
  constexpr int ROWB = DP * 2 + 16;
  constexpr int SMEM = 5 * 64 * ROWB;
  static bool configured = false;
  if (!configured) {
    SASPA_CUDA(cudaFuncSetAttribute(flash_attn_kernel<DP>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
    configured = true;
  }
  dim3 grid(ceil_div(tq, BQ), batch * heads);
  flash_attn_kernel<DP><<<grid, ATT_THREADS, SMEM, stream>>>(static_cast<const __nv_bfloat16*>(q), ldq, static_cast<const __nv_bfloat16*>(k), ldk,
                                                             static_cast<const __nv_bfloat16*>(v), ldv, static_cast<__nv_bfloat16*>(o), ldo, heads,
                                                             tq, tkv, d, scale * 1.4426950408889634f, causal);
  SASPA_LAUNCH_CHECK();
  return SASPA_OK;
}

// ------------------------------------------------------------------------------------------------
// Cross-attention over a SHORT key sequence (tkv <= 128: the 77 text tokens), non-causal.
// The work is HBM-bound (read Q, write O; 4*tq*77*d FLOP is nothing), and the per-CTA fixed cost dominated the
// generic kernels (one CTA per 64 / 256 queries re-staging K and V).  Here K and V of one (batch, head) are staged in
// shared memory ONCE per CTA (padded to 80 or 128 keys) and the CTA then streams `qt_per_cta` consecutive 64-query tiles
// through them: Q tiles double-buffered with cp.async, the whole score row / one-pass softmax / P V in registers
// (mma.sync m16n8k16, 16 query rows per warp), O staged through the consumed Q buffer so it leaves as 16-byte
// row-contiguous stores.
// ------------------------------------------------------------------------------------------------
template <int DP, int NK16>  // NK16 = key groups of 16 held resident (5 -> up to 80 keys: the 77 text tokens; 8 -> up to 128)
__global__ void __launch_bounds__(ATT_THREADS) xattn_resident_kernel(const __nv_bfloat16* __restrict__ q, int ldq, const __nv_bfloat16* __restrict__ k,
                                                                      int ldk, const __nv_bfloat16* __restrict__ v, int ldv,
                                                                      __nv_bfloat16* __restrict__ o, int ldo, int heads, int tq, int tkv, int d,
                                                                      float scale_log2, int qt_per_cta) {
  constexpr int ROWB = DP * 2 + 16;
  constexpr int KS = DP / 16;
  constexpr int NT = DP / 8;
  constexpr int NKEY = NK16 * 16;
  constexpr int CH = DP / 8;
  extern __shared__ __align__(16) uint8_t att_smem[];
  uint8_t* sQ = att_smem;                  // 2 buffers of 64 rows
  uint8_t* sK = att_smem + 2 * 64 * ROWB;  // NKEY rows (keys >= tkv zero-filled)
  uint8_t* sV = sK + NKEY * ROWB;          // NKEY rows

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int bh = blockIdx.y, b = bh / heads, h = bh % heads;
  const int n_qt = (tq + BQ - 1) / BQ;
  const int qt0 = blockIdx.x * qt_per_cta, qt1 = min(qt0 + qt_per_cta, n_qt);
  const __nv_bfloat16* qg = q + ((long long)b * tq) * ldq + (long long)h * d;
  const __nv_bfloat16* kg = k + ((long long)b * tkv) * ldk + (long long)h * d;
  const __nv_bfloat16* vg = v + ((long long)b * tkv) * ldv + (long long)h * d;
  __nv_bfloat16* og = o + ((long long)b * tq) * ldo + (long long)h * d;

  for (int i = tid; i < NKEY * CH; i += ATT_THREADS) {
    const int r = i / CH, c = i % CH;
    const bool ok = (r < tkv) && (c * 8 < d);
    cp_async16(sK + r * ROWB + c * 16, ok ? kg + (long long)r * ldk + c * 8 : kg, ok);
    cp_async16(sV + r * ROWB + c * 16, ok ? vg + (long long)r * ldv + c * 8 : vg, ok);
  }
  load_tile<DP>(reinterpret_cast<__nv_bfloat16*>(sQ), qg + (long long)qt0 * BQ * ldq, ldq, min(BQ, tq - qt0 * BQ), d, tid);
  cp_async_commit();

  for (int qt = qt0; qt < qt1; ++qt) {
    const int buf = (qt - qt0) & 1;
    const int q0 = qt * BQ;
    if (qt + 1 < qt1) {
      load_tile<DP>(reinterpret_cast<__nv_bfloat16*>(sQ + (buf ^ 1) * 64 * ROWB), qg + (long long)(q0 + BQ) * ldq, ldq, min(BQ, tq - q0 - BQ), d, tid);
      cp_async_commit();
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();
    uint8_t* qb = sQ + buf * 64 * ROWB;
    uint32_t qf[KS][4];
#pragma unroll
    for (int ks = 0; ks < KS; ++ks) {
      int row = warp * 16 + (lane & 7) + ((lane >> 3) & 1) * 8;
      int col = ks * 16 + (lane >> 4) * 8;
      ldmatrix_x4(qf[ks][0], qf[ks][1], qf[ks][2], qf[ks][3], qb + row * ROWB + col * 2);
    }
    // ---- S = Q K^T over all resident keys (16 x NKEY per warp), one softmax pass, no running rescale ----
    float sacc[NK16 * 2][4];
#pragma unroll
    for (int i = 0; i < NK16 * 2; ++i) sacc[i][0] = sacc[i][1] = sacc[i][2] = sacc[i][3] = 0.0f;
#pragma unroll
    for (int ks = 0; ks < KS; ++ks) {
#pragma unroll
      for (int np = 0; np < NK16; ++np) {
        uint32_t b0, b1, b2, b3;
        int key = np * 16 + (lane & 7) + (lane >> 4) * 8;
        int col = ks * 16 + ((lane >> 3) & 1) * 8;
        ldmatrix_x4(b0, b1, b2, b3, sK + key * ROWB + col * 2);
        mma_bf16_16816(sacc[np * 2], qf[ks], b0, b1);
        mma_bf16_16816(sacc[np * 2 + 1], qf[ks], b2, b3);
      }
    }
    // The kernel is issue-bound (ncu: 70 % of the issue slots, profiles/r1_ncu_full_xattn.txt), so the softmax is kept to four
    // instructions per score: row max on the raw accumulators (only the n-tiles that reach past tkv are masked), then one FFMA
    // (scale, minus the scaled max) and one ex2.approx per element.
    float mx[2] = {-INFINITY, -INFINITY};
#pragma unroll
    for (int nt = 0; nt < NK16 * 2; ++nt) {
      if (nt * 8 + 8 > tkv) {  // warp-uniform: only the last n-tile(s) hold keys >= tkv (zero-filled K rows)
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int key = nt * 8 + (lane & 3) * 2 + (e & 1);
          if (key >= tkv) sacc[nt][e] = -INFINITY;
        }
      }
      mx[0] = fmaxf(mx[0], fmaxf(sacc[nt][0], sacc[nt][1]));
      mx[1] = fmaxf(mx[1], fmaxf(sacc[nt][2], sacc[nt][3]));
    }
    float nm[2];
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 1));
      mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 2));
      nm[r] = (mx[r] == -INFINITY) ? 0.0f : -mx[r] * scale_log2;  // scale_log2 > 0: max commutes with the scaling
    }
    float rs[2] = {0.0f, 0.0f};
    uint32_t pf[NK16][4];
#pragma unroll
    for (int nt = 0; nt < NK16 * 2; ++nt) {
      const float p0 = ex2_approx_f(fmaf(sacc[nt][0], scale_log2, nm[0]));
      const float p1 = ex2_approx_f(fmaf(sacc[nt][1], scale_log2, nm[0]));
      const float p2 = ex2_approx_f(fmaf(sacc[nt][2], scale_log2, nm[1]));
      const float p3 = ex2_approx_f(fmaf(sacc[nt][3], scale_log2, nm[1]));
      rs[0] += p0 + p1;
      rs[1] += p2 + p3;
      pf[nt >> 1][(nt & 1) * 2 + 0] = pack_bf16(p0, p1);
      pf[nt >> 1][(nt & 1) * 2 + 1] = pack_bf16(p2, p3);
    }
    float oacc[NT][4];
#pragma unroll
    for (int i = 0; i < NT; ++i) oacc[i][0] = oacc[i][1] = oacc[i][2] = oacc[i][3] = 0.0f;
#pragma unroll
    for (int ks = 0; ks < NK16; ++ks) {
#pragma unroll
      for (int np = 0; np < NT / 2; ++np) {
        uint32_t b0, b1, b2, b3;
        int key = ks * 16 + (lane & 7) + ((lane >> 3) & 1) * 8;
        int col = np * 16 + (lane >> 4) * 8;
        ldmatrix_x4_trans(b0, b1, b2, b3, sV + key * ROWB + col * 2);
        mma_bf16_16816(oacc[np * 2], pf[ks], b0, b1);
        mma_bf16_16816(oacc[np * 2 + 1], pf[ks], b2, b3);
      }
    }
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      rs[r] += __shfl_xor_sync(0xffffffffu, rs[r], 1);
      rs[r] += __shfl_xor_sync(0xffffffffu, rs[r], 2);
    }
    const float inv0 = rs[0] > 0.0f ? 1.0f / rs[0] : 0.0f;
    const float inv1 = rs[1] > 0.0f ? 1.0f / rs[1] : 0.0f;
    // O through this warp's own 16 rows of the consumed Q buffer (its fragments are in registers), then out in rows
    __syncwarp();
    const int lr = warp * 16 + (lane >> 2);
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {
      const int col = nt * 8 + (lane & 3) * 2;
      *reinterpret_cast<uint32_t*>(qb + lr * ROWB + col * 2) = pack_bf16(oacc[nt][0] * inv0, oacc[nt][1] * inv0);
      *reinterpret_cast<uint32_t*>(qb + (lr + 8) * ROWB + col * 2) = pack_bf16(oacc[nt][2] * inv1, oacc[nt][3] * inv1);
    }
    __syncwarp();
    const int ch = d / 8;  // 16-byte chunks per output row
    for (int i = lane; i < 16 * ch; i += 32) {
      const int r = i / ch, c = i - r * ch;
      const int row = q0 + warp * 16 + r;
      if (row < tq) *reinterpret_cast<uint4*>(og + (long long)row * ldo + c * 8) = *reinterpret_cast<const uint4*>(qb + (warp * 16 + r) * ROWB + c * 16);
    }
    __syncthreads();  // every warp is done with sQ[buf] before the prefetch of tile qt + 2 lands in it
  }
}

template <int DP, int NK16>
int launch_xattn_n(const void* q, int ldq, const void* k, int ldk, const void* v, int ldv, void* o, int ldo, int batch, int heads, int tq, int tkv,
                   int d, float scale, cudaStream_t stream) {
  constexpr int ROWB = DP * 2 + 16;
  constexpr int SMEM = (2 * 64 + 2 * NK16 * 16) * ROWB;
  static bool configured = false;
  if (!configured) {
    SASPA_CUDA(cudaFuncSetAttribute((xattn_resident_kernel<DP, NK16>), cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
    configured = true;
  }
  const int n_qt = ceil_div(tq, BQ);
  // enough CTAs for ~4 waves of the machine, at most 16 query tiles per CTA
  int per = 16;
  while (per > 1 && (long long)ceil_div(n_qt, per) * batch * heads < 4LL * saspa_num_sms()) per >>= 1;
  dim3 grid(ceil_div(n_qt, per), batch * heads);
  xattn_resident_kernel<DP, NK16><<<grid, ATT_THREADS, SMEM, stream>>>(static_cast<const __nv_bfloat16*>(q), ldq, static_cast<const __nv_bfloat16*>(k),
                                                                       ldk, static_cast<const __nv_bfloat16*>(v), ldv, static_cast<__nv_bfloat16*>(o), ldo,
                                                                       heads, tq, tkv, d, scale * 1.4426950408889634f, per);
  SASPA_LAUNCH_CHECK();
  return SASPA_OK;
}
template <int DP>
int launch_xattn(const void* q, int ldq, const void* k, int ldk, const void* v, int ldv, void* o, int ldo, int batch, int heads, int tq, int tkv,
                 int d, float scale, cudaStream_t stream) {
  if (tkv <= 80) return launch_xattn_n<DP, 5>(q, ldq, k, ldk, v, ldv, o, ldo, batch, heads, tq, tkv, d, scale, stream);
  return launch_xattn_n<DP, 8>(q, ldq, k, ldk, v, ldv, o, ldo, batch, heads, tq, tkv, d, scale, stream);
}

// ---- row softmax (in place capable), one warp per row, fp32 math ----
__global__ void __launch_bounds__(256) softmax_rows_kernel(const __nv_bfloat16* __restrict__ x, int ldx, __nv_bfloat16* __restrict__ y, int ldy,
                                                           long long rows, int cols, float scale) {
  const int lane = threadIdx.x & 31;
  const int wpb = blockDim.x >> 5;
  for (long long row = (long long)blockIdx.x * wpb + (threadIdx.x >> 5); row < rows; row += (long long)gridDim.x * wpb) {
    const __nv_bfloat16* xr = x + row * ldx;
    __nv_bfloat16* yr = y + row * ldy;
    float mx = -INFINITY;
    for (int c = lane * 8; c < cols; c += 256) {
      uint4 u = *reinterpret_cast<const uint4*>(xr + c);
      mx = fmaxf(mx, fmaxf(fmaxf(fmaxf(bf16_lo(u.x), bf16_hi(u.x)), fmaxf(bf16_lo(u.y), bf16_hi(u.y))),
                           fmaxf(fmaxf(bf16_lo(u.z), bf16_hi(u.z)), fmaxf(bf16_lo(u.w), bf16_hi(u.w)))));
    }
    mx = warp_max(mx) * scale;
    float sum = 0.0f;
    for (int c = lane * 8; c < cols; c += 256) {
      uint4 u = *reinterpret_cast<const uint4*>(xr + c);
      sum += __expf(bf16_lo(u.x) * scale - mx) + __expf(bf16_hi(u.x) * scale - mx) + __expf(bf16_lo(u.y) * scale - mx) +
             __expf(bf16_hi(u.y) * scale - mx) + __expf(bf16_lo(u.z) * scale - mx) + __expf(bf16_hi(u.z) * scale - mx) +
             __expf(bf16_lo(u.w) * scale - mx) + __expf(bf16_hi(u.w) * scale - mx);
    }
    const float inv = 1.0f / warp_sum(sum);
    for (int c = lane * 8; c < cols; c += 256) {
      uint4 u = *reinterpret_cast<const uint4*>(xr + c);
      uint4 w;
      w.x = pack_bf16(__expf(bf16_lo(u.x) * scale - mx) * inv, __expf(bf16_hi(u.x) * scale - mx) * inv);
      w.y = pack_bf16(__expf(bf16_lo(u.y) * scale - mx) * inv, __expf(bf16_hi(u.y) * scale - mx) * inv);
      w.z = pack_bf16(__expf(bf16_lo(u.z) * scale - mx) * inv, __expf(bf16_hi(u.z) * scale - mx) * inv);
      w.w = pack_bf16(__expf(bf16_lo(u.w) * scale - mx) * inv, __expf(bf16_hi(u.w) * scale - mx) * inv);
      *reinterpret_cast<uint4*>(yr + c) = w;
    }
  }
}

// ---- batched 2-D transpose: x [batch, rows, cols] (row stride ldx) -> y [batch, cols, rows] (row stride ldy) ----
__global__ void transpose_kernel(const __nv_bfloat16* __restrict__ x, long long ldx, long long bsx, __nv_bfloat16* __restrict__ y, long long ldy,
                                 long long bsy, int rows, int cols) {
  __shared__ __nv_bfloat16 tile[32][33];
  const int b = blockIdx.z;
  const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    int r = r0 + i, c = c0 + threadIdx.x;
    if (r < rows && c < cols) tile[i][threadIdx.x] = x[b * bsx + (long long)r * ldx + c];
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    int c = c0 + i, r = r0 + threadIdx.x;
    if (r < rows && c < cols) y[b * bsy + (long long)c * ldy + r] = tile[threadIdx.x][i];
  }
}

std::atomic<int> g_attention_impl{0};  // 0 = auto (K/V-resident kernel for tkv <= 128, else tcgen05 when the shape has an instantiation), 1 = mma.sync flash kernel, 2 = tcgen05 only

}  // namespace

int saspa_attention_tc(const void* q, int ldq, const void* k, int ldk, const void* v, int ldv, void* o, int ldo, int batch, int heads, int tq,
                       int tkv, int d, float scale, int causal, cudaStream_t stream);
int saspa_xattention_tc(const void* q, int ldq, const void* k, int ldk, const void* v, int ldv, void* o, int ldo, int batch, int heads, int tq,
                        int tkv, int d, float scale, cudaStream_t stream);

extern "C" int saspa_attention_impl(int impl) {
  const int prev = g_attention_impl.load();
  if (impl >= 0 && impl <= 3) g_attention_impl = impl;
  return prev;
}

extern "C" int saspa_attention_bf16(const void* q, int ldq, const void* k, int ldk, const void* v, int ldv, void* o, int ldo, int batch,
                                    int heads, int tq, int tkv, int d, float scale, int causal, cudaStream_t stream) {
  SASPA_CHECK_ARG(batch >= 0 && heads > 0 && tq >= 0 && tkv > 0 && d > 0, "saspa_attention_bf16: bad shape");
  SASPA_CHECK_ARG(d % 8 == 0 && d <= 160, "saspa_attention_bf16: head_dim must be a multiple of 8 and <= 160 (got %d); d=512 runs as GEMM-softmax-GEMM", d);
  SASPA_CHECK_ARG(ldq % 8 == 0 && ldk % 8 == 0 && ldv % 8 == 0 && ldo % 2 == 0, "saspa_attention_bf16: row strides must be multiples of 8");
  if (batch == 0 || tq == 0) return SASPA_OK;
  SASPA_CHECK_ARG(q && k && v && o, "saspa_attention_bf16: null pointer");
  SASPA_CHECK_ARG(((reinterpret_cast<uintptr_t>(q) | reinterpret_cast<uintptr_t>(k) | reinterpret_cast<uintptr_t>(v)) & 15) == 0 &&
                      (reinterpret_cast<uintptr_t>(o) & 3) == 0,
                  "saspa_attention_bf16: q/k/v must be 16-byte aligned");
  SASPA_CHECK_ARG((long long)batch * heads <= 65535, "saspa_attention_bf16: batch*heads must be <= 65535");
  // short key sequences (cross-attention over the text tokens): persistent tcgen05 kernel with K / V resident (xattention_tc.cu);
  // impl 3 forces the older mma.sync K/V-resident kernel for A/B timing
  if ((g_attention_impl == 0 || g_attention_impl == 2) && !causal && tkv <= 128 && tq >= 256) {
    const int rc = saspa_xattention_tc(q, ldq, k, ldk, v, ldv, o, ldo, batch, heads, tq, tkv, d, scale, stream);
    if (rc != SASPA_ERR_UNSUPPORTED) return rc;
  }
  if ((g_attention_impl == 0 || g_attention_impl == 3) && !causal && scale > 0.0f && tkv <= 128 && tq >= 256 && (ldo % 8) == 0 && (reinterpret_cast<uintptr_t>(o) & 15) == 0) {
    if (d <= 48) return launch_xattn<48>(q, ldq, k, ldk, v, ldv, o, ldo, batch, heads, tq, tkv, d, scale, stream);
    if (d <= 64) return launch_xattn<64>(q, ldq, k, ldk, v, ldv, o, ldo, batch, heads, tq, tkv, d, scale, stream);
    if (d <= 80) return launch_xattn<80>(q, ldq, k, ldk, v, ldv, o, ldo, batch, heads, tq, tkv, d, scale, stream);
    if (d <= 128) return launch_xattn<128>(q, ldq, k, ldk, v, ldv, o, ldo, batch, heads, tq, tkv, d, scale, stream);
    // d = 160 (the 16 x 16 level): one resident CTA per SM loses to the tcgen05 kernel (b32 h8 256x77: 45 vs 31 us), which is
    // tried below; the resident kernel remains the fallback for head dims without a tcgen05 instantiation
    if (d != 160) return launch_xattn<160>(q, ldq, k, ldk, v, ldv, o, ldo, batch, heads, tq, tkv, d, scale, stream);
  }
  if (g_attention_impl != 1) {
    const int rc = saspa_attention_tc(q, ldq, k, ldk, v, ldv, o, ldo, batch, heads, tq, tkv, d, scale, causal, stream);
    if (rc != SASPA_ERR_UNSUPPORTED) return rc;
    SASPA_CHECK_ARG(g_attention_impl != 2, "saspa_attention_bf16: no tcgen05 instantiation for head_dim %d", d);
  }
  if (d <= 48) return launch_flash<48>(q, ldq, k, ldk, v, ldv, o, ldo, batch, heads, tq, tkv, d, scale, causal, stream);
  if (d <= 64) return launch_flash<64>(q, ldq, k, ldk, v, ldv, o, ldo, batch, heads, tq, tkv, d, scale, causal, stream);
  if (d <= 80) return launch_flash<80>(q, ldq, k, ldk, v, ldv, o, ldo, batch, heads, tq, tkv, d, scale, causal, stream);
  if (d <= 128) return launch_flash<128>(q, ldq, k, ldk, v, ldv, o, ldo, batch, heads, tq, tkv, d, scale, causal, stream);
  return launch_flash<160>(q, ldq, k, ldk, v, ldv, o, ldo, batch, heads, tq, tkv, d, scale, causal, stream);
}

extern "C" int saspa_softmax_rows_bf16(const void* x, int ldx, void* y, int ldy, long long rows, int cols, float scale, cudaStream_t stream) {
  SASPA_CHECK_ARG(rows >= 0 && cols > 0 && cols % 8 == 0 && ldx % 8 == 0 && ldy % 8 == 0, "saspa_softmax_rows_bf16: cols and strides must be multiples of 8");
  if (rows == 0) return SASPA_OK;
  SASPA_CHECK_ARG(x && y, "saspa_softmax_rows_bf16: null pointer");
  long long g = ceil_div_ll(rows, 8);
  long long cap = (long long)saspa_num_sms() * 8;
  softmax_rows_kernel<<<(int)(g < cap ? g : cap), 256, 0, stream>>>(static_cast<const __nv_bfloat16*>(x), ldx, static_cast<__nv_bfloat16*>(y), ldy, rows, cols, scale);
  SASPA_LAUNCH_CHECK();
  return SASPA_OK;
}

extern "C" int saspa_transpose_bf16(const void* x, int ldx, long long batch_stride_x, void* y, int ldy, long long batch_stride_y, int batch,
                                    int rows, int cols, cudaStream_t stream) {
  SASPA_CHECK_ARG(batch >= 0 && rows >= 0 && cols >= 0, "saspa_transpose_bf16: bad shape");
  if (batch == 0 || rows == 0 || cols == 0) return SASPA_OK;
  SASPA_CHECK_ARG(x && y && batch <= 65535, "saspa_transpose_bf16: bad arguments");
  dim3 grid(ceil_div(cols, 32), ceil_div(rows, 32), batch);
  transpose_kernel<<<grid, dim3(32, 8), 0, stream>>>(static_cast<const __nv_bfloat16*>(x), ldx, batch_stride_x, static_cast<__nv_bfloat16*>(y), ldy,
                                                     batch_stride_y, rows, cols);
  SASPA_LAUNCH_CHECK();
  return SASPA_OK;
}
