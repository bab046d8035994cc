// HBM-bound glue kernels of the denoising path (NHWC bf16 activations): im2col for the few strided /
// odd-channel convs, GroupNorm(+SiLU), LayerNorm, activation, add, nearest x2 upsample, pooling,
// latent layout casts, sinusoidal timestep projection, CFG + scheduler update, VAE u8 quantise.
// All are coalesced along the channel dimension with 16-byte vectors where alignment allows,
// warp-shuffle reductions, grids sized from the problem (grid-stride loops capped at 148 x 16 CTAs).
//
// Reference arithmetic replaced: diffusers 0.32.2 modules run by pipe(**pipe_args)
// (run_aug/run_aug.py:278): GroupNorm/LayerNorm/SiLU/Upsample2D in models/resnet.py, attention.py,
// embeddings.py (Timesteps), schedulers/scheduling_{ddim,unipc_multistep,pndm}.py step(),
// image_processor.py postprocess.
#include "common.cuh"
#include <atomic>

#include "../../include/saspa_b200.h"
#include "tuning_hooks.h"

namespace {

inline int grid_for(long long work_items, int block, int cap_mult = 16) {
  long long g = ceil_div_ll(work_items, block);
  long long cap = (long long)saspa_num_sms() * cap_mult;
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return (int)g;
}

__device__ __forceinline__ float act_f(float v, int act) {
  switch (act) {
    case SASPA_ACT_SILU: return silu_f(v);
    case SASPA_ACT_GELU: return gelu_erf_f(v);
    case SASPA_ACT_RELU: return fmaxf(v, 0.0f);
    case SASPA_ACT_QUICKGELU: return quick_gelu_f(v);
    default: return v;
  }
}

__device__ __forceinline__ void unpack8(const uint4& u, float* f) {
  f[0] = bf16_lo(u.x); f[1] = bf16_hi(u.x); f[2] = bf16_lo(u.y); f[3] = bf16_hi(u.y);
  f[4] = bf16_lo(u.z); f[5] = bf16_hi(u.z); f[6] = bf16_lo(u.w); f[7] = bf16_hi(u.w);
}
__device__ __forceinline__ uint4 pack8(const float* f) {
  uint4 u;
  u.x = pack_bf16(f[0], f[1]); u.y = pack_bf16(f[2], f[3]); u.z = pack_bf16(f[4], f[5]); u.w = pack_bf16(f[6], f[7]);
  return u;
}

// ------------------------------------------------------------------------------------------------
// im2col
// ------------------------------------------------------------------------------------------------
template <bool VEC>
__global__ void im2col_kernel(const __nv_bfloat16* __restrict__ x, int ldx, int n, int h, int w, int c, int kh, int kw, int stride,
                              int pt, int pl, int oh, int ow, __nv_bfloat16* __restrict__ cols, int kpad) {
  constexpr int V = VEC ? 8 : 1;
  const int kv = kpad / V;
  const long long total = (long long)n * oh * ow * kv;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    int kq = (int)(i % kv);
    long long row = i / kv;
    int ox = (int)(row % ow);
    long long t = row / ow;
    int oy = (int)(t % oh);
    int img = (int)(t / oh);
    int k = kq * V;
    __nv_bfloat16* dst = cols + row * kpad + k;
    bool zero = true;
    const __nv_bfloat16* src = nullptr;
    if (k < kh * kw * c) {
      int tap = k / c, ch = k - tap * c;
      int ky = tap / kw, kx = tap - ky * kw;
      int iy = oy * stride - pt + ky, ix = ox * stride - pl + kx;
      if (iy >= 0 && iy < h && ix >= 0 && ix < w) {
        zero = false;
        src = x + (((long long)img * h + iy) * w + ix) * ldx + ch;
      }
    }
    if (VEC) {
      uint4 v = zero ? make_uint4(0, 0, 0, 0) : __ldg(reinterpret_cast<const uint4*>(src));
      *reinterpret_cast<uint4*>(dst) = v;
    } else {
      *dst = zero ? __float2bfloat16(0.0f) : *src;
    }
  }
}

// ------------------------------------------------------------------------------------------------
// GroupNorm (NHWC), one launch, x read from HBM once, deterministic (batch-composition invariant).
// Each image is split over `ctas_per_img` (<= 32, a function of hw only) CTAs.  Phase 1: per-CTA (sum, sum of squares) per
// group over its pixel range -> workspace (fixed-order reductions only, no float atomics).  A per-image
// arrive/spin counter orders phase 1 before phase 2 across the image's CTAs.  Phase 2: every CTA folds the
// partials in the same order (double), stages per-channel scale/shift in smem and normalises its own pixel
// range again (the re-read is served by L2), optional SiLU, 16-byte vectors.
// ------------------------------------------------------------------------------------------------
constexpr int GN_THREADS = 512;

__global__ void __launch_bounds__(GN_THREADS) gn_fused_kernel(const __nv_bfloat16* __restrict__ x, int ldx, int hw, int c, int groups, float eps,
                                                              const float* __restrict__ gamma, const float* __restrict__ beta, int act,
                                                              __nv_bfloat16* __restrict__ y, int ldy, int ctas_per_img, int pix_per_cta,
                                                              float2* __restrict__ partials, unsigned int* __restrict__ counters) {
  extern __shared__ float gn_smem[];  // phase 1: [pix_par][c] sums + [pix_par][c] squares; phase 2: scale[c], shift[c]
  __shared__ float s_stat[2 * 64];    // per-group mean, rstd (groups <= 64)
  const int img = blockIdx.x / ctas_per_img, part = blockIdx.x % ctas_per_img;
  const int p0 = part * pix_per_cta;
  const int p1 = min(p0 + pix_per_cta, hw);
  const int cv = c / 8;  // 16-byte vectors per pixel
  const int tpp = cv;    // threads per pixel (cv <= GN_THREADS)
  const int pix_par = GN_THREADS / tpp;
  const int my_v = threadIdx.x % tpp, my_p = threadIdx.x / tpp;
  const __nv_bfloat16* xi = x + (size_t)img * hw * ldx;
  float* s_sum = gn_smem;
  float* s_sq = gn_smem + (size_t)pix_par * c;

  // ---- phase 1: partial statistics of my pixel range ----
  if (my_p < pix_par) {
    float s[8], ss[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) s[j] = ss[j] = 0.0f;
    const __nv_bfloat16* base = xi + my_v * 8;
    int p = p0 + my_p;
    for (; p + 3 * pix_par < p1; p += 4 * pix_par) {  // four loads in flight
      uint4 u[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) u[k] = __ldg(reinterpret_cast<const uint4*>(base + (size_t)(p + k * pix_par) * ldx));
#pragma unroll
      for (int k = 0; k < 4; k += 2) {
        float f[8], g[8];
        unpack8(u[k], f);
        unpack8(u[k + 1], g);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          s[j] += f[j] + g[j];
          ss[j] += f[j] * f[j] + g[j] * g[j];
        }
      }
    }
    for (; p < p1; p += pix_par) {
      uint4 u = __ldg(reinterpret_cast<const uint4*>(base + (size_t)p * ldx));
      float f[8];
      unpack8(u, f);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        s[j] += f[j];
        ss[j] += f[j] * f[j];
      }
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      s_sum[my_p * c + my_v * 8 + j] = s[j];
      s_sq[my_p * c + my_v * 8 + j] = ss[j];
    }
  }
  __syncthreads();
  {
    // one warp per group, lanes stride over the (pixel lane, channel) partials, shuffle tree: fixed order
    const int cg = c / groups;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int g = warp; g < groups; g += GN_THREADS / 32) {
      float a = 0.0f, b = 0.0f;
      for (int i = lane; i < pix_par * cg; i += 32) {
        const int pp = i / cg, ch = g * cg + (i - pp * cg);
        a += s_sum[pp * c + ch];
        b += s_sq[pp * c + ch];
      }
      a = warp_sum(a);
      b = warp_sum(b);
      if (lane == 0) partials[((size_t)img * ctas_per_img + part) * groups + g] = make_float2(a, b);
    }
  }
  // ---- all CTAs of this image have published their partials ----
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    atomicAdd(&counters[img], 1u);
    unsigned int seen;
    do {
      asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(counters + img) : "memory");
      if (seen < (unsigned int)ctas_per_img) __nanosleep(32);
    } while (seen < (unsigned int)ctas_per_img);
  }
  __syncthreads();

  // ---- phase 2: fold partials (same order in every CTA), per-channel scale / shift, apply ----
  if (threadIdx.x < groups) {
    const int g = threadIdx.x;
    double a = 0.0, b = 0.0;
    for (int q = 0; q < ctas_per_img; ++q) {
      float2 v = __ldcg(&partials[((size_t)img * ctas_per_img + q) * groups + g]);
      a += (double)v.x;
      b += (double)v.y;
    }
    const double cnt = (double)hw * (c / groups);
    const double mean = a / cnt;
    double var = b / cnt - mean * mean;
    if (var < 0.0) var = 0.0;
    s_stat[g] = (float)mean;
    s_stat[64 + g] = (float)(1.0 / sqrt(var + (double)eps));
  }
  __syncthreads();
  float* s_scale = gn_smem;
  float* s_shift = gn_smem + c;
  for (int ch = threadIdx.x; ch < c; ch += GN_THREADS) {
    const int g = ch / (c / groups);
    const float ga = gamma ? gamma[ch] : 1.0f, be = beta ? beta[ch] : 0.0f;
    const float sc = s_stat[64 + g] * ga;
    s_scale[ch] = sc;
    s_shift[ch] = be - s_stat[g] * sc;
  }
  __syncthreads();
  if (my_p < pix_par) {
    float sc[8], sh[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      sc[j] = s_scale[my_v * 8 + j];
      sh[j] = s_shift[my_v * 8 + j];
    }
    const __nv_bfloat16* base = xi + my_v * 8;
    __nv_bfloat16* yb = y + (size_t)img * hw * ldy + my_v * 8;
    int p = p0 + my_p;
    for (; p + 3 * pix_par < p1; p += 4 * pix_par) {  // four loads in flight
      uint4 u[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) u[k] = __ldcg(reinterpret_cast<const uint4*>(base + (size_t)(p + k * pix_par) * ldx));
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        float f[8];
        unpack8(u[k], f);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float t = fmaf(f[j], sc[j], sh[j]);
          f[j] = (act == SASPA_ACT_SILU) ? silu_f(t) : t;
        }
        *reinterpret_cast<uint4*>(yb + (size_t)(p + k * pix_par) * ldy) = pack8(f);
      }
    }
    for (; p < p1; p += pix_par) {
      uint4 u = __ldcg(reinterpret_cast<const uint4*>(base + (size_t)p * ldx));
      float f[8];
      unpack8(u, f);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float t = fmaf(f[j], sc[j], sh[j]);
        f[j] = (act == SASPA_ACT_SILU) ? silu_f(t) : t;
      }
      *reinterpret_cast<uint4*>(yb + (size_t)p * ldy) = pack8(f);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// GroupNorm, register-resident variant (every UNet / ControlNet level: hw <= ~8K pixels per image).
// A CTA owns (image, channel slice, pixel part): a slice is a whole number of groups (and of 16-byte vectors), a part
// is pix_par x GN_VPT pixels.  Every thread issues all of its <= GN_VPT 16-byte loads up front (memory-level
// parallelism 16) and keeps them in registers: x is read from HBM exactly once, nothing is re-read.  Statistics:
// per-thread per-channel sums -> shared memory -> one warp per group (fixed order) -> if the image slice spans several
// parts, partials go through the workspace and the parts of ONE (image, slice) wait for each other on an arrive
// counter (they are adjacent in dispatch order, so all of them are resident) -> every CTA folds the partials in the
// same order (double).  Deterministic and independent of what else shares the batch.  Then scale / shift (+SiLU) is
// applied to the registers and written out.
// ------------------------------------------------------------------------------------------------
constexpr int GN_VPT = 16;

__global__ void __launch_bounds__(GN_THREADS, 1)
    gn_reg_kernel(const __nv_bfloat16* __restrict__ x, int ldx, int hw, int c, int groups, float eps, const float* __restrict__ gamma,
                  const float* __restrict__ beta, int act, __nv_bfloat16* __restrict__ y, int ldy, int c_slice, int slices, int parts,
                  int pix_per_cta, float2* __restrict__ partials, unsigned int* __restrict__ counters) {
  extern __shared__ float gn_smem[];  // [pix_par][c_slice] sums + [pix_par][c_slice] squares; later scale[c_slice], shift[c_slice]
  __shared__ float s_stat[2 * 64];
  const int part = blockIdx.x % parts;
  const int slice = (blockIdx.x / parts) % slices;
  const int img = blockIdx.x / (parts * slices);
  const int cg = c / groups, gps = c_slice / cg;  // channels per group, groups per slice
  const int tpp = c_slice / 8, pix_par = GN_THREADS / tpp;
  const int my_v = threadIdx.x % tpp, my_p = threadIdx.x / tpp;
  const bool active = my_p < pix_par;
  const int p0 = part * pix_per_cta;
  const size_t col = (size_t)slice * c_slice + my_v * 8;
  const __nv_bfloat16* xb = x + (size_t)img * hw * ldx + col;

  uint4 v[GN_VPT];
#pragma unroll
  for (int k = 0; k < GN_VPT; ++k) {
    const int p = p0 + my_p + k * pix_par;
    v[k] = (active && p < hw && k * pix_par < pix_per_cta) ? __ldg(reinterpret_cast<const uint4*>(xb + (size_t)p * ldx)) : make_uint4(0, 0, 0, 0);
  }
  float* s_sum = gn_smem;
  float* s_sq = gn_smem + (size_t)pix_par * c_slice;
  if (active) {
    float s[8], ss[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) s[j] = ss[j] = 0.0f;
#pragma unroll
    for (int k = 0; k < GN_VPT; ++k) {
      float f[8];
      unpack8(v[k], f);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        s[j] += f[j];
        ss[j] = fmaf(f[j], f[j], ss[j]);
      }
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      s_sum[my_p * c_slice + my_v * 8 + j] = s[j];
      s_sq[my_p * c_slice + my_v * 8 + j] = ss[j];
    }
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const double cnt = (double)hw * cg;
  for (int g = warp; g < gps; g += GN_THREADS / 32) {
    float a = 0.0f, b = 0.0f;
    for (int i = lane; i < pix_par * cg; i += 32) {
      const int pp = i / cg, ch = g * cg + (i - pp * cg);
      a += s_sum[pp * c_slice + ch];
      b += s_sq[pp * c_slice + ch];
    }
    a = warp_sum(a);
    b = warp_sum(b);
    if (lane == 0) {
      if (parts == 1) {
        const double mean = (double)a / cnt;
        double var = (double)b / cnt - mean * mean;
        if (var < 0.0) var = 0.0;
        s_stat[g] = (float)mean;
        s_stat[64 + g] = (float)(1.0 / sqrt(var + (double)eps));
      } else {
        partials[(((size_t)img * slices + slice) * parts + part) * gps + g] = make_float2(a, b);
      }
    }
  }
  if (parts > 1) {
    __syncthreads();
    if (threadIdx.x == 0) {
      __threadfence();
      unsigned int* ctr = counters + (size_t)img * slices + slice;
      atomicAdd(ctr, 1u);
      unsigned int seen;
      do {
        asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(ctr) : "memory");
        if (seen < (unsigned int)parts) __nanosleep(32);
      } while (seen < (unsigned int)parts);
    }
    __syncthreads();
    if (threadIdx.x < gps) {
      const int g = threadIdx.x;
      double a = 0.0, b = 0.0;
      for (int q = 0; q < parts; ++q) {
        const float2 t = __ldcg(&partials[(((size_t)img * slices + slice) * parts + q) * gps + g]);
        a += (double)t.x;
        b += (double)t.y;
      }
      const double mean = a / cnt;
      double var = b / cnt - mean * mean;
      if (var < 0.0) var = 0.0;
      s_stat[g] = (float)mean;
      s_stat[64 + g] = (float)(1.0 / sqrt(var + (double)eps));
    }
  }
  __syncthreads();
  float* s_scale = gn_smem;
  float* s_shift = gn_smem + c_slice;
  for (int ch = threadIdx.x; ch < c_slice; ch += GN_THREADS) {
    const int g = ch / cg, gc = slice * c_slice + ch;
    const float ga = gamma ? gamma[gc] : 1.0f, be = beta ? beta[gc] : 0.0f;
    const float sc = s_stat[64 + g] * ga;
    s_scale[ch] = sc;
    s_shift[ch] = be - s_stat[g] * sc;
  }
  __syncthreads();
  if (active) {
    float sc[8], sh[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      sc[j] = s_scale[my_v * 8 + j];
      sh[j] = s_shift[my_v * 8 + j];
    }
    __nv_bfloat16* yb = y + (size_t)img * hw * ldy + col;
#pragma unroll
    for (int k = 0; k < GN_VPT; ++k) {
      const int p = p0 + my_p + k * pix_par;
      if (p < hw && k * pix_par < pix_per_cta) {
        float f[8];
        unpack8(v[k], f);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float t = fmaf(f[j], sc[j], sh[j]);
          f[j] = (act == SASPA_ACT_SILU) ? silu_f(t) : t;
        }
        *reinterpret_cast<uint4*>(yb + (size_t)p * ldy) = pack8(f);
      }
    }
  }
}

struct GnRegPlan {
  bool ok;
  int c_slice, slices, parts, pix_per_cta;
};
// Slice = a multiple of lcm(channels per group, 8) that divides c, as close to 320 channels as possible (<= 512).
GnRegPlan gn_reg_plan(int hw, int c, int groups) {
  GnRegPlan r = {false, 0, 0, 0, 0};
  const int cg = c / groups;
  int a = cg, b = 8;
  while (b) {
    int t = a % b;
    a = b;
    b = t;
  }
  const int unit = cg / a * 8;  // lcm(cg, 8)
  int best = 0;
  for (int m = unit; m <= 512 && m <= c; m += unit) {
    if (c % m != 0) continue;
    if (best == 0 || abs(m - 320) < abs(best - 320)) best = m;
  }
  if (best == 0 || best / cg > 64) return r;
  r.c_slice = best;
  r.slices = c / best;
  const int pix_par = GN_THREADS / (best / 8);
  r.pix_per_cta = pix_par * GN_VPT;
  r.parts = (hw + r.pix_per_cta - 1) / r.pix_per_cta;
  // the parts of one (image, slice) wait for each other: keep them well inside one wave.  Measured (profiles/
  // r1_hbm_microbench_i.txt): one 16-warp CTA per SM wins up to 16 x 16 pixels per image (1.2-2x at 16x16 / 8x8, where the
  // two-pass kernel runs one CTA per image); from 32 x 32 on the two-pass kernel's four CTAs per SM overlap better
  // (32x1024x640: 43 vs 53 us; 64x1024x640: 67 vs 92 us).  An 8-vector / two-CTA-per-SM variant measured no better.
  r.ok = r.parts <= 64 && hw <= 512;
  return r;
}

// workspace layout: [n] u32 arrive counters (padded to 256 B) | [n][ctas_per_img][groups] float2 partials
struct GnPlan {
  int ctas_per_img, pix_per_cta;
  size_t counters_bytes, total_bytes;
};
GnPlan gn_plan(int n, int hw, int groups) {
  GnPlan p;
  // The split of an image over CTAs depends on hw only -- never on n -- so an image's statistics are reduced in the
  // same order whatever else shares the batch (bit-exact batch invariance).  Only the CTAs of ONE image wait for each
  // other (<= 32 of them), so in-order block dispatch always makes progress.
  int per = hw / 256;
  if (per < 1) per = 1;
  if (per > 32) per = 32;
  p.pix_per_cta = ceil_div(hw, per);
  p.ctas_per_img = ceil_div(hw, p.pix_per_cta);
  // sized for either kernel: counters per (image, slice) with slices <= groups; partials per (image, part, group)
  // with parts <= max(ctas_per_img, 64) (register-resident variant)
  p.counters_bytes = ((size_t)n * groups * 4 + 255) / 256 * 256;
  const int max_parts = p.ctas_per_img > 64 ? p.ctas_per_img : 64;
  p.total_bytes = p.counters_bytes + (size_t)n * max_parts * groups * sizeof(float2);
  return p;
}

// ------------------------------------------------------------------------------------------------
// LayerNorm: one warp per row, row kept in registers (c <= 2048), two-pass variance.
// ------------------------------------------------------------------------------------------------
constexpr int LN_MAXV = 8;  // 16-byte vectors per lane
__global__ void __launch_bounds__(256) layernorm_kernel(const __nv_bfloat16* __restrict__ x, int ldx, int rows, int c, float eps,
                                                        const float* __restrict__ gamma, const float* __restrict__ beta,
                                                        __nv_bfloat16* __restrict__ y, int ldy) {
  const int warps_per_block = blockDim.x >> 5;
  const int lane = threadIdx.x & 31;
  const int cv = c / 8;
  for (int row = blockIdx.x * warps_per_block + (threadIdx.x >> 5); row < rows; row += gridDim.x * warps_per_block) {
    const __nv_bfloat16* xr = x + (size_t)row * ldx;
    float f[LN_MAXV][8];
    float sum = 0.0f;
#pragma unroll
    for (int i = 0; i < LN_MAXV; ++i) {
      int v = lane + i * 32;
      if (v < cv) {
        uint4 u = __ldg(reinterpret_cast<const uint4*>(xr + v * 8));
        unpack8(u, f[i]);
#pragma unroll
        for (int j = 0; j < 8; ++j) sum += f[i][j];
      }
    }
    const float mean = warp_sum(sum) / (float)c;
    float sq = 0.0f;
#pragma unroll
    for (int i = 0; i < LN_MAXV; ++i) {
      int v = lane + i * 32;
      if (v < cv) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          float d = f[i][j] - mean;
          sq += d * d;
        }
      }
    }
    const float rstd = rsqrtf(warp_sum(sq) / (float)c + eps);
    __nv_bfloat16* yr = y + (size_t)row * ldy;
#pragma unroll
    for (int i = 0; i < LN_MAXV; ++i) {
      int v = lane + i * 32;
      if (v < cv) {
        float o[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          float ga = gamma ? __ldg(gamma + v * 8 + j) : 1.0f, be = beta ? __ldg(beta + v * 8 + j) : 0.0f;
          o[j] = (f[i][j] - mean) * rstd * ga + be;
        }
        *reinterpret_cast<uint4*>(yr + v * 8) = pack8(o);
      }
    }
  }
}

// LayerNorm for the transformer widths of the denoising path (c = 8 * LPR * VPL): LPR lanes per row, 32 / LPR rows
// per warp, VPL 16-byte vectors per lane (row stays in registers), shuffle reductions over LPR lanes, two-pass variance.
template <int LPR, int VPL>
__global__ void __launch_bounds__(256, 3) layernorm_sub_kernel(const __nv_bfloat16* __restrict__ x, int ldx, int rows, float eps,
                                                               const float* __restrict__ gamma, const float* __restrict__ beta,
                                                               __nv_bfloat16* __restrict__ y, int ldy) {
  constexpr int RPW = 32 / LPR;  // rows per warp
  constexpr int C = 8 * LPR * VPL;
  // gamma / beta live in shared memory (not 2 x VPL x 8 registers): the kernel is HBM-bound and needs the occupancy -- three
  // 256-thread CTAs per SM, each thread with its NEXT row's 16-byte loads already in flight while it reduces the current one.
  __shared__ __align__(16) float s_ga[C];
  __shared__ __align__(16) float s_be[C];
  for (int ch = threadIdx.x; ch < C; ch += blockDim.x) {
    s_ga[ch] = gamma ? __ldg(gamma + ch) : 1.0f;
    s_be[ch] = beta ? __ldg(beta + ch) : 0.0f;
  }
  __syncthreads();
  const int lane = threadIdx.x & 31, sub = lane / LPR, l = lane % LPR;
  const long long warp_global = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const long long warps_total = (long long)gridDim.x * (blockDim.x >> 5);
  uint4 nxt[VPL];
  {
    const long long row = warp_global * RPW + sub;
    const __nv_bfloat16* xr = x + (size_t)(row < rows ? row : 0) * ldx;
#pragma unroll
    for (int i = 0; i < VPL; ++i) nxt[i] = __ldg(reinterpret_cast<const uint4*>(xr + (l + i * LPR) * 8));
  }
  for (long long r0 = warp_global * RPW; r0 < rows; r0 += warps_total * RPW) {
    const long long row = r0 + sub;
    const bool ok = row < rows;
    uint4 cur[VPL];  // the row stays PACKED in registers (unpacking bf16 is a shift/mask; 8 x VPL floats would spill at this occupancy)
    float sum = 0.0f;
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
      cur[i] = nxt[i];
      float f[8];
      unpack8(cur[i], f);
#pragma unroll
      for (int j = 0; j < 8; ++j) sum += f[j];
    }
    {
      const long long rn = row + warps_total * RPW;  // prefetch this lane's next row
      const __nv_bfloat16* xn = x + (size_t)(rn < rows ? rn : 0) * ldx;
#pragma unroll
      for (int i = 0; i < VPL; ++i) nxt[i] = __ldg(reinterpret_cast<const uint4*>(xn + (l + i * LPR) * 8));
    }
#pragma unroll
    for (int o = LPR / 2; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    const float mean = sum * (1.0f / (float)C);
    float sq = 0.0f;
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
      float f[8];
      unpack8(cur[i], f);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float d = f[j] - mean;
        sq += d * d;
      }
    }
#pragma unroll
    for (int o = LPR / 2; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
    const float rstd = rsqrtf(sq * (1.0f / (float)C) + eps);
    if (ok) {
      __nv_bfloat16* yr = y + (size_t)row * ldy;
#pragma unroll
      for (int i = 0; i < VPL; ++i) {
        const int ch = (l + i * LPR) * 8;
        const float4 g0 = *reinterpret_cast<const float4*>(s_ga + ch), g1 = *reinterpret_cast<const float4*>(s_ga + ch + 4);
        const float4 b0 = *reinterpret_cast<const float4*>(s_be + ch), b1 = *reinterpret_cast<const float4*>(s_be + ch + 4);
        const float ga[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
        const float be[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
        float f[8], o[8];
        unpack8(cur[i], f);
#pragma unroll
        for (int j = 0; j < 8; ++j) o[j] = (f[j] - mean) * rstd * ga[j] + be[j];
        *reinterpret_cast<uint4*>(yr + (l + i * LPR) * 8) = pack8(o);
      }
    }
  }
}

template <int LPR, int VPL>
void launch_ln_sub(const void* x, int ldx, int rows, float eps, const float* gamma, const float* beta, void* y, int ldy, cudaStream_t stream) {
  constexpr int RPW = 32 / LPR;
  const int grid = grid_for(ceil_div_ll(rows, RPW), 8, 3);
  layernorm_sub_kernel<LPR, VPL><<<grid, 256, 0, stream>>>(static_cast<const __nv_bfloat16*>(x), ldx, rows, eps, gamma, beta,
                                                           static_cast<__nv_bfloat16*>(y), ldy);
}

// ------------------------------------------------------------------------------------------------
// simple elementwise
// ------------------------------------------------------------------------------------------------
__global__ void act_kernel(const __nv_bfloat16* __restrict__ x, __nv_bfloat16* __restrict__ y, size_t count, int act) {
  const size_t nv = count / 8;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < nv; i += stride) {
    uint4 u = __ldg(reinterpret_cast<const uint4*>(x) + i);
    float f[8];
    unpack8(u, f);
#pragma unroll
    for (int j = 0; j < 8; ++j) f[j] = act_f(f[j], act);
    reinterpret_cast<uint4*>(y)[i] = pack8(f);
  }
  for (size_t i = nv * 8 + (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += stride)
    y[i] = __float2bfloat16(act_f(__bfloat162float(x[i]), act));
}

__global__ void add_kernel(const __nv_bfloat16* __restrict__ a, int lda, const __nv_bfloat16* __restrict__ b, int ldb,
                           __nv_bfloat16* __restrict__ y, int ldy, int rows, int c) {
  const int cv = c / 8;
  const long long total = (long long)rows * cv;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    int v = (int)(i % cv);
    size_t r = (size_t)(i / cv);
    uint4 ua = __ldg(reinterpret_cast<const uint4*>(a + r * lda + v * 8));
    uint4 ub = __ldg(reinterpret_cast<const uint4*>(b + r * ldb + v * 8));
    float fa[8], fb[8];
    unpack8(ua, fa);
    unpack8(ub, fb);
#pragma unroll
    for (int j = 0; j < 8; ++j) fa[j] += fb[j];
    *reinterpret_cast<uint4*>(y + r * ldy + v * 8) = pack8(fa);
  }
}

__global__ void upsample2x_kernel(const uint4* __restrict__ x, int n, int h, int w, int cv, uint4* __restrict__ y) {
  const long long total = (long long)n * (2 * h) * (2 * w) * cv;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    int v = (int)(i % cv);
    long long t = i / cv;
    int ox = (int)(t % (2 * w));
    t /= (2 * w);
    int oy = (int)(t % (2 * h));
    int img = (int)(t / (2 * h));
    y[i] = __ldg(x + (((long long)img * h + (oy >> 1)) * w + (ox >> 1)) * cv + v);
  }
}

__global__ void nchw_f32_to_nhwc_bf16_kernel(const float* __restrict__ x, int n, int c, int h, int w, __nv_bfloat16* __restrict__ y,
                                             int ldy, float scale) {
  const long long total = (long long)n * h * w * c;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    int ch = (int)(i % c);
    long long pix = i / c;
    int img = (int)(pix / ((long long)h * w));
    long long sp = pix % ((long long)h * w);
    y[pix * ldy + ch] = __float2bfloat16(x[((long long)img * c + ch) * h * w + sp] * scale);
  }
}

__global__ void nhwc_to_nchw_f32_kernel(const void* __restrict__ x, int ldx, int is_fp32, int n, int c, int h, int w, float* __restrict__ y) {
  const long long total = (long long)n * c * h * w;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    long long sp = i % ((long long)h * w);
    long long t = i / ((long long)h * w);
    int ch = (int)(t % c);
    int img = (int)(t / c);
    long long src = ((long long)img * h * w + sp) * ldx + ch;
    y[i] = is_fp32 ? reinterpret_cast<const float*>(x)[src] : __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(x)[src]);
  }
}

__global__ void pool2d_kernel(const __nv_bfloat16* __restrict__ x, int n, int h, int w, int c, int k, int stride, int pad, int is_max,
                              __nv_bfloat16* __restrict__ y, int oh, int ow) {
  const int cv = c / 8;
  const long long total = (long long)n * oh * ow * cv;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    int v = (int)(i % cv);
    long long t = i / cv;
    int ox = (int)(t % ow);
    t /= ow;
    int oy = (int)(t % oh);
    int img = (int)(t / oh);
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = is_max ? -INFINITY : 0.0f;
    for (int ky = 0; ky < k; ++ky) {
      int iy = oy * stride - pad + ky;
      if (iy < 0 || iy >= h) continue;
      for (int kx = 0; kx < k; ++kx) {
        int ix = ox * stride - pad + kx;
        if (ix < 0 || ix >= w) continue;
        uint4 u = __ldg(reinterpret_cast<const uint4*>(x + (((long long)img * h + iy) * w + ix) * c + v * 8));
        float f[8];
        unpack8(u, f);
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] = is_max ? fmaxf(acc[j], f[j]) : acc[j] + f[j];
      }
    }
    if (!is_max) {
      float inv = 1.0f / (float)(k * k);  // count_include_pad semantics (pad == 0 for every avg pool on the path)
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] *= inv;
    }
    *reinterpret_cast<uint4*>(y + (((long long)img * oh + oy) * ow + ox) * c + v * 8) = pack8(acc);
  }
}

__global__ void sinusoid_kernel(const float* __restrict__ t, int rows, int dim, int flip, float freq_shift, __nv_bfloat16* __restrict__ out) {
  const int half = dim / 2;
  const int total = rows * half;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    int r = i / half, k = i % half;
    // diffusers get_timestep_embedding: exponent = -ln(10000) * k / (half - freq_shift)
    float freq = expf(-9.210340371976184f * (float)k / ((float)half - freq_shift));
    float a = t[r] * freq;
    float s = sinf(a), c = cosf(a);
    __nv_bfloat16* o = out + (size_t)r * dim;
    if (flip) {
      o[k] = __float2bfloat16(c);
      o[half + k] = __float2bfloat16(s);
    } else {
      o[k] = __float2bfloat16(s);
      o[half + k] = __float2bfloat16(c);
    }
  }
}

struct LinCombDev {
  const float* in[8];
  float* out[4];
  float coef[32];
  float pre[8];   // clip term: out[j] += post[j] * clamp(sum_i pre[i] * in[i], -clip, clip)   (clip <= 0: off)
  float post[4];
  float clip;
  int n_in, n_out;
};

__global__ void cfg_sched_kernel(const float* __restrict__ eu, const float* __restrict__ ec, float g, LinCombDev lc, size_t count) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += stride) {
    float v[8];
    float c = ec[i];
    float e = eu ? (eu[i] + g * (c - eu[i])) : c;
    v[0] = lc.in[0] ? lc.in[0][i] : 0.0f;
    v[1] = e;
#pragma unroll
    for (int k = 2; k < 8; ++k) v[k] = (k < lc.n_in && lc.in[k]) ? lc.in[k][i] : 0.0f;
    float clipped = 0.0f;
    if (lc.clip > 0.0f) {
#pragma unroll
      for (int k = 0; k < 8; ++k)
        if (k < lc.n_in) clipped += lc.pre[k] * v[k];
      clipped = fminf(fmaxf(clipped, -lc.clip), lc.clip);
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      if (j < lc.n_out) {
        float acc = lc.clip > 0.0f ? lc.post[j] * clipped : 0.0f;
#pragma unroll
        for (int k = 0; k < 8; ++k)
          if (k < lc.n_in) acc += lc.coef[j * lc.n_in + k] * v[k];
        lc.out[j][i] = acc;
      }
    }
  }
}

// img2img start: z0 = (mean + exp(0.5*clamp(logvar,-30,20)) * n_post) * scaling ; latents = a * z0 + s * n_diff
// moments fp32 NHWC [n,h,w,8] (mean 0:4, logvar 4:8); noises / latents fp32 NCHW [n,4,h,w].
__global__ void vae_sample_kernel(const float* __restrict__ moments, const float* __restrict__ n_post, const float* __restrict__ n_diff,
                                  float scaling, float a, float s, int n, int h, int w, int lc, float* __restrict__ latents,
                                  float* __restrict__ z0_out) {
  const long long total = (long long)n * lc * h * w;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    long long sp = i % ((long long)h * w);
    long long t = i / ((long long)h * w);
    int ch = (int)(t % lc);
    long long img = t / lc;
    const float* m = moments + (img * h * w + sp) * (2 * lc);
    float mean = m[ch];
    float lv = fminf(fmaxf(m[lc + ch], -30.0f), 20.0f);
    float z0 = (mean + expf(0.5f * lv) * (n_post ? n_post[i] : 0.0f)) * scaling;
    if (z0_out) z0_out[i] = z0;
    latents[i] = a * z0 + s * (n_diff ? n_diff[i] : 0.0f);
  }
}

__global__ void vae_quant_kernel(const void* __restrict__ x, int ldx, int is_fp32, size_t pixels, uint8_t* __restrict__ out) {
  const size_t total = pixels * 3;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    size_t p = i / 3;
    int ch = (int)(i % 3);
    float v = is_fp32 ? reinterpret_cast<const float*>(x)[p * ldx + ch] : __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(x)[p * ldx + ch]);
    v = fminf(fmaxf(v * 0.5f + 0.5f, 0.0f), 1.0f);
    out[i] = (uint8_t)__float2int_rn(v * 255.0f);  // round-half-even, as numpy .round()
  }
}

}  // namespace

extern "C" int saspa_im2col_bf16(const void* x, int ldx, int n, int h, int w, int c, int kh, int kw, int stride, int pad_top,
                                 int pad_left, int oh, int ow, void* cols, int kpad, cudaStream_t stream) {
  SASPA_CHECK_ARG(n >= 0 && h > 0 && w > 0 && c > 0 && kh > 0 && kw > 0 && stride > 0 && oh >= 0 && ow >= 0, "saspa_im2col_bf16: bad shape");
  SASPA_CHECK_ARG(kpad >= kh * kw * c && kpad % 8 == 0, "saspa_im2col_bf16: kpad must be >= kh*kw*c and a multiple of 8");
  if (n == 0 || oh == 0 || ow == 0) return SASPA_OK;
  SASPA_CHECK_ARG(x && cols, "saspa_im2col_bf16: null pointer");
  const bool vec = (c % 8 == 0) && (ldx % 8 == 0) && ((reinterpret_cast<uintptr_t>(x) & 15) == 0) && ((reinterpret_cast<uintptr_t>(cols) & 15) == 0);
  const long long items = (long long)n * oh * ow * (vec ? kpad / 8 : kpad);
  const int grid = grid_for(items, 256, 32);
  if (vec)
    im2col_kernel<true><<<grid, 256, 0, stream>>>(static_cast<const __nv_bfloat16*>(x), ldx, n, h, w, c, kh, kw, stride, pad_top, pad_left, oh,
                                                  ow, static_cast<__nv_bfloat16*>(cols), kpad);
  else
    im2col_kernel<false><<<grid, 256, 0, stream>>>(static_cast<const __nv_bfloat16*>(x), ldx, n, h, w, c, kh, kw, stride, pad_top, pad_left,
                                                   oh, ow, static_cast<__nv_bfloat16*>(cols), kpad);
  SASPA_LAUNCH_CHECK();
  return SASPA_OK;
}

std::atomic<int> g_gn_impl{0};  // 0 auto, 1 two-pass kernel only (tests / A-B timing)
extern "C" int saspa_groupnorm_impl(int impl) {
  const int prev = g_gn_impl.load();
  if (impl == 0 || impl == 1) g_gn_impl = impl;
  return prev;
}

extern "C" size_t saspa_groupnorm_workspace_bytes(int n, int hw, int groups) {
  if (n <= 0 || hw <= 0 || groups <= 0) return 256;
  return gn_plan(n, hw, groups).total_bytes;
}

extern "C" int saspa_groupnorm_nhwc_bf16(const void* x, int ldx, int n, int hw, int c, int groups, float eps, const float* gamma,
                                         const float* beta, int act, void* y, int ldy, void* stats_ws, size_t ws_bytes, cudaStream_t stream) {
  SASPA_CHECK_ARG(n >= 0 && hw >= 0 && c > 0 && groups > 0 && groups <= 64 && c % groups == 0, "saspa_groupnorm_nhwc_bf16: bad shape (c=%d groups=%d)", c, groups);
  SASPA_CHECK_ARG(c % 8 == 0 && ldx % 8 == 0 && ldy % 8 == 0 && c <= 8 * GN_THREADS, "saspa_groupnorm_nhwc_bf16: c, ldx, ldy must be multiples of 8 and c <= %d, got c=%d", 8 * GN_THREADS, c);
  SASPA_CHECK_ARG(act == SASPA_ACT_NONE || act == SASPA_ACT_SILU, "saspa_groupnorm_nhwc_bf16: act must be NONE or SILU");
  if (n == 0 || hw == 0) return SASPA_OK;
  SASPA_CHECK_ARG(x && y && stats_ws, "saspa_groupnorm_nhwc_bf16: null pointer");
  SASPA_CHECK_ARG((reinterpret_cast<uintptr_t>(stats_ws) & 255) == 0, "saspa_groupnorm_nhwc_bf16: stats_ws must be 256-byte aligned");
  const GnPlan p = gn_plan(n, hw, groups);
  if (ws_bytes < p.total_bytes) {
    saspa_set_error("saspa_groupnorm_nhwc_bf16: workspace too small (%zu < %zu bytes)", ws_bytes, p.total_bytes);
    return SASPA_ERR_WORKSPACE;
  }
  unsigned int* counters = reinterpret_cast<unsigned int*>(stats_ws);
  float2* partials = reinterpret_cast<float2*>(reinterpret_cast<uint8_t*>(stats_ws) + p.counters_bytes);
  const GnRegPlan rp = gn_reg_plan(hw, c, groups);
  if (rp.ok && g_gn_impl != 1) {
    if (rp.parts > 1) SASPA_CUDA(cudaMemsetAsync(counters, 0, (size_t)n * rp.slices * 4, stream));
    const int pp = GN_THREADS / (rp.c_slice / 8);
    const size_t sm = sizeof(float) * 2 * (size_t)pp * rp.c_slice;
    static size_t sm_configured = 0;
    if (sm > 48 * 1024 && sm > sm_configured) {
      SASPA_CUDA(cudaFuncSetAttribute(gn_reg_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
      sm_configured = sm;
    }
    gn_reg_kernel<<<n * rp.slices * rp.parts, GN_THREADS, sm, stream>>>(static_cast<const __nv_bfloat16*>(x), ldx, hw, c, groups, eps, gamma, beta, act,
                                                                       static_cast<__nv_bfloat16*>(y), ldy, rp.c_slice, rp.slices, rp.parts, rp.pix_per_cta,
                                                                       partials, counters);
    SASPA_LAUNCH_CHECK();
    return SASPA_OK;
  }
  SASPA_CUDA(cudaMemsetAsync(counters, 0, (size_t)n * 4, stream));
  const int pix_par = GN_THREADS / (c / 8);
  const size_t smem = sizeof(float) * 2 * (size_t)(pix_par > 1 ? pix_par : 1) * c;
  static size_t smem_configured = 0;
  if (smem > 48 * 1024 && smem > smem_configured) {
    SASPA_CUDA(cudaFuncSetAttribute(gn_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    smem_configured = smem;
  }
  gn_fused_kernel<<<n * p.ctas_per_img, GN_THREADS, smem, stream>>>(static_cast<const __nv_bfloat16*>(x), ldx, hw, c, groups, eps, gamma, beta, act,
                                                                    static_cast<__nv_bfloat16*>(y), ldy, p.ctas_per_img, p.pix_per_cta, partials, counters);
  SASPA_LAUNCH_CHECK();
  return SASPA_OK;
}

extern "C" int saspa_layernorm_bf16(const void* x, int ldx, int rows, int c, float eps, const float* gamma, const float* beta, void* y,
                                    int ldy, cudaStream_t stream) {
  SASPA_CHECK_ARG(rows >= 0 && c > 0 && c % 8 == 0 && c <= 8 * 32 * LN_MAXV, "saspa_layernorm_bf16: c must be a multiple of 8 and <= %d, got %d", 8 * 32 * LN_MAXV, c);
  SASPA_CHECK_ARG(ldx % 8 == 0 && ldy % 8 == 0, "saspa_layernorm_bf16: row strides must be multiples of 8");
  if (rows == 0) return SASPA_OK;
  SASPA_CHECK_ARG(x && y, "saspa_layernorm_bf16: null pointer");
  switch (c) {  // widths of the SD v1.5 / SDXL transformer blocks and the CLIP towers
    case 320: launch_ln_sub<8, 5>(x, ldx, rows, eps, gamma, beta, y, ldy, stream); SASPA_LAUNCH_CHECK(); return SASPA_OK;
    case 640: launch_ln_sub<16, 5>(x, ldx, rows, eps, gamma, beta, y, ldy, stream); SASPA_LAUNCH_CHECK(); return SASPA_OK;
    case 1280: launch_ln_sub<32, 5>(x, ldx, rows, eps, gamma, beta, y, ldy, stream); SASPA_LAUNCH_CHECK(); return SASPA_OK;
    case 512: launch_ln_sub<32, 2>(x, ldx, rows, eps, gamma, beta, y, ldy, stream); SASPA_LAUNCH_CHECK(); return SASPA_OK;
    case 768: launch_ln_sub<32, 3>(x, ldx, rows, eps, gamma, beta, y, ldy, stream); SASPA_LAUNCH_CHECK(); return SASPA_OK;
    case 1024: launch_ln_sub<32, 4>(x, ldx, rows, eps, gamma, beta, y, ldy, stream); SASPA_LAUNCH_CHECK(); return SASPA_OK;
    default: break;
  }
  const int grid = grid_for(rows, 8, 8);
  layernorm_kernel<<<grid, 256, 0, stream>>>(static_cast<const __nv_bfloat16*>(x), ldx, rows, c, eps, gamma, beta, static_cast<__nv_bfloat16*>(y), ldy);
  SASPA_LAUNCH_CHECK();
  return SASPA_OK;
}

extern "C" int saspa_act_bf16(const void* x, void* y, size_t count, int act, cudaStream_t stream) {
  if (count == 0) return SASPA_OK;
  SASPA_CHECK_ARG(x && y, "saspa_act_bf16: null pointer");
  SASPA_CHECK_ARG(((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(y)) & 15) == 0, "saspa_act_bf16: 16-byte alignment required");
  act_kernel<<<grid_for((long long)(count / 8 + 1), 256), 256, 0, stream>>>(static_cast<const __nv_bfloat16*>(x), static_cast<__nv_bfloat16*>(y), count, act);
  SASPA_LAUNCH_CHECK();
  return SASPA_OK;
}

extern "C" int saspa_add_bf16(const void* a, int lda, const void* b, int ldb, void* y, int ldy, int rows, int c, cudaStream_t stream) {
  SASPA_CHECK_ARG(rows >= 0 && c >= 0 && c % 8 == 0 && lda % 8 == 0 && ldb % 8 == 0 && ldy % 8 == 0, "saspa_add_bf16: c and strides must be multiples of 8");
  if (rows == 0 || c == 0) return SASPA_OK;
  SASPA_CHECK_ARG(a && b && y, "saspa_add_bf16: null pointer");
  add_kernel<<<grid_for((long long)rows * (c / 8), 256), 256, 0, stream>>>(static_cast<const __nv_bfloat16*>(a), lda, static_cast<const __nv_bfloat16*>(b), ldb,
                                                                        static_cast<__nv_bfloat16*>(y), ldy, rows, c);
  SASPA_LAUNCH_CHECK();
  return SASPA_OK;
}

extern "C" int saspa_upsample_nearest2x_bf16(const void* x, int n, int h, int w, int c, void* y, cudaStream_t stream) {
  SASPA_CHECK_ARG(n >= 0 && h >= 0 && w >= 0 && c % 8 == 0, "saspa_upsample_nearest2x_bf16: c must be a multiple of 8");
  if (n == 0 || h == 0 || w == 0 || c == 0) return SASPA_OK;
  SASPA_CHECK_ARG(x && y, "saspa_upsample_nearest2x_bf16: null pointer");
  upsample2x_kernel<<<grid_for((long long)n * 4 * h * w * (c / 8), 256), 256, 0, stream>>>(static_cast<const uint4*>(x), n, h, w, c / 8, static_cast<uint4*>(y));
  SASPA_LAUNCH_CHECK();
  return SASPA_OK;
}

extern "C" int saspa_nchw_f32_to_nhwc_bf16(const float* x, int n, int c, int h, int w, void* y, int ldy, float scale, cudaStream_t stream) {
  if (n <= 0 || c <= 0 || h <= 0 || w <= 0) return SASPA_OK;
  SASPA_CHECK_ARG(x && y && ldy >= c, "saspa_nchw_f32_to_nhwc_bf16: bad arguments");
  nchw_f32_to_nhwc_bf16_kernel<<<grid_for((long long)n * c * h * w, 256), 256, 0, stream>>>(x, n, c, h, w, static_cast<__nv_bfloat16*>(y), ldy, scale);
  SASPA_LAUNCH_CHECK();
  return SASPA_OK;
}

extern "C" int saspa_nhwc_to_nchw_f32(const void* x, int ldx, int x_is_fp32, int n, int c, int h, int w, float* y, cudaStream_t stream) {
  if (n <= 0 || c <= 0 || h <= 0 || w <= 0) return SASPA_OK;
  SASPA_CHECK_ARG(x && y && ldx >= c, "saspa_nhwc_to_nchw_f32: bad arguments");
  nhwc_to_nchw_f32_kernel<<<grid_for((long long)n * c * h * w, 256), 256, 0, stream>>>(x, ldx, x_is_fp32, n, c, h, w, y);
  SASPA_LAUNCH_CHECK();
  return SASPA_OK;
}

extern "C" int saspa_pool2d_nhwc_bf16(const void* x, int n, int h, int w, int c, int k, int stride, int pad, int is_max, void* y, int oh,
                                      int ow, cudaStream_t stream) {
  SASPA_CHECK_ARG(c % 8 == 0 && k > 0 && stride > 0 && pad >= 0, "saspa_pool2d_nhwc_bf16: bad arguments");
  if (n <= 0 || oh <= 0 || ow <= 0) return SASPA_OK;
  SASPA_CHECK_ARG(x && y, "saspa_pool2d_nhwc_bf16: null pointer");
  pool2d_kernel<<<grid_for((long long)n * oh * ow * (c / 8), 256), 256, 0, stream>>>(static_cast<const __nv_bfloat16*>(x), n, h, w, c, k, stride, pad, is_max,
                                                                                  static_cast<__nv_bfloat16*>(y), oh, ow);
  SASPA_LAUNCH_CHECK();
  return SASPA_OK;
}

extern "C" int saspa_timestep_sinusoid_bf16(const float* t, int rows, int dim, int flip_sin_to_cos, float freq_shift, void* out, cudaStream_t stream) {
  SASPA_CHECK_ARG(rows >= 0 && dim > 0 && dim % 2 == 0, "saspa_timestep_sinusoid_bf16: dim must be even");
  if (rows == 0) return SASPA_OK;
  SASPA_CHECK_ARG(t && out, "saspa_timestep_sinusoid_bf16: null pointer");
  sinusoid_kernel<<<grid_for((long long)rows * dim / 2, 128), 128, 0, stream>>>(t, rows, dim, flip_sin_to_cos, freq_shift, static_cast<__nv_bfloat16*>(out));
  SASPA_LAUNCH_CHECK();
  return SASPA_OK;
}

static int cfg_sched_launch(const float* eps_uncond, const float* eps_cond, float guidance, const saspa_lincomb* lc, const float* pre,
                            const float* post, float clip, size_t count, cudaStream_t stream);

extern "C" int saspa_cfg_sched_step(const float* eps_uncond, const float* eps_cond, float guidance, const saspa_lincomb* lc, size_t count,
                                    cudaStream_t stream) {
  return cfg_sched_launch(eps_uncond, eps_cond, guidance, lc, nullptr, nullptr, 0.0f, count, stream);
}

extern "C" int saspa_cfg_sched_step_clip(const float* eps_uncond, const float* eps_cond, float guidance, const saspa_lincomb* lc,
                                         const float* pre_host, const float* post_host, float clip_range, size_t count, cudaStream_t stream) {
  SASPA_CHECK_ARG(clip_range <= 0.0f || (pre_host && post_host), "saspa_cfg_sched_step_clip: clip_range > 0 needs pre and post rows");
  return cfg_sched_launch(eps_uncond, eps_cond, guidance, lc, pre_host, post_host, clip_range, count, stream);
}

static int cfg_sched_launch(const float* eps_uncond, const float* eps_cond, float guidance, const saspa_lincomb* lc, const float* pre,
                            const float* post, float clip, size_t count, cudaStream_t stream) {
  SASPA_CHECK_ARG(lc && eps_cond, "saspa_cfg_sched_step: null pointer");
  SASPA_CHECK_ARG(lc->n_in >= 2 && lc->n_in <= 8 && lc->n_out >= 1 && lc->n_out <= 4, "saspa_cfg_sched_step: n_in in [2,8], n_out in [1,4]");
  if (count == 0) return SASPA_OK;
  LinCombDev d;
  for (int i = 0; i < 8; ++i) d.in[i] = lc->in[i];
  for (int i = 0; i < 4; ++i) d.out[i] = lc->out[i];
  for (int i = 0; i < 32; ++i) d.coef[i] = lc->coef[i];
  d.n_in = lc->n_in;
  d.n_out = lc->n_out;
  d.clip = clip > 0.0f ? clip : 0.0f;
  for (int i = 0; i < 8; ++i) d.pre[i] = (d.clip > 0.0f && i < d.n_in) ? pre[i] : 0.0f;
  for (int j = 0; j < 4; ++j) d.post[j] = (d.clip > 0.0f && j < d.n_out) ? post[j] : 0.0f;
  for (int j = 0; j < d.n_out; ++j) SASPA_CHECK_ARG(d.out[j], "saspa_cfg_sched_step: out[%d] is null", j);
  cfg_sched_kernel<<<grid_for((long long)count, 256), 256, 0, stream>>>(eps_uncond, eps_cond, guidance, d, count);
  SASPA_LAUNCH_CHECK();
  return SASPA_OK;
}

extern "C" int saspa_vae_quantize_u8(const void* x, int ldx, int x_is_fp32, size_t pixels, uint8_t* out, cudaStream_t stream) {
  if (pixels == 0) return SASPA_OK;
  SASPA_CHECK_ARG(x && out && ldx >= 3, "saspa_vae_quantize_u8: bad arguments");
  vae_quant_kernel<<<grid_for((long long)pixels * 3, 256), 256, 0, stream>>>(x, ldx, x_is_fp32, pixels, out);
  SASPA_LAUNCH_CHECK();
  return SASPA_OK;
}

extern "C" int saspa_vae_sample_add_noise(const float* moments, const float* noise_posterior, const float* noise_diffusion, float scaling,
                                          float alpha, float sigma, int n, int h, int w, int latent_channels, float* latents, float* z0_out,
                                          cudaStream_t stream) {
  if (n <= 0 || h <= 0 || w <= 0) return SASPA_OK;
  SASPA_CHECK_ARG(moments && latents && latent_channels > 0, "saspa_vae_sample_add_noise: bad arguments");
  vae_sample_kernel<<<grid_for((long long)n * latent_channels * h * w, 256), 256, 0, stream>>>(moments, noise_posterior, noise_diffusion, scaling, alpha,
                                                                                             sigma, n, h, w, latent_channels, latents, z0_out);
  SASPA_LAUNCH_CHECK();
  return SASPA_OK;
}
