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
#include "../../include/saspa_b200.h"

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
// GroupNorm (NHWC).  Pass 1: per-(image, channel) partial sums -> per-(image, group) double atomics.
// Pass 2: per-channel scale/shift staged in smem, applied with 16-byte vectors, optional SiLU.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) gn_stats_kernel(const __nv_bfloat16* __restrict__ x, int ldx, int hw, int c, int groups,
                                                       int pix_per_block, double* __restrict__ stats) {
  extern __shared__ float s_acc[];  // [2][c]
  const int img = blockIdx.y;
  const int p0 = blockIdx.x * pix_per_block;
  const int p1 = min(p0 + pix_per_block, hw);
  for (int i = threadIdx.x; i < 2 * c; i += blockDim.x) s_acc[i] = 0.0f;
  __syncthreads();
  const int cv = c / 8;  // 16-byte vectors per pixel
  // thread -> fixed channel vector(s), strided over pixels: vpt vectors per thread, tpp threads per pixel
  const int vpt = (cv + blockDim.x - 1) / blockDim.x;  // 1 or 2 (c <= 4096)
  const int tpp = (cv + vpt - 1) / vpt;
  const int pix_par = blockDim.x / tpp;  // >= 1
  const int my_l = threadIdx.x % tpp;
  const int my_p = threadIdx.x / tpp;
  if (my_p < pix_par) {
#pragma unroll
    for (int jv = 0; jv < 2; ++jv) {
      const int v = my_l + jv * tpp;
      if (jv < vpt && v < cv) {
        float s[8], ss[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) s[j] = ss[j] = 0.0f;
        const __nv_bfloat16* base = x + ((size_t)img * hw) * ldx + v * 8;
        for (int p = p0 + my_p; p < p1; p += pix_par) {
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
          atomicAdd(&s_acc[v * 8 + j], s[j]);
          atomicAdd(&s_acc[c + v * 8 + j], ss[j]);
        }
      }
    }
  }
  __syncthreads();
  const int cg = c / groups;
  for (int g = threadIdx.x; g < groups; g += blockDim.x) {
    double a = 0.0, b = 0.0;
    for (int j = 0; j < cg; ++j) {
      a += (double)s_acc[g * cg + j];
      b += (double)s_acc[c + g * cg + j];
    }
    atomicAdd(&stats[((size_t)img * groups + g) * 2 + 0], a);
    atomicAdd(&stats[((size_t)img * groups + g) * 2 + 1], b);
  }
}

__global__ void __launch_bounds__(256) gn_apply_kernel(const __nv_bfloat16* __restrict__ x, int ldx, int hw, int c, int groups, float eps,
                                                       const float* __restrict__ gamma, const float* __restrict__ beta, int act,
                                                       __nv_bfloat16* __restrict__ y, int ldy, int pix_per_block,
                                                       const double* __restrict__ stats) {
  extern __shared__ float s_ab[];  // scale[c], shift[c]
  const int img = blockIdx.y;
  const int cg = c / groups;
  const double cnt = (double)hw * cg;
  for (int ch = threadIdx.x; ch < c; ch += blockDim.x) {
    int g = ch / cg;
    double sum = stats[((size_t)img * groups + g) * 2 + 0], sq = stats[((size_t)img * groups + g) * 2 + 1];
    double mean = sum / cnt;
    double var = sq / cnt - mean * mean;
    if (var < 0.0) var = 0.0;
    float rstd = (float)(1.0 / sqrt(var + (double)eps));
    float ga = gamma ? gamma[ch] : 1.0f, be = beta ? beta[ch] : 0.0f;
    s_ab[ch] = rstd * ga;
    s_ab[c + ch] = be - (float)mean * rstd * ga;
  }
  __syncthreads();
  const int cv = c / 8;
  const int p0 = blockIdx.x * pix_per_block;
  const int p1 = min(p0 + pix_per_block, hw);
  const long long total = (long long)(p1 - p0) * cv;
  for (long long i = threadIdx.x; i < total; i += blockDim.x) {
    int v = (int)(i % cv);
    int p = p0 + (int)(i / cv);
    size_t pix = (size_t)img * hw + p;
    uint4 u = __ldg(reinterpret_cast<const uint4*>(x + pix * ldx + v * 8));
    float f[8];
    unpack8(u, f);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float t = f[j] * s_ab[v * 8 + j] + s_ab[c + v * 8 + j];
      f[j] = (act == SASPA_ACT_SILU) ? silu_f(t) : t;
    }
    *reinterpret_cast<uint4*>(y + pix * ldy + v * 8) = pack8(f);
  }
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
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      if (j < lc.n_out) {
        float acc = 0.0f;
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

extern "C" int saspa_groupnorm_nhwc_bf16(const void* x, int ldx, int n, int hw, int c, int groups, float eps, const float* gamma,
                                         const float* beta, int act, void* y, int ldy, void* stats_ws, cudaStream_t stream) {
  SASPA_CHECK_ARG(n >= 0 && hw >= 0 && c > 0 && groups > 0 && c % groups == 0, "saspa_groupnorm_nhwc_bf16: bad shape (c=%d groups=%d)", c, groups);
  SASPA_CHECK_ARG(c % 8 == 0 && ldx % 8 == 0 && ldy % 8 == 0 && c <= 4096, "saspa_groupnorm_nhwc_bf16: c, ldx, ldy must be multiples of 8 and c <= 4096, got c=%d", c);
  SASPA_CHECK_ARG(act == SASPA_ACT_NONE || act == SASPA_ACT_SILU, "saspa_groupnorm_nhwc_bf16: act must be NONE or SILU");
  if (n == 0 || hw == 0) return SASPA_OK;
  SASPA_CHECK_ARG(x && y && stats_ws, "saspa_groupnorm_nhwc_bf16: null pointer");
  SASPA_CHECK_ARG((reinterpret_cast<uintptr_t>(stats_ws) & 7) == 0, "saspa_groupnorm_nhwc_bf16: stats_ws must be 8-byte aligned");
  double* stats = reinterpret_cast<double*>(stats_ws);
  SASPA_CUDA(cudaMemsetAsync(stats, 0, sizeof(double) * 2 * (size_t)n * groups, stream));
  // ~4 CTAs per SM over the whole batch
  int blocks_x = ceil_div(saspa_num_sms() * 4, n);
  int pix_per_block = ceil_div(hw, blocks_x);
  if (pix_per_block < 32) pix_per_block = 32;
  blocks_x = ceil_div(hw, pix_per_block);
  dim3 grid(blocks_x, n);
  size_t smem = sizeof(float) * 2 * c;
  gn_stats_kernel<<<grid, 256, smem, stream>>>(static_cast<const __nv_bfloat16*>(x), ldx, hw, c, groups, pix_per_block, stats);
  SASPA_LAUNCH_CHECK();
  gn_apply_kernel<<<grid, 256, smem, stream>>>(static_cast<const __nv_bfloat16*>(x), ldx, hw, c, groups, eps, gamma, beta, act,
                                               static_cast<__nv_bfloat16*>(y), ldy, pix_per_block, stats);
  SASPA_LAUNCH_CHECK();
  return SASPA_OK;
}

extern "C" int saspa_layernorm_bf16(const void* x, int ldx, int rows, int c, float eps, const float* gamma, const float* beta, void* y,
                                    int ldy, cudaStream_t stream) {
  SASPA_CHECK_ARG(rows >= 0 && c > 0 && c % 8 == 0 && c <= 8 * 32 * LN_MAXV, "saspa_layernorm_bf16: c must be a multiple of 8 and <= %d, got %d", 8 * 32 * LN_MAXV, c);
  SASPA_CHECK_ARG(ldx % 8 == 0 && ldy % 8 == 0, "saspa_layernorm_bf16: row strides must be multiples of 8");
  if (rows == 0) return SASPA_OK;
  SASPA_CHECK_ARG(x && y, "saspa_layernorm_bf16: null pointer");
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

extern "C" int saspa_cfg_sched_step(const float* eps_uncond, const float* eps_cond, float guidance, const saspa_lincomb* lc, size_t count,
                                    cudaStream_t stream) {
  SASPA_CHECK_ARG(lc && eps_cond, "saspa_cfg_sched_step: null pointer");
  SASPA_CHECK_ARG(lc->n_in >= 2 && lc->n_in <= 8 && lc->n_out >= 1 && lc->n_out <= 4, "saspa_cfg_sched_step: n_in in [2,8], n_out in [1,4]");
  if (count == 0) return SASPA_OK;
  LinCombDev d;
  for (int i = 0; i < 8; ++i) d.in[i] = lc->in[i];
  for (int i = 0; i < 4; ++i) d.out[i] = lc->out[i];
  for (int i = 0; i < 32; ++i) d.coef[i] = lc->coef[i];
  d.n_in = lc->n_in;
  d.n_out = lc->n_out;
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
