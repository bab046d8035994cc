// Filter-side kernels (integer resize + classifier / CLIP heads).
//
//   saspa_pil_coeffs_host / saspa_resize_pil_u8 : bit-exact Pillow antialiased resize -- the arithmetic behind
//       torchvision Resize in all_utils/dataset_utils.py:78-85 (bilinear 256x256) and openai-clip's _transform
//       (bicubic 224), applied at all_utils/utils.py:360 and :171/:404.
//   saspa_crop_normalize_bf16 : CenterCrop + ToTensor + Normalize, fused with the cast to NHWC bf16.
//   saspa_bap_head            : WS-DAN bilinear attention pooling (fgvc/models/cal.py:53-86).
//   saspa_fc_f32              : the 65536-wide classifier FC (fgvc/models/cal.py:164, :212), HBM-bound on W.
//   saspa_topk_contains       : `correct_label in logits.topk(k)[1]` (all_utils/utils.py:363).
//   saspa_clip_score_argmax   : CLIP_selector.forward cosine logits + argmax (all_utils/utils.py:152-177).
#include <math.h>
#include <stdlib.h>

#include "common.cuh"
#include "../../include/saspa_b200.h"

namespace {

constexpr int PRECISION_BITS = 32 - 8 - 2;

double filt_bilinear(double x) {
  if (x < 0.0) x = -x;
  if (x < 1.0) return 1.0 - x;
  return 0.0;
}
double filt_bicubic(double x) {
  const double a = -0.5;
  if (x < 0.0) x = -x;
  if (x < 1.0) return ((a + 2.0) * x - (a + 3.0)) * x * x + 1;
  if (x < 2.0) return (((x - 5) * x + 8) * x - 4) * a;
  return 0.0;
}

// One pass of the separable resample along `axis_len` -> `out_len`.
//   in  [outer, axis_len, inner] u8,  out [outer, out_len, inner] u8
__global__ void resample_pass_kernel(const uint8_t* __restrict__ in, uint8_t* __restrict__ out, long long outer, int axis_len, int out_len, int inner,
                                     const int32_t* __restrict__ bounds, const int32_t* __restrict__ coeffs, int ksize) {
  const long long total = outer * out_len * inner;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    int in_i = (int)(i % inner);
    long long t = i / inner;
    int o = (int)(t % out_len);
    long long ou = t / out_len;
    int lo = __ldg(bounds + 2 * o), cnt = __ldg(bounds + 2 * o + 1);
    const int32_t* k = coeffs + (size_t)o * ksize;
    const uint8_t* src = in + (ou * axis_len + lo) * inner + in_i;
    int32_t ss = 1 << (PRECISION_BITS - 1);
    for (int x = 0; x < cnt; ++x) ss += (int32_t)__ldg(src + (size_t)x * inner) * __ldg(k + x);
    int v = ss >> PRECISION_BITS;
    out[i] = (uint8_t)(v < 0 ? 0 : (v > 255 ? 255 : v));
  }
}

__global__ void crop_normalize_kernel(const uint8_t* __restrict__ img, int h, int w, int cy, int cx, int ch, int cw, float m0, float m1, float m2,
                                      float s0, float s1, float s2, __nv_bfloat16* __restrict__ out, int out_c, long long total_pix) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total_pix; i += (long long)gridDim.x * blockDim.x) {
    int x = (int)(i % cw);
    long long t = i / cw;
    int y = (int)(t % ch);
    long long n = t / ch;
    const uint8_t* p = img + ((n * h + (cy + y)) * w + (cx + x)) * 3;
    __nv_bfloat16* o = out + i * out_c;
    // ToTensor: u8 / 255 (fp32), Normalize: (v - mean) / std
    o[0] = __float2bfloat16(((float)p[0] / 255.0f - m0) / s0);
    o[1] = __float2bfloat16(((float)p[1] / 255.0f - m1) / s1);
    o[2] = __float2bfloat16(((float)p[2] / 255.0f - m2) / s2);
    for (int c = 3; c < out_c; ++c) o[c] = __float2bfloat16(0.0f);
  }
}

// BAP: p[i][j] = sum_hw att[hw][i] * feat[hw][j] / hw ; sign-sqrt; (L2 norm applied by bap_norm_kernel)
// grid (c / 256, n); block 256 threads: thread j owns feature channel j, loops over hw with att row in smem.
__global__ void __launch_bounds__(256) bap_pool_kernel(const __nv_bfloat16* __restrict__ feat, int ldf, const __nv_bfloat16* __restrict__ att, int lda,
                                                       int hw, int c, int m, float* __restrict__ fm, float* __restrict__ sq_sum) {
  extern __shared__ float s_att[];  // [hw][m]
  const int img = blockIdx.y;
  for (int i = threadIdx.x; i < hw * m; i += blockDim.x) {
    int p = i / m, a = i % m;
    s_att[i] = __bfloat162float(att[((size_t)img * hw + p) * lda + a]);
  }
  __syncthreads();
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  float local_sq = 0.0f;
  if (j < c) {
    float acc[32];
    for (int a0 = 0; a0 < m; a0 += 32) {
      const int ma = min(32, m - a0);
#pragma unroll
      for (int a = 0; a < 32; ++a) acc[a] = 0.0f;
      for (int p = 0; p < hw; ++p) {
        float f = __bfloat162float(feat[((size_t)img * hw + p) * ldf + j]);
#pragma unroll
        for (int a = 0; a < 32; ++a)
          if (a < ma) acc[a] += s_att[p * m + a0 + a] * f;
      }
#pragma unroll
      for (int a = 0; a < 32; ++a)
        if (a < ma) {
          float v = acc[a] / (float)hw;
          float sgn = (v > 0.0f) ? 1.0f : ((v < 0.0f) ? -1.0f : 0.0f);
          float r = sgn * sqrtf(fabsf(v) + 1e-6f);
          fm[(size_t)img * m * c + (size_t)(a0 + a) * c + j] = r;
          local_sq += r * r;
        }
    }
  }
  local_sq = warp_sum(local_sq);
  if ((threadIdx.x & 31) == 0) atomicAdd(&sq_sum[img], local_sq);
}

__global__ void bap_norm_kernel(float* __restrict__ fm, const float* __restrict__ sq_sum, long long per_img, float mul) {
  const int img = blockIdx.y;
  // F.normalize: x / max(||x||, 1e-12)
  const float inv = mul / fmaxf(sqrtf(sq_sum[img]), 1e-12f);
  float* p = fm + (size_t)img * per_img;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < per_img; i += (long long)gridDim.x * blockDim.x) p[i] *= inv;
}

// logits[n][cls] = x[n][:] . W[cls][:] + b ; one CTA per class row chunk, all n images at once (n small), W streamed once.
constexpr int FC_MAX_N = 8;
__global__ void __launch_bounds__(256) fc_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ bias, int n0,
                                                 int nb, int k, int classes, float* __restrict__ logits) {
  const int cls = blockIdx.x;
  float acc[FC_MAX_N];
#pragma unroll
  for (int i = 0; i < FC_MAX_N; ++i) acc[i] = 0.0f;
  const float* wr = w + (size_t)cls * k;
  for (int c = threadIdx.x * 8; c < k; c += blockDim.x * 8) {
    float4 w0 = __ldg(reinterpret_cast<const float4*>(wr + c)), w1 = __ldg(reinterpret_cast<const float4*>(wr + c) + 1);
    float wf[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
#pragma unroll
    for (int i = 0; i < FC_MAX_N; ++i) {
      if (i < nb) {
        const float4* xp = reinterpret_cast<const float4*>(x + (size_t)(n0 + i) * k + c);
        float4 a = __ldg(xp), b = __ldg(xp + 1);
        acc[i] += a.x * wf[0] + a.y * wf[1] + a.z * wf[2] + a.w * wf[3] + b.x * wf[4] + b.y * wf[5] + b.z * wf[6] + b.w * wf[7];
      }
    }
  }
  __shared__ float red[FC_MAX_N][8];
#pragma unroll
  for (int i = 0; i < FC_MAX_N; ++i) {
    float v = warp_sum(acc[i]);
    if ((threadIdx.x & 31) == 0) red[i][threadIdx.x >> 5] = v;
  }
  __syncthreads();
  if (threadIdx.x < nb) {
    float v = 0.0f;
    for (int wq = 0; wq < 8; ++wq) v += red[threadIdx.x][wq];
    logits[(size_t)(n0 + threadIdx.x) * classes + cls] = v + (bias ? bias[cls] : 0.0f);
  }
}

// rank of the label = #{j : logit[j] > logit[label]  or (logit[j] == logit[label] and j < label)}; keep iff rank < k.
__global__ void topk_contains_kernel(const float* __restrict__ logits, int n, int classes, const int32_t* __restrict__ label, int k,
                                     uint8_t* __restrict__ keep, float* __restrict__ margin) {
  const int img = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (img >= n) return;
  const float* row = logits + (size_t)img * classes;
  const int lab = label[img];
  const float lv = row[lab];
  int rank = 0;
  for (int j = lane; j < classes; j += 32) {
    float v = row[j];
    rank += (v > lv) || (v == lv && j < lab);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) rank += __shfl_xor_sync(0xffffffffu, rank, o);
  if (lane == 0) {
    keep[img] = rank < k ? 1 : 0;
    if (margin) margin[img] = (float)(k - rank);  // > 0: inside the top-k by that many places
  }
}

// softmax(logits[i, :])[idx[i]], max_j logits[i, j] and its (first) argmax: the confidence tests of the reference's optional filters
// (all_utils/utils.py:186-191 get_clip_filtering, :370-376 filter_confidence_higher_than, :411-418 alia_conf_filtering).
__global__ void softmax_at_kernel(const float* __restrict__ logits, int n, int classes, const int32_t* __restrict__ idx, float* __restrict__ prob,
                                  float* __restrict__ max_logit, int32_t* __restrict__ argmax) {
  const int i = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (i >= n) return;
  const float* row = logits + (size_t)i * classes;
  float m = -INFINITY;
  int mj = 0x7fffffff;
  for (int j = lane; j < classes; j += 32) {
    const float v = row[j];
    if (v > m) {
      m = v;
      mj = j;
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float om = __shfl_xor_sync(0xffffffffu, m, o);
    const int oj = __shfl_xor_sync(0xffffffffu, mj, o);
    if (om > m || (om == m && oj < mj)) {  // first maximum wins, as torch.argmax / max
      m = om;
      mj = oj;
    }
  }
  float s = 0.0f;
  for (int j = lane; j < classes; j += 32) s += expf(row[j] - m);
  s = warp_sum(s);
  if (lane == 0) {
    if (prob) prob[i] = expf(row[idx[i]] - m) / s;
    if (max_logit) max_logit[i] = m;
    if (argmax) argmax[i] = mj;
  }
}

__global__ void clip_score_kernel(const float* __restrict__ img, const float* __restrict__ txt, int n, int p, int d, float logit_scale,
                                  float* __restrict__ logits, int32_t* __restrict__ argmax) {
  const int i = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (i >= n) return;
  const float* a = img + (size_t)i * d;
  float na = 0.0f;
  for (int c = lane; c < d; c += 32) na += a[c] * a[c];
  na = sqrtf(warp_sum(na));
  float best = -INFINITY;
  int best_j = 0;
  for (int j = 0; j < p; ++j) {
    const float* t = txt + (size_t)j * d;
    float dot = 0.0f, nt = 0.0f;
    for (int c = lane; c < d; c += 32) {
      dot += a[c] * t[c];
      nt += t[c] * t[c];
    }
    dot = warp_sum(dot);
    nt = sqrtf(warp_sum(nt));
    float l = logit_scale * dot / (na * nt);
    if (lane == 0 && logits) logits[(size_t)i * p + j] = l;
    if (l > best) {  // first maximum wins, as torch.argmax
      best = l;
      best_j = j;
    }
  }
  if (lane == 0) argmax[i] = best_j;
}

}  // namespace

extern "C" int saspa_pil_ksize(int in_size, int out_size, int filter) {
  if (in_size <= 0 || out_size <= 0) return -1;
  double filterscale = (double)in_size / out_size;
  if (filterscale < 1.0) filterscale = 1.0;
  double support = (filter ? 2.0 : 1.0) * filterscale;
  return (int)ceil(support) * 2 + 1;
}

extern "C" int saspa_pil_coeffs_host(int in_size, int out_size, int filter, int* ksize_out, int32_t* bounds, int32_t* coeffs, int coeffs_capacity) {
  SASPA_CHECK_ARG(in_size > 0 && out_size > 0 && (filter == 0 || filter == 1), "saspa_pil_coeffs_host: bad arguments");
  SASPA_CHECK_ARG(bounds && coeffs, "saspa_pil_coeffs_host: null pointer");
  double (*f)(double) = filter ? filt_bicubic : filt_bilinear;
  double scale, filterscale;
  filterscale = scale = (double)in_size / out_size;
  if (filterscale < 1.0) filterscale = 1.0;
  const double support = (filter ? 2.0 : 1.0) * filterscale;
  const int ksize = (int)ceil(support) * 2 + 1;
  SASPA_CHECK_ARG(coeffs_capacity >= out_size * ksize, "saspa_pil_coeffs_host: coefficient buffer too small");
  double* k = (double*)malloc(sizeof(double) * ksize);
  if (!k) return SASPA_ERR_ARG;
  for (int xx = 0; xx < out_size; ++xx) {
    double center = (xx + 0.5) * scale;
    double ww = 0.0, ss = 1.0 / filterscale;
    int xmin = (int)(center - support + 0.5);
    if (xmin < 0) xmin = 0;
    int xmax = (int)(center + support + 0.5);
    if (xmax > in_size) xmax = in_size;
    xmax -= xmin;
    int x;
    for (x = 0; x < xmax; ++x) {
      double w = f((x + xmin - center + 0.5) * ss);
      k[x] = w;
      ww += w;
    }
    for (x = 0; x < xmax; ++x)
      if (ww != 0.0) k[x] /= ww;
    for (; x < ksize; ++x) k[x] = 0;
    for (x = 0; x < ksize; ++x) {
      if (k[x] < 0)
        coeffs[xx * ksize + x] = (int32_t)(-0.5 + k[x] * (1 << PRECISION_BITS));
      else
        coeffs[xx * ksize + x] = (int32_t)(0.5 + k[x] * (1 << PRECISION_BITS));
    }
    bounds[xx * 2] = xmin;
    bounds[xx * 2 + 1] = xmax;
  }
  free(k);
  if (ksize_out) *ksize_out = ksize;
  return SASPA_OK;
}

extern "C" int saspa_resize_pil_u8(const uint8_t* img, int n, int h, int w, int c, uint8_t* tmp, uint8_t* out, int out_h, int out_w,
                                   const int32_t* bounds_x, const int32_t* coeffs_x, int ksize_x, const int32_t* bounds_y,
                                   const int32_t* coeffs_y, int ksize_y, cudaStream_t stream) {
  SASPA_CHECK_ARG(n >= 0 && h > 0 && w > 0 && c > 0 && out_h > 0 && out_w > 0, "saspa_resize_pil_u8: bad shape");
  if (n == 0) return SASPA_OK;
  SASPA_CHECK_ARG(img && tmp && out && bounds_x && coeffs_x && bounds_y && coeffs_y, "saspa_resize_pil_u8: null pointer");
  // horizontal pass: [n*h, w, c] -> [n*h, out_w, c]; vertical pass: [n, h, out_w*c] -> [n, out_h, out_w*c]
  long long t1 = (long long)n * h * out_w * c, t2 = (long long)n * out_h * out_w * c;
  long long cap = (long long)saspa_num_sms() * 16;
  long long g1 = ceil_div_ll(t1, 256), g2 = ceil_div_ll(t2, 256);
  resample_pass_kernel<<<(int)(g1 < cap ? g1 : cap), 256, 0, stream>>>(img, tmp, (long long)n * h, w, out_w, c, bounds_x, coeffs_x, ksize_x);
  SASPA_LAUNCH_CHECK();
  resample_pass_kernel<<<(int)(g2 < cap ? g2 : cap), 256, 0, stream>>>(tmp, out, n, h, out_h, out_w * c, bounds_y, coeffs_y, ksize_y);
  SASPA_LAUNCH_CHECK();
  return SASPA_OK;
}

extern "C" int saspa_crop_normalize_bf16(const uint8_t* img, int n, int h, int w, int crop_y, int crop_x, int crop_h, int crop_w, float mean0,
                                         float mean1, float mean2, float std0, float std1, float std2, void* out, int out_c,
                                         cudaStream_t stream) {
  SASPA_CHECK_ARG(n >= 0 && crop_y >= 0 && crop_x >= 0 && crop_y + crop_h <= h && crop_x + crop_w <= w && out_c >= 3, "saspa_crop_normalize_bf16: bad crop");
  if (n == 0 || crop_h == 0 || crop_w == 0) return SASPA_OK;
  SASPA_CHECK_ARG(img && out, "saspa_crop_normalize_bf16: null pointer");
  long long total = (long long)n * crop_h * crop_w;
  long long cap = (long long)saspa_num_sms() * 16, g = ceil_div_ll(total, 256);
  crop_normalize_kernel<<<(int)(g < cap ? g : cap), 256, 0, stream>>>(img, h, w, crop_y, crop_x, crop_h, crop_w, mean0, mean1, mean2, std0, std1, std2,
                                                                    static_cast<__nv_bfloat16*>(out), out_c, total);
  SASPA_LAUNCH_CHECK();
  return SASPA_OK;
}

extern "C" int saspa_bap_head(const void* feat, int ldf, const void* att, int lda, int n, int hw, int c, int m, float* fm, float* sq_ws, cudaStream_t stream) {
  SASPA_CHECK_ARG(n >= 0 && hw > 0 && c > 0 && m > 0, "saspa_bap_head: bad shape");
  if (n == 0) return SASPA_OK;
  SASPA_CHECK_ARG(feat && att && fm && sq_ws, "saspa_bap_head: null pointer");
  size_t smem = sizeof(float) * (size_t)hw * m;
  SASPA_CHECK_ARG(smem <= 200 * 1024, "saspa_bap_head: hw*m too large for shared memory");
  static bool configured = false;
  if (!configured) {
    SASPA_CUDA(cudaFuncSetAttribute(bap_pool_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    configured = true;
  }
  SASPA_CUDA(cudaMemsetAsync(sq_ws, 0, sizeof(float) * n, stream));
  dim3 grid(ceil_div(c, 256), n);
  bap_pool_kernel<<<grid, 256, smem, stream>>>(static_cast<const __nv_bfloat16*>(feat), ldf, static_cast<const __nv_bfloat16*>(att), lda, hw, c, m, fm, sq_ws);
  SASPA_LAUNCH_CHECK();
  bap_norm_kernel<<<dim3(64, n), 256, 0, stream>>>(fm, sq_ws, (long long)m * c, 100.0f);
  SASPA_LAUNCH_CHECK();
  return SASPA_OK;
}

extern "C" int saspa_fc_f32(const float* x, const float* w, const float* bias, int n, int k, int classes, float* logits, cudaStream_t stream) {
  SASPA_CHECK_ARG(n >= 0 && k > 0 && classes > 0 && k % 8 == 0, "saspa_fc_f32: k must be a multiple of 8");
  if (n == 0) return SASPA_OK;
  SASPA_CHECK_ARG(x && w && logits, "saspa_fc_f32: null pointer");
  for (int n0 = 0; n0 < n; n0 += FC_MAX_N) {
    int nb = n - n0 < FC_MAX_N ? n - n0 : FC_MAX_N;
    fc_kernel<<<classes, 256, 0, stream>>>(x, w, bias, n0, nb, k, classes, logits);
    SASPA_LAUNCH_CHECK();
  }
  return SASPA_OK;
}

extern "C" int saspa_topk_contains(const float* logits, int n, int classes, const int32_t* label, int k, uint8_t* keep, float* margin, cudaStream_t stream) {
  SASPA_CHECK_ARG(n >= 0 && classes > 0 && k > 0, "saspa_topk_contains: bad shape");
  if (n == 0) return SASPA_OK;
  SASPA_CHECK_ARG(logits && label && keep, "saspa_topk_contains: null pointer");
  topk_contains_kernel<<<ceil_div(n, 4), 128, 0, stream>>>(logits, n, classes, label, k, keep, margin);
  SASPA_LAUNCH_CHECK();
  return SASPA_OK;
}

extern "C" int saspa_softmax_at_f32(const float* logits, int n, int classes, const int32_t* idx, float* prob, float* max_logit, int32_t* argmax,
                                    cudaStream_t stream) {
  SASPA_CHECK_ARG(n >= 0 && classes > 0, "saspa_softmax_at_f32: bad shape");
  if (n == 0) return SASPA_OK;
  SASPA_CHECK_ARG(logits && (idx || !prob), "saspa_softmax_at_f32: null pointer");
  softmax_at_kernel<<<ceil_div(n, 4), 128, 0, stream>>>(logits, n, classes, idx, prob, max_logit, argmax);
  SASPA_LAUNCH_CHECK();
  return SASPA_OK;
}

extern "C" int saspa_clip_score_argmax(const float* img, const float* txt, int n, int p, int d, float logit_scale, float* logits, int32_t* argmax,
                                       cudaStream_t stream) {
  SASPA_CHECK_ARG(n >= 0 && p > 0 && d > 0, "saspa_clip_score_argmax: bad shape");
  if (n == 0) return SASPA_OK;
  SASPA_CHECK_ARG(img && txt && argmax, "saspa_clip_score_argmax: null pointer");
  clip_score_kernel<<<ceil_div(n, 4), 128, 0, stream>>>(img, txt, n, p, d, logit_scale, logits, argmax);
  SASPA_LAUNCH_CHECK();
  return SASPA_OK;
}
