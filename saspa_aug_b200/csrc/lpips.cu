// LPIPS distance pieces of the reference's optional filter (all_utils/utils.py:269-270, :377-381, calc_lpips_distance :576-590;
// algorithm of the un-vendored `lpips` package, restated in oracle/lpips_alex.py): the PIL "L" conversion that precedes the resize, and
// the per-layer score  mean_pixels sum_c w_c (f0_c / |f0| - f1_c / |f1|)^2  on NHWC bf16 feature maps.  The AlexNet convolutions run on the
// tcgen05 GEMM (im2col for the 11x11 / 5x5 stems), the max pools on saspa_pool2d_nhwc_bf16.
#include "common.cuh"
#include "../../include/saspa_b200.h"

namespace {

// PIL Image.convert("L") (ITU-R 601-2 luma, 16-bit fixed point, libImaging Convert.c rgb2l) followed by convert("RGB") (replication)
__global__ void luma3_kernel(const uint8_t* __restrict__ img, uint8_t* __restrict__ out, long long pixels) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < pixels; i += (long long)gridDim.x * blockDim.x) {
    const uint8_t* p = img + i * 3;
    const uint32_t l = ((uint32_t)p[0] * 19595u + (uint32_t)p[1] * 38470u + (uint32_t)p[2] * 7471u + 0x8000u) >> 16;
    uint8_t* o = out + i * 3;
    o[0] = o[1] = o[2] = (uint8_t)l;
  }
}

constexpr int LP_THREADS = 256;
constexpr int LP_MAXV = 2;  // 16-byte vectors per lane: c <= 8 * 32 * LP_MAXV = 512

// One CTA per image; a warp owns pixels warp, warp + 8, ...; its lanes own the channel vectors.  Fixed pixel order per warp and fixed
// warp order in the final sum: deterministic.  accum[img] += mean over pixels (launches of the five layers are stream-ordered).
__global__ void __launch_bounds__(LP_THREADS) lpips_layer_kernel(const __nv_bfloat16* __restrict__ f0, const __nv_bfloat16* __restrict__ f1,
                                                                 const float* __restrict__ w, int hw, int c, float* __restrict__ accum) {
  __shared__ float s_part[LP_THREADS / 32];
  const int img = blockIdx.x, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int cv = c / 8;
  const __nv_bfloat16* a = f0 + (size_t)img * hw * c;
  const __nv_bfloat16* b = f1 + (size_t)img * hw * c;
  float wv[LP_MAXV][8];
#pragma unroll
  for (int k = 0; k < LP_MAXV; ++k) {
    const int v = lane + 32 * k;
#pragma unroll
    for (int j = 0; j < 8; ++j) wv[k][j] = v < cv ? __ldg(w + v * 8 + j) : 0.0f;
  }
  float total = 0.0f;
  for (int p = warp; p < hw; p += LP_THREADS / 32) {
    float x[LP_MAXV][8], y[LP_MAXV][8];
    float sx = 0.0f, sy = 0.0f;
#pragma unroll
    for (int k = 0; k < LP_MAXV; ++k) {
      const int v = lane + 32 * k;
      if (v < cv) {
        const uint4 ua = __ldg(reinterpret_cast<const uint4*>(a + (size_t)p * c + v * 8));
        const uint4 ub = __ldg(reinterpret_cast<const uint4*>(b + (size_t)p * c + v * 8));
        x[k][0] = bf16_lo(ua.x); x[k][1] = bf16_hi(ua.x); x[k][2] = bf16_lo(ua.y); x[k][3] = bf16_hi(ua.y);
        x[k][4] = bf16_lo(ua.z); x[k][5] = bf16_hi(ua.z); x[k][6] = bf16_lo(ua.w); x[k][7] = bf16_hi(ua.w);
        y[k][0] = bf16_lo(ub.x); y[k][1] = bf16_hi(ub.x); y[k][2] = bf16_lo(ub.y); y[k][3] = bf16_hi(ub.y);
        y[k][4] = bf16_lo(ub.z); y[k][5] = bf16_hi(ub.z); y[k][6] = bf16_lo(ub.w); y[k][7] = bf16_hi(ub.w);
      } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) x[k][j] = y[k][j] = 0.0f;
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        sx = fmaf(x[k][j], x[k][j], sx);
        sy = fmaf(y[k][j], y[k][j], sy);
      }
    }
    sx = warp_sum(sx);
    sy = warp_sum(sy);
    const float ix = 1.0f / (sqrtf(sx) + 1e-10f), iy = 1.0f / (sqrtf(sy) + 1e-10f);
    float d = 0.0f;
#pragma unroll
    for (int k = 0; k < LP_MAXV; ++k) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float t = __fsub_rn(__fmul_rn(x[k][j], ix), __fmul_rn(y[k][j], iy));  // no FMA contraction: identical inputs give exactly 0
        d = fmaf(wv[k][j] * t, t, d);
      }
    }
    total += warp_sum(d);
  }
  if (lane == 0) s_part[warp] = total;
  __syncthreads();
  if (threadIdx.x == 0) {
    float s = 0.0f;
    for (int k = 0; k < LP_THREADS / 32; ++k) s += s_part[k];
    accum[img] += s / (float)hw;
  }
}

}  // namespace

extern "C" int saspa_rgb_to_luma3_u8(const uint8_t* img, long long pixels, uint8_t* out, cudaStream_t stream) {
  SASPA_CHECK_ARG(pixels >= 0, "saspa_rgb_to_luma3_u8: bad size");
  if (pixels == 0) return SASPA_OK;
  SASPA_CHECK_ARG(img && out, "saspa_rgb_to_luma3_u8: null pointer");
  long long g = ceil_div_ll(pixels, 256), cap = (long long)saspa_num_sms() * 16;
  luma3_kernel<<<(int)(g < cap ? g : cap), 256, 0, stream>>>(img, out, pixels);
  SASPA_LAUNCH_CHECK();
  return SASPA_OK;
}

extern "C" int saspa_lpips_layer_accum(const void* f0, const void* f1, const float* w, int n, int hw, int c, float* accum, cudaStream_t stream) {
  SASPA_CHECK_ARG(n >= 0 && hw > 0 && c > 0 && c % 8 == 0 && c <= 8 * 32 * LP_MAXV, "saspa_lpips_layer_accum: c must be a multiple of 8 and <= %d (got %d)",
                  8 * 32 * LP_MAXV, c);
  if (n == 0) return SASPA_OK;
  SASPA_CHECK_ARG(f0 && f1 && w && accum, "saspa_lpips_layer_accum: null pointer");
  lpips_layer_kernel<<<n, LP_THREADS, 0, stream>>>(static_cast<const __nv_bfloat16*>(f0), static_cast<const __nv_bfloat16*>(f1), w, hw, c, accum);
  SASPA_LAUNCH_CHECK();
  return SASPA_OK;
}
