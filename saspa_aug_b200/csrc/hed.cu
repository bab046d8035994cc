// HED conditioning (run_aug/run_aug.py:311-312, :438-439: controlnet_aux.HEDdetector, un-vendored; the detector of the ControlNet
// annotators, restated in oracle/hed.py): the tail that turns the network's five side outputs into the control image.  The VGG trunk
// and the 1x1 projections run on the tcgen05 implicit-GEMM convolutions (ReLU epilogue) and saspa_pool2d_nhwc_bf16; this file holds
// what follows them in the detector:
//   edges_k = cv2.resize(side_k, (W, H), INTER_LINEAR)          fp32, half-pixel centres, x clamped with weight reset, rows clipped
//   edge    = 1 / (1 + exp(-mean_k(edges_k)))                    fp32 mean in numpy's reduction order, sigmoid in fp64
//   [safe]  edge = trunc(float(edge) * 3) / 2                    controlnet_aux util.safe_step(step = 2)
//   out     = u8(clip(edge * 255, 0, 255)) replicated to `out_channels` (HWC3)
// One thread per output pixel; the side maps are at most (1 + 1/4 + ... ) * 4 B per pixel and stay in L1/L2.
#include "common.cuh"
#include "../../include/saspa_b200.h"

namespace {

struct HedSides {
  const float* p[5];
  int h[5], w[5], ld[5];
  double sx[5], sy[5];  // cv2's scale_x = 1. / (dst / src), in double like resize.cpp
};

// cv2.resize INTER_LINEAR on a single-channel float map, one destination sample (HResizeLinear then VResizeLinear, no FMA contraction)
__device__ __forceinline__ float hed_sample(const float* __restrict__ s, int h, int w, int ld, double scale_x, double scale_y, int dx, int dy) {
  float fx = (float)(((double)dx + 0.5) * scale_x - 0.5);
  int ix = (int)floorf(fx);
  fx -= (float)ix;
  if (ix < 0) {
    fx = 0.0f;
    ix = 0;
  }
  if (ix >= w - 1) {
    fx = 0.0f;
    ix = w - 1;
  }
  const int ix1 = min(ix + 1, w - 1);
  float fy = (float)(((double)dy + 0.5) * scale_y - 0.5);
  const int iy = (int)floorf(fy);
  fy -= (float)iy;
  const int y0 = min(max(iy, 0), h - 1), y1 = min(max(iy + 1, 0), h - 1);
  const float a0 = 1.0f - fx, a1 = fx, b0 = 1.0f - fy, b1 = fy;
  const float* r0 = s + (size_t)y0 * w * ld;
  const float* r1 = s + (size_t)y1 * w * ld;
  const float h0 = __fadd_rn(__fmul_rn(__ldg(r0 + (size_t)ix * ld), a0), __fmul_rn(__ldg(r0 + (size_t)ix1 * ld), a1));
  const float h1 = __fadd_rn(__fmul_rn(__ldg(r1 + (size_t)ix * ld), a0), __fmul_rn(__ldg(r1 + (size_t)ix1 * ld), a1));
  return __fadd_rn(__fmul_rn(h0, b0), __fmul_rn(h1, b1));
}

__global__ void __launch_bounds__(256) hed_fuse_kernel(const HedSides sd, int H, int W, int safe, uint8_t* __restrict__ out, int out_c, long long total) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int x = (int)(i % W);
    const long long t = i / W;
    const int y = (int)(t % H);
    const long long img = t / H;
    float e[5];
#pragma unroll
    for (int k = 0; k < 5; ++k)
      e[k] = hed_sample(sd.p[k] + (size_t)img * sd.h[k] * sd.w[k] * sd.ld[k], sd.h[k], sd.w[k], sd.ld[k], sd.sx[k], sd.sy[k], x, y);
    // np.mean(float32 [H, W, 5], axis=2): first element + sequential sum of the other four (numpy's reduce inner loop), then / 5 in fp32
    const float rest = __fadd_rn(__fadd_rn(__fadd_rn(e[1], e[2]), e[3]), e[4]);
    const float m = __fdiv_rn(__fadd_rn(e[0], rest), 5.0f);
    double edge = 1.0 / (1.0 + exp(-(double)m));
    if (safe) edge = (double)((float)(int)(__fmul_rn((float)edge, 3.0f)) / 2.0f);
    double v = edge * 255.0;
    v = v < 0.0 ? 0.0 : (v > 255.0 ? 255.0 : v);
    const uint8_t q = (uint8_t)v;  // astype(np.uint8): truncation
    uint8_t* o = out + i * out_c;
    for (int c = 0; c < out_c; ++c) o[c] = q;
  }
}

}  // namespace

extern "C" int saspa_hed_fuse_u8(const float* const* side, const int* side_h, const int* side_w, const int* side_ld, int n, int H, int W, int safe,
                                 uint8_t* out, int out_channels, cudaStream_t stream) {
  SASPA_CHECK_ARG(n >= 0 && H > 0 && W > 0 && (out_channels == 1 || out_channels == 3), "saspa_hed_fuse_u8: bad shape (n=%d H=%d W=%d out_channels=%d)", n, H, W,
                  out_channels);
  if (n == 0) return SASPA_OK;
  SASPA_CHECK_ARG(side && side_h && side_w && side_ld && out, "saspa_hed_fuse_u8: null pointer");
  HedSides sd;
  for (int k = 0; k < 5; ++k) {
    SASPA_CHECK_ARG(side[k] && side_h[k] > 0 && side_w[k] > 0 && side_ld[k] > 0, "saspa_hed_fuse_u8: side output %d is empty", k);
    sd.p[k] = side[k];
    sd.h[k] = side_h[k];
    sd.w[k] = side_w[k];
    sd.ld[k] = side_ld[k];
    sd.sx[k] = 1.0 / ((double)W / (double)side_w[k]);
    sd.sy[k] = 1.0 / ((double)H / (double)side_h[k]);
  }
  const long long total = (long long)n * H * W;
  const long long cap = (long long)saspa_num_sms() * 16, g = ceil_div_ll(total, 256);
  hed_fuse_kernel<<<(int)(g < cap ? g : cap), 256, 0, stream>>>(sd, H, W, safe, out, out_channels, total);
  SASPA_LAUNCH_CHECK();
  return SASPA_OK;
}
