// cv2.resize(u8 HWC, dsize, interpolation = INTER_AREA | INTER_LANCZOS4) on the device, bit for bit: the resize the reference applies to every source that
// is at least `resolution` pixels on its short side (utils.resize_image, all_utils/utils.py:58-79: k <= 1 -> cv2.INTER_AREA; again inside
// preprocess_canny :93-94 and in controlnet_aux's detectors).  OpenCV takes one of three 8-bit code paths (modules/imgproc/src/resize.cpp;
// restated and pinned against the installed cv2 in oracle/cv2_area.py), all three are here:
//   mode 0  both scale factors integers        ResizeAreaFast: integer block sums; 2 x 2: (s + 2) >> 2, else cvRound(float(s) * (1.f / area))
//   mode 1  both >= 1, not both integers       ResizeArea: per-axis (source index, float weight) tables, buf += S * alpha along x (in table
//                                              order), sum (+)= beta * buf along y, cvRound -- float32, unfused multiply and add
//   mode 2  one axis < 1 (the x64 rounding of resize_image can make ONE axis a slight up-scale)
//                                              the fixed-point bilinear kernels with INTER_AREA's coefficient rule (11-bit weights)
//   lanczos (k > 1: sources under `resolution` px)  INTER_LANCZOS4: interpolateLanczos4 weights (double sin / cos on the host), 11-bit fixed point,
//                                              8 x 8 taps with a replicated border, (v + 2^21) >> 22
// The tables are built on the host in double precision exactly as OpenCV builds them and copied into the caller's workspace; one thread
// per output element (a 512 x 704 x 3 destination is 1.08 M elements: HBM-bound on the source read).
#include <math.h>

#include <vector>

#include "common.cuh"
#include "../../include/saspa_b200.h"

namespace {

struct AreaTab {
  std::vector<int> ofs;    // [dsize + 1] CSR offsets into idx / w
  std::vector<int> idx;    // source index of every entry
  std::vector<float> w;    // weight of every entry
};

// computeResizeAreaTab
AreaTab area_tab(int ssize, int dsize) {
  AreaTab t;
  const double scale = (double)ssize / dsize;
  t.ofs.push_back(0);
  for (int dx = 0; dx < dsize; ++dx) {
    const double fsx1 = dx * scale, fsx2 = fsx1 + scale;
    const double cell = fmin(scale, ssize - fsx1);
    int sx1 = (int)ceil(fsx1), sx2 = (int)floor(fsx2);
    sx2 = sx2 < ssize - 1 ? sx2 : ssize - 1;
    sx1 = sx1 < sx2 ? sx1 : sx2;
    if (sx1 - fsx1 > 1e-3) {
      t.idx.push_back(sx1 - 1);
      t.w.push_back((float)((sx1 - fsx1) / cell));
    }
    for (int sx = sx1; sx < sx2; ++sx) {
      t.idx.push_back(sx);
      t.w.push_back((float)(1.0 / cell));
    }
    if (fsx2 - sx2 > 1e-3) {
      t.idx.push_back(sx2);
      t.w.push_back((float)(fmin(fmin(fsx2 - sx2, 1.0), cell) / cell));
    }
    t.ofs.push_back((int)t.idx.size());
  }
  return t;
}

struct LinTab {
  std::vector<int> ofs;  // [dsize] left / top source index
  std::vector<int> w;    // [dsize][2] 11-bit weights
  int dmax;              // first destination index without a right / bottom neighbour
};

// the INTER_AREA branch of resize()'s general path (area_mode): bilinear taps with area coefficients
LinTab lin_tab(int ssize, int dsize) {
  LinTab t;
  const double scale = (double)ssize / dsize, inv = (double)dsize / ssize;
  t.dmax = dsize;
  for (int d = 0; d < dsize; ++d) {
    int s = (int)floor(d * scale);
    float f = (float)((d + 1) - (s + 1) * inv);
    f = f <= 0 ? 0.f : f - floorf(f);
    if (s < 0) {
      f = 0.f;
      s = 0;
    }
    if (s + 1 >= ssize) {
      t.dmax = t.dmax < d ? t.dmax : d;
      if (s >= ssize - 1) {
        f = 0.f;
        s = ssize - 1;
      }
    }
    t.ofs.push_back(s);
    t.w.push_back((int)lrintf((1.f - f) * 2048.f));
    t.w.push_back((int)lrintf(f * 2048.f));
  }
  return t;
}

struct LanczosTab {
  std::vector<int> ofs;  // [dsize] floor of the source coordinate
  std::vector<int> w;    // [dsize][8] 11-bit weights of taps ofs - 3 .. ofs + 4
};

// interpolateLanczos4 + the fixed-point conversion of resize()'s general path (INTER_LANCZOS4, 8-bit)
LanczosTab lanczos_tab(int ssize, int dsize) {
  static const double s45 = 0.70710678118654752440084436210485;
  static const double cs[8][2] = {{1, 0}, {-s45, -s45}, {0, 1}, {s45, -s45}, {-1, 0}, {s45, s45}, {0, -1}, {-s45, s45}};
  const double pi = 3.1415926535897932384626433832795;
  LanczosTab t;
  const double scale = (double)ssize / dsize;
  for (int d = 0; d < dsize; ++d) {
    float fx = (float)((d + 0.5) * scale - 0.5);
    const int s = (int)floorf(fx);
    fx -= (float)s;
    float co[8], sum = 0.f;
    const double y0 = -(fx + 3) * pi * 0.25, s0 = sin(y0), c0 = cos(y0);
    for (int i = 0; i < 8; ++i) {
      const float y0_ = (fx + 3 - i);
      if (fabsf(y0_) >= 1e-6f) {
        const double y = -y0_ * pi * 0.25;
        co[i] = (float)((cs[i][0] * s0 + cs[i][1] * c0) / (y * y));
      } else {
        co[i] = 1e30f;
      }
      sum += co[i];
    }
    sum = 1.f / sum;
    t.ofs.push_back(s);
    for (int i = 0; i < 8; ++i) {
      long v = lrintf(co[i] * sum * 2048.f);
      t.w.push_back((int)(v < -32768 ? -32768 : (v > 32767 ? 32767 : v)));
    }
  }
  return t;
}

__device__ __forceinline__ uint8_t sat_u8(int v) { return (uint8_t)(v < 0 ? 0 : (v > 255 ? 255 : v)); }

__global__ void area_fast_kernel(const uint8_t* __restrict__ src, int sw, int c, uint8_t* __restrict__ dst, int dh, int dw, int iy, int ix, float scale) {
  const long long total = (long long)dh * dw * c;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int ch = (int)(i % c);
    const int dx = (int)((i / c) % dw), dy = (int)(i / ((long long)c * dw));
    int s = 0;
    for (int y = 0; y < iy; ++y) {
      const uint8_t* row = src + ((size_t)(dy * iy + y) * sw + (size_t)dx * ix) * c + ch;
      for (int x = 0; x < ix; ++x) s += row[(size_t)x * c];
    }
    dst[i] = (ix == 2 && iy == 2) ? (uint8_t)((s + 2) >> 2) : sat_u8(__float2int_rn(__fmul_rn((float)s, scale)));
  }
}

__global__ void area_float_kernel(const uint8_t* __restrict__ src, int sw, int c, uint8_t* __restrict__ dst, int dh, int dw,
                                  const int* __restrict__ xofs, const int* __restrict__ xidx, const float* __restrict__ xw,
                                  const int* __restrict__ yofs, const int* __restrict__ yidx, const float* __restrict__ yw) {
  const long long total = (long long)dh * dw * c;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int ch = (int)(i % c);
    const int dx = (int)((i / c) % dw), dy = (int)(i / ((long long)c * dw));
    const int x0 = xofs[dx], x1 = xofs[dx + 1], y0 = yofs[dy], y1 = yofs[dy + 1];
    float sum = 0.0f;
    for (int ky = y0; ky < y1; ++ky) {
      const uint8_t* row = src + (size_t)yidx[ky] * sw * c + ch;
      float buf = 0.0f;
      for (int kx = x0; kx < x1; ++kx) buf = __fadd_rn(buf, __fmul_rn((float)row[(size_t)xidx[kx] * c], xw[kx]));
      const float t = __fmul_rn(yw[ky], buf);
      sum = ky == y0 ? t : __fadd_rn(sum, t);
    }
    dst[i] = sat_u8(__float2int_rn(sum));
  }
}

__global__ void area_linear_kernel(const uint8_t* __restrict__ src, int sh, int sw, int c, uint8_t* __restrict__ dst, int dh, int dw,
                                   const int* __restrict__ xofs, const int* __restrict__ xw, int xmax, const int* __restrict__ yofs,
                                   const int* __restrict__ yw) {
  const long long total = (long long)dh * dw * c;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int ch = (int)(i % c);
    const int dx = (int)((i / c) % dw), dy = (int)(i / ((long long)c * dw));
    const int sx = xofs[dx], a0 = xw[2 * dx], a1 = xw[2 * dx + 1];
    const int ys = yofs[dy];
    const int r0 = min(max(ys, 0), sh - 1), r1 = min(max(ys + 1, 0), sh - 1);
    const uint8_t* p0 = src + ((size_t)r0 * sw + sx) * c + ch;
    const uint8_t* p1 = src + ((size_t)r1 * sw + sx) * c + ch;
    int h0, h1;
    if (dx < xmax) {
      h0 = p0[0] * a0 + p0[c] * a1;
      h1 = p1[0] * a0 + p1[c] * a1;
    } else {
      h0 = p0[0] * 2048;
      h1 = p1[0] * 2048;
    }
    const int v = (((yw[2 * dy] * (h0 >> 4)) >> 16) + ((yw[2 * dy + 1] * (h1 >> 4)) >> 16) + 2) >> 2;
    dst[i] = sat_u8(v);
  }
}

// INTER_LANCZOS4, 8-bit: HResizeLanczos4 (int32 sums of 11-bit weights, replicated border) then VResizeLanczos4 with
// FixedPtCast<int, uchar, 22>; one thread per output element (64 taps)
__global__ void lanczos4_kernel(const uint8_t* __restrict__ src, int sh, int sw, int c, uint8_t* __restrict__ dst, int dh, int dw,
                                const int* __restrict__ xofs, const int* __restrict__ xw, const int* __restrict__ yofs, const int* __restrict__ yw) {
  const long long total = (long long)dh * dw * c;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int ch = (int)(i % c);
    const int dx = (int)((i / c) % dw), dy = (int)(i / ((long long)c * dw));
    const int sx = xofs[dx] - 3, sy = yofs[dy] - 3;
    int wx[8], cx[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      wx[k] = xw[8 * dx + k];
      cx[k] = min(max(sx + k, 0), sw - 1) * c + ch;
    }
    int v = 0;
#pragma unroll
    for (int ky = 0; ky < 8; ++ky) {
      const uint8_t* row = src + (size_t)min(max(sy + ky, 0), sh - 1) * sw * c;
      int h = 0;
#pragma unroll
      for (int k = 0; k < 8; ++k) h += row[cx[k]] * wx[k];
      v += h * yw[8 * dy + ky];
    }
    dst[i] = sat_u8((v + (1 << 21)) >> 22);
  }
}

size_t align256(size_t v) { return (v + 255) / 256 * 256; }

}  // namespace

extern "C" size_t saspa_resize_area_workspace_bytes(int sh, int sw, int dh, int dw) {
  if (sh <= 0 || sw <= 0 || dh <= 0 || dw <= 0) return 256;
  // float-area tables: at most (source + 2 * destination) entries per axis, 8 bytes each, plus the CSR offsets
  return align256((size_t)(sw + 2 * dw + sh + 2 * dh) * 8 + (size_t)(dw + dh + 2) * 4 + 1024);
}

extern "C" int saspa_resize_area_u8(const uint8_t* src, int sh, int sw, int c, uint8_t* dst, int dh, int dw, void* workspace, size_t ws_bytes,
                                    cudaStream_t stream) {
  SASPA_CHECK_ARG(sh > 0 && sw > 0 && dh > 0 && dw > 0 && c >= 1 && c <= 4, "saspa_resize_area_u8: bad shape (%d x %d x %d -> %d x %d)", sh, sw, c, dh, dw);
  SASPA_CHECK_ARG(src && dst, "saspa_resize_area_u8: null pointer");
  const long long total = (long long)dh * dw * c;
  const long long cap = (long long)saspa_num_sms() * 32, g = ceil_div_ll(total, 256);
  const int grid = (int)(g < cap ? g : cap);
  if (sh == dh && sw == dw) {
    SASPA_CUDA(cudaMemcpyAsync(dst, src, (size_t)total, cudaMemcpyDeviceToDevice, stream));
    return SASPA_OK;
  }
  const double scale_x = (double)sw / dw, scale_y = (double)sh / dh;
  if (scale_x >= 1 && scale_y >= 1 && sw % dw == 0 && sh % dh == 0) {
    const int ix = sw / dw, iy = sh / dh;
    area_fast_kernel<<<grid, 256, 0, stream>>>(src, sw, c, dst, dh, dw, iy, ix, 1.0f / (float)(ix * iy));
    SASPA_LAUNCH_CHECK();
    return SASPA_OK;
  }
  SASPA_CHECK_ARG(workspace && (reinterpret_cast<uintptr_t>(workspace) & 255) == 0, "saspa_resize_area_u8: workspace must be 256-byte aligned");
  if (ws_bytes < saspa_resize_area_workspace_bytes(sh, sw, dh, dw)) {
    saspa_set_error("saspa_resize_area_u8: workspace too small (%zu < %zu bytes)", ws_bytes, saspa_resize_area_workspace_bytes(sh, sw, dh, dw));
    return SASPA_ERR_WORKSPACE;
  }
  uint8_t* ws = static_cast<uint8_t*>(workspace);
  auto put = [&](const void* host, size_t bytes) -> void* {  // pageable -> device: the runtime stages the bytes before it returns
    void* d = ws;
    cudaMemcpyAsync(d, host, bytes, cudaMemcpyHostToDevice, stream);
    ws += (bytes + 15) / 16 * 16;
    return d;
  };
  if (scale_x >= 1 && scale_y >= 1) {
    const AreaTab tx = area_tab(sw, dw), ty = area_tab(sh, dh);
    const int* xofs = static_cast<const int*>(put(tx.ofs.data(), tx.ofs.size() * 4));
    const int* xidx = static_cast<const int*>(put(tx.idx.data(), tx.idx.size() * 4));
    const float* xw = static_cast<const float*>(put(tx.w.data(), tx.w.size() * 4));
    const int* yofs = static_cast<const int*>(put(ty.ofs.data(), ty.ofs.size() * 4));
    const int* yidx = static_cast<const int*>(put(ty.idx.data(), ty.idx.size() * 4));
    const float* yw = static_cast<const float*>(put(ty.w.data(), ty.w.size() * 4));
    SASPA_CUDA(cudaGetLastError());
    area_float_kernel<<<grid, 256, 0, stream>>>(src, sw, c, dst, dh, dw, xofs, xidx, xw, yofs, yidx, yw);
  } else {
    const LinTab tx = lin_tab(sw, dw), ty = lin_tab(sh, dh);
    const int* xofs = static_cast<const int*>(put(tx.ofs.data(), tx.ofs.size() * 4));
    const int* xw = static_cast<const int*>(put(tx.w.data(), tx.w.size() * 4));
    const int* yofs = static_cast<const int*>(put(ty.ofs.data(), ty.ofs.size() * 4));
    const int* yw = static_cast<const int*>(put(ty.w.data(), ty.w.size() * 4));
    SASPA_CUDA(cudaGetLastError());
    area_linear_kernel<<<grid, 256, 0, stream>>>(src, sh, sw, c, dst, dh, dw, xofs, xw, tx.dmax, yofs, yw);
  }
  SASPA_LAUNCH_CHECK();
  return SASPA_OK;
}

extern "C" size_t saspa_resize_lanczos4_workspace_bytes(int dh, int dw) {
  if (dh <= 0 || dw <= 0) return 256;
  return align256((size_t)(dw + dh) * 9 * 4 + 256);
}

extern "C" int saspa_resize_lanczos4_u8(const uint8_t* src, int sh, int sw, int c, uint8_t* dst, int dh, int dw, void* workspace, size_t ws_bytes,
                                        cudaStream_t stream) {
  SASPA_CHECK_ARG(sh > 0 && sw > 0 && dh > 0 && dw > 0 && c >= 1 && c <= 4, "saspa_resize_lanczos4_u8: bad shape (%d x %d x %d -> %d x %d)", sh, sw, c, dh, dw);
  SASPA_CHECK_ARG(src && dst, "saspa_resize_lanczos4_u8: null pointer");
  const long long total = (long long)dh * dw * c;
  if (sh == dh && sw == dw) {
    SASPA_CUDA(cudaMemcpyAsync(dst, src, (size_t)total, cudaMemcpyDeviceToDevice, stream));
    return SASPA_OK;
  }
  SASPA_CHECK_ARG(workspace && (reinterpret_cast<uintptr_t>(workspace) & 255) == 0, "saspa_resize_lanczos4_u8: workspace must be 256-byte aligned");
  if (ws_bytes < saspa_resize_lanczos4_workspace_bytes(dh, dw)) {
    saspa_set_error("saspa_resize_lanczos4_u8: workspace too small (%zu < %zu bytes)", ws_bytes, saspa_resize_lanczos4_workspace_bytes(dh, dw));
    return SASPA_ERR_WORKSPACE;
  }
  const LanczosTab tx = lanczos_tab(sw, dw), ty = lanczos_tab(sh, dh);
  uint8_t* ws = static_cast<uint8_t*>(workspace);
  auto put = [&](const void* host, size_t bytes) -> const int* {
    void* d = ws;
    cudaMemcpyAsync(d, host, bytes, cudaMemcpyHostToDevice, stream);
    ws += (bytes + 15) / 16 * 16;
    return static_cast<const int*>(d);
  };
  const int* xofs = put(tx.ofs.data(), tx.ofs.size() * 4);
  const int* xw = put(tx.w.data(), tx.w.size() * 4);
  const int* yofs = put(ty.ofs.data(), ty.ofs.size() * 4);
  const int* yw = put(ty.w.data(), ty.w.size() * 4);
  SASPA_CUDA(cudaGetLastError());
  const long long cap = (long long)saspa_num_sms() * 32, g = ceil_div_ll(total, 256);
  lanczos4_kernel<<<(int)(g < cap ? g : cap), 256, 0, stream>>>(src, sh, sw, c, dst, dh, dw, xofs, xw, yofs, yw);
  SASPA_LAUNCH_CHECK();
  return SASPA_OK;
}
