// Batched, bit-exact cv2.Canny (aperture 3, L1 gradient, multi-channel argmax) for sm_100a.
//
// Replaces the per-prompt CPU call at reference run_aug/run_aug.py:436-437
//   -> all_utils/utils.py:102-109 generate_canny -> :87-99 preprocess_canny -> :81-85 cv2.Canny.
//
// Integer/byte work, HBM-bound (algorithmic 4 B/pixel: 3 B RGB in, 1 B edge out):
//   kernel 1 (one CTA per 64x32 tile): RGB tile + 2-px halo staged in shared memory with
//     replicated borders, Sobel -> L1 magnitude -> first-max channel, NMS against a
//     zero-bordered magnitude tile, double threshold, then tile-local hysteresis in smem.
//     Writes labels {0 none, 1 weak, 2 strong} (1 B/pixel) and a per-tile "weak left" flag.
//   kernel 2 (cooperative, persistent, <= 1 wave): cross-tile hysteresis to the global fixed
//     point.  Only tiles that still hold weak pixels are revisited; a grid barrier separates
//     sweeps; three rotating "changed" flags decide termination.  The fixed point of the
//     monotone propagation is unique, so the result is deterministic and order independent.
//     Its tail converts labels to the requested outputs (u8 x1/x3 channels, bf16 {0,1} NHWC3).
#include "common.cuh"
#include "../../include/saspa_b200.h"

namespace {

constexpr int TW = 64, TH = 32, NT = 256;
constexpr int RW = TW + 4, RH = TH + 4;  // RGB staging region (2-px halo)
constexpr int MW = TW + 2, MH = TH + 2;  // magnitude region (1-px halo)
constexpr int TG22 = 13573;              // tan(22.5deg) in Q15, as OpenCV

struct CannyCtl {
  unsigned barrier;
  unsigned flags[3];
};

__device__ __forceinline__ void grid_barrier(unsigned* counter, unsigned& target) {
  __syncthreads();
  if (threadIdx.x == 0) {
    target += gridDim.x;
    __threadfence();
    atomicAdd(counter, 1u);
    unsigned v;
    do {
      asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(counter) : "memory");
    } while (v < target);
    __threadfence();
  }
  __syncthreads();
}

// Tile-local 8-connected hysteresis on s_lab (1-px border).  Returns true if any pixel was promoted.
__device__ __forceinline__ bool local_hysteresis(uint8_t (*s_lab)[MW + 2]) {
  bool any = false;
  while (true) {
    bool changed = false;
#pragma unroll
    for (int k = 0; k < (TW * TH) / NT; ++k) {
      int idx = threadIdx.x + k * NT;
      int ty = idx / TW + 1, tx = idx % TW + 1;
      if (s_lab[ty][tx] == 1) {
        bool nb = (s_lab[ty - 1][tx - 1] == 2) | (s_lab[ty - 1][tx] == 2) | (s_lab[ty - 1][tx + 1] == 2) |
                  (s_lab[ty][tx - 1] == 2) | (s_lab[ty][tx + 1] == 2) | (s_lab[ty + 1][tx - 1] == 2) |
                  (s_lab[ty + 1][tx] == 2) | (s_lab[ty + 1][tx + 1] == 2);
        if (nb) {
          s_lab[ty][tx] = 2;
          changed = true;
        }
      }
    }
    any |= changed;
    if (!__syncthreads_or(changed)) break;
  }
  return any;
}

// Write the TWxTH centre of s_lab to the global label plane; returns whether weak pixels remain.
__device__ __forceinline__ bool store_labels(uint8_t (*s_lab)[MW + 2], uint8_t* plane, int h, int w, int y0, int x0) {
  bool weak = false;
  const bool vec = ((w & 3) == 0) && ((reinterpret_cast<uintptr_t>(plane) & 3) == 0);
#pragma unroll
  for (int k = 0; k < (TW * TH / 4) / NT; ++k) {
    int q = threadIdx.x + k * NT;
    int ty = q / (TW / 4), tx = (q % (TW / 4)) * 4;
    int gy = y0 + ty, gx = x0 + tx;
    uint8_t v0 = s_lab[ty + 1][tx + 1], v1 = s_lab[ty + 1][tx + 2], v2 = s_lab[ty + 1][tx + 3], v3 = s_lab[ty + 1][tx + 4];
    if (gy < h) {
      if (vec && gx + 3 < w) {
        weak |= (v0 == 1) | (v1 == 1) | (v2 == 1) | (v3 == 1);
        uint32_t pk = v0 | (v1 << 8) | (v2 << 16) | (v3 << 24);
        *reinterpret_cast<uint32_t*>(plane + (size_t)gy * w + gx) = pk;
      } else {
        uint8_t v[4] = {v0, v1, v2, v3};
        for (int i = 0; i < 4; ++i)
          if (gx + i < w) {
            weak |= v[i] == 1;
            plane[(size_t)gy * w + gx + i] = v[i];
          }
      }
    }
  }
  return weak;
}

template <int C>
__global__ void __launch_bounds__(NT) canny_nms_kernel(const uint8_t* __restrict__ img, int h, int w, int low, int high,
                                                       uint8_t* __restrict__ lab, uint8_t* __restrict__ tile_weak) {
  constexpr int ROWB = ((RW * C + 2 + 3) / 4) * 4 + 4;  // bytes per staged row, data starts at byte 2
  __shared__ __align__(16) uint8_t s_rgb[RH][ROWB];
  __shared__ int16_t s_mag[MH][MW];
  __shared__ int s_dxy[MH][MW];
  __shared__ uint8_t s_lab[MH][MW + 2];

  const int x0 = blockIdx.x * TW, y0 = blockIdx.y * TH;
  const size_t plane = (size_t)h * w;
  const uint8_t* im = img + (size_t)blockIdx.z * plane * C;
  const int tid = threadIdx.x;

  // ---- stage RGB (replicated border) ----
  const size_t rowbytes = (size_t)w * C;
  const bool fast = (C == 3) && (x0 >= 2) && (x0 + TW + 3 <= w) && ((rowbytes & 3) == 0) &&
                    ((reinterpret_cast<uintptr_t>(im) & 3) == 0) && ((((x0 - 2) * C) & 3) == 2);
  if (fast) {
    constexpr int WORDS = (RW * C + 2 + 3) / 4;  // 52 for C == 3
    for (int i = tid; i < RH * WORDS; i += NT) {
      int ry = i / WORDS, wd = i % WORDS;
      int gy = min(max(y0 - 2 + ry, 0), h - 1);
      const uint32_t* src = reinterpret_cast<const uint32_t*>(im + (size_t)gy * rowbytes + (size_t)(x0 - 2) * C - 2);
      reinterpret_cast<uint32_t*>(&s_rgb[ry][0])[wd] = __ldg(src + wd);
    }
  } else {
    for (int i = tid; i < RH * RW * C; i += NT) {
      int ry = i / (RW * C), rb = i % (RW * C);
      int sx = rb / C, ch = rb % C;
      int gy = min(max(y0 - 2 + ry, 0), h - 1);
      int gx = min(max(x0 - 2 + sx, 0), w - 1);
      s_rgb[ry][2 + rb] = __ldg(im + ((size_t)gy * w + gx) * C + ch);
    }
  }
  for (int i = tid; i < MH * (MW + 2); i += NT) (&s_lab[0][0])[i] = 0;
  __syncthreads();

  // ---- Sobel, L1 magnitude, first-max channel ----
  for (int i = tid; i < MH * MW; i += NT) {
    int my = i / MW, mx = i % MW;
    int gy = y0 - 1 + my, gx = x0 - 1 + mx;
    int best = 0, bdx = 0, bdy = 0;
    if (gy >= 0 && gy < h && gx >= 0 && gx < w) {
      best = -1;
#pragma unroll
      for (int ch = 0; ch < C; ++ch) {
        const uint8_t* r0 = &s_rgb[my][2 + mx * C + ch];
        const uint8_t* r1 = &s_rgb[my + 1][2 + mx * C + ch];
        const uint8_t* r2 = &s_rgb[my + 2][2 + mx * C + ch];
        int tl = r0[0], tc = r0[C], tr = r0[2 * C];
        int ml = r1[0], mr = r1[2 * C];
        int bl = r2[0], bc = r2[C], br = r2[2 * C];
        int gxv = (tr + 2 * mr + br) - (tl + 2 * ml + bl);
        int gyv = (bl + 2 * bc + br) - (tl + 2 * tc + tr);
        int m = abs(gxv) + abs(gyv);
        if (m > best) {
          best = m;
          bdx = gxv;
          bdy = gyv;
        }
      }
    }
    s_mag[my][mx] = (int16_t)best;
    s_dxy[my][mx] = (bdx & 0xffff) | (bdy << 16);
  }
  __syncthreads();

  // ---- non-maximum suppression + double threshold ----
#pragma unroll
  for (int k = 0; k < (TW * TH) / NT; ++k) {
    int idx = tid + k * NT;
    int ty = idx / TW, tx = idx % TW;
    int my = ty + 1, mx = tx + 1;
    int m = s_mag[my][mx];
    uint8_t l = 0;
    if (m > low && (y0 + ty) < h && (x0 + tx) < w) {
      int pk = s_dxy[my][mx];
      int xs = (int)(int16_t)(pk & 0xffff), ys = pk >> 16;
      int ax = abs(xs);
      long long ay = (long long)abs(ys) << 15;
      long long tg22x = (long long)ax * TG22;
      bool keep;
      if (ay < tg22x) {
        keep = (m > s_mag[my][mx - 1]) && (m >= s_mag[my][mx + 1]);
      } else {
        long long tg67x = tg22x + ((long long)ax << 16);
        if (ay > tg67x) {
          keep = (m > s_mag[my - 1][mx]) && (m >= s_mag[my + 1][mx]);
        } else {
          int s = ((xs ^ ys) < 0) ? -1 : 1;
          keep = (m > s_mag[my - 1][mx - s]) && (m > s_mag[my + 1][mx + s]);
        }
      }
      if (keep) l = (m > high) ? 2 : 1;
    }
    s_lab[my][mx] = l;
  }
  __syncthreads();

  local_hysteresis(s_lab);

  bool weak = store_labels(s_lab, lab + (size_t)blockIdx.z * plane, h, w, y0, x0);
  weak = __syncthreads_or(weak);
  if (tid == 0) tile_weak[((size_t)blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x] = weak ? 1 : 0;
}

__global__ void canny_ctl_reset_kernel(CannyCtl* ctl) {
  ctl->barrier = 0;
  ctl->flags[0] = ctl->flags[1] = ctl->flags[2] = 0;
}

__global__ void __launch_bounds__(NT) canny_hysteresis_kernel(uint8_t* lab, uint8_t* tile_weak, CannyCtl* ctl, int n, int h, int w,
                                                              int tiles_x, int tiles_y, uint8_t* out_u8, int out_channels,
                                                              __nv_bfloat16* out_ctrl) {
  __shared__ uint8_t s_lab[MH][MW + 2];
  const int tid = threadIdx.x;
  const int tiles = n * tiles_x * tiles_y;
  const size_t plane = (size_t)h * w;
  unsigned bar_target = 0;

  for (int sweep = 0;; ++sweep) {
    if (blockIdx.x == 0 && tid == 0) ctl->flags[(sweep + 1) % 3] = 0;  // last read two barriers ago
    bool block_changed = false;
    for (int t = blockIdx.x; t < tiles; t += gridDim.x) {
      if (__ldcg(tile_weak + t) == 0) continue;  // uniform per block
      int im = t / (tiles_x * tiles_y), r = t % (tiles_x * tiles_y);
      int y0 = (r / tiles_x) * TH, x0 = (r % tiles_x) * TW;
      uint8_t* pl = lab + (size_t)im * plane;
      __syncthreads();
      for (int i = tid; i < MH * MW; i += NT) {
        int my = i / MW, mx = i % MW;
        int gy = y0 - 1 + my, gx = x0 - 1 + mx;
        uint8_t v = 0;
        if (gy >= 0 && gy < h && gx >= 0 && gx < w) v = __ldcg(pl + (size_t)gy * w + gx);
        s_lab[my][mx] = v;
      }
      __syncthreads();
      bool ch = local_hysteresis(s_lab);
      ch = __syncthreads_or(ch);
      if (ch) {
        bool weak = store_labels(s_lab, pl, h, w, y0, x0);
        weak = __syncthreads_or(weak);
        if (tid == 0) tile_weak[t] = weak ? 1 : 0;
        block_changed = true;
      }
    }
    if (block_changed && tid == 0) atomicOr(&ctl->flags[sweep % 3], 1u);
    grid_barrier(&ctl->barrier, bar_target);
    unsigned f;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(f) : "l"(&ctl->flags[sweep % 3]) : "memory");
    if (f == 0) break;
  }

  // ---- finalize: labels -> requested outputs ----
  const size_t total = (size_t)n * plane;
  const size_t stride = (size_t)gridDim.x * NT;
  const bool vec = (total & 3) == 0 && ((reinterpret_cast<uintptr_t>(lab) & 3) == 0) &&
                   ((reinterpret_cast<uintptr_t>(out_u8) & 3) == 0) && ((reinterpret_cast<uintptr_t>(out_ctrl) & 3) == 0);
  if (vec) {
    for (size_t q = (size_t)blockIdx.x * NT + tid; q < total / 4; q += stride) {
      uint32_t v = __ldcg(reinterpret_cast<const uint32_t*>(lab) + q);
      uint32_t e = 0;
#pragma unroll
      for (int i = 0; i < 4; ++i)
        if (((v >> (8 * i)) & 0xff) == 2) e |= 0xffu << (8 * i);
      if (out_u8) {
        if (out_channels == 1) {
          reinterpret_cast<uint32_t*>(out_u8)[q] = e;
        } else {
          uint32_t b0 = e & 0xff, b1 = (e >> 8) & 0xff, b2 = (e >> 16) & 0xff, b3 = e >> 24;
          uint32_t* o = reinterpret_cast<uint32_t*>(out_u8) + q * 3;
          o[0] = b0 | (b0 << 8) | (b0 << 16) | (b1 << 24);
          o[1] = b1 | (b1 << 8) | (b2 << 16) | (b2 << 24);
          o[2] = b2 | (b3 << 8) | (b3 << 16) | (b3 << 24);
        }
      }
      if (out_ctrl) {
        // 4 pixels x 3 channels bf16 = 24 B; 1.0bf16 = 0x3f80
        uint32_t p[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) p[i] = ((e >> (8 * i)) & 1) ? 0x3f80u : 0u;
        uint32_t* o = reinterpret_cast<uint32_t*>(out_ctrl) + q * 6;
        o[0] = p[0] | (p[0] << 16);
        o[1] = p[0] | (p[1] << 16);
        o[2] = p[1] | (p[1] << 16);
        o[3] = p[2] | (p[2] << 16);
        o[4] = p[2] | (p[3] << 16);
        o[5] = p[3] | (p[3] << 16);
      }
    }
  } else {
    for (size_t i = (size_t)blockIdx.x * NT + tid; i < total; i += stride) {
      uint8_t e = (__ldcg(lab + i) == 2) ? 255 : 0;
      if (out_u8)
        for (int c = 0; c < out_channels; ++c) out_u8[i * out_channels + c] = e;
      if (out_ctrl) {
        __nv_bfloat16 v = __float2bfloat16(e ? 1.0f : 0.0f);
        out_ctrl[i * 3 + 0] = v;
        out_ctrl[i * 3 + 1] = v;
        out_ctrl[i * 3 + 2] = v;
      }
    }
  }
}

inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

}  // namespace

extern "C" size_t saspa_canny_workspace_bytes(int n, int h, int w) {
  if (n <= 0 || h <= 0 || w <= 0) return 256;
  size_t tiles = (size_t)n * ceil_div(w, TW) * ceil_div(h, TH);
  return 256 + align_up((size_t)n * h * w, 256) + align_up(tiles, 256);
}

extern "C" int saspa_canny_u8(const uint8_t* img, int n, int h, int w, int c, int low_threshold, int high_threshold, uint8_t* out_u8,
                              int out_channels, void* out_ctrl_bf16, void* workspace, size_t workspace_bytes, cudaStream_t stream) {
  SASPA_CHECK_ARG(n >= 0 && h >= 0 && w >= 0, "saspa_canny_u8: negative shape");
  SASPA_CHECK_ARG(c == 1 || c == 3, "saspa_canny_u8: channels must be 1 or 3 (apply HWC3 first), got %d", c);
  SASPA_CHECK_ARG(out_channels == 1 || out_channels == 3, "saspa_canny_u8: out_channels must be 1 or 3");
  if (n == 0 || h == 0 || w == 0) return SASPA_OK;
  SASPA_CHECK_ARG(img && (out_u8 || out_ctrl_bf16) && workspace, "saspa_canny_u8: null pointer");
  SASPA_CHECK_ARG(h <= 65535 * TH / 1 && (long long)h * w < (1ll << 31), "saspa_canny_u8: image too large");
  if (workspace_bytes < saspa_canny_workspace_bytes(n, h, w)) {
    saspa_set_error("saspa_canny_u8: workspace too small (%zu < %zu)", workspace_bytes, saspa_canny_workspace_bytes(n, h, w));
    return SASPA_ERR_WORKSPACE;
  }
  if (low_threshold > high_threshold) {
    int t = low_threshold;
    low_threshold = high_threshold;
    high_threshold = t;
  }
  const int tiles_x = ceil_div(w, TW), tiles_y = ceil_div(h, TH);
  uint8_t* ws = static_cast<uint8_t*>(workspace);
  CannyCtl* ctl = reinterpret_cast<CannyCtl*>(ws);
  uint8_t* lab = ws + 256;
  uint8_t* tile_weak = lab + align_up((size_t)n * h * w, 256);

  canny_ctl_reset_kernel<<<1, 1, 0, stream>>>(ctl);
  SASPA_LAUNCH_CHECK();
  // grid.z is limited to 65535 images per launch
  for (int i0 = 0; i0 < n; i0 += 65535) {
    int nb = n - i0 < 65535 ? n - i0 : 65535;
    dim3 grid(tiles_x, tiles_y, nb);
    const uint8_t* src = img + (size_t)i0 * h * w * c;
    uint8_t* l = lab + (size_t)i0 * h * w;
    uint8_t* tw = tile_weak + (size_t)i0 * tiles_x * tiles_y;
    if (c == 3)
      canny_nms_kernel<3><<<grid, NT, 0, stream>>>(src, h, w, low_threshold, high_threshold, l, tw);
    else
      canny_nms_kernel<1><<<grid, NT, 0, stream>>>(src, h, w, low_threshold, high_threshold, l, tw);
    SASPA_LAUNCH_CHECK();
  }

  static int blocks_per_sm = 0;
  if (blocks_per_sm == 0) {
    SASPA_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks_per_sm, canny_hysteresis_kernel, NT, 0));
    if (blocks_per_sm <= 0) blocks_per_sm = 1;
    if (blocks_per_sm > 4) blocks_per_sm = 4;
  }
  long long tiles = (long long)n * tiles_x * tiles_y;
  long long want = ceil_div_ll((long long)n * h * w / 4, NT * 4);
  if (want < tiles) want = tiles;
  int grid = (int)(want < (long long)saspa_num_sms() * blocks_per_sm ? want : (long long)saspa_num_sms() * blocks_per_sm);
  if (grid < 1) grid = 1;
  __nv_bfloat16* ctrl = static_cast<__nv_bfloat16*>(out_ctrl_bf16);
  void* args[] = {&lab, &tile_weak, &ctl, &n, &h, &w, (void*)&tiles_x, (void*)&tiles_y, &out_u8, &out_channels, &ctrl};
  SASPA_CUDA(cudaLaunchCooperativeKernel((const void*)canny_hysteresis_kernel, dim3(grid), dim3(NT), args, 0, stream));
  return SASPA_OK;
}
