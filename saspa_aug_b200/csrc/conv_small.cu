// Direct 3x3 convolution for SMALL channel counts (Cin <= 32, Cout <= 128) on full-resolution maps: the ControlNet conditioning
// embedding (3->16, 16->16, 16->32 s2, 32->32, 32->96 s2 at 512^2 / 256^2; diffusers ControlNetConditioningEmbedding,
// models/controlnets/controlnet.py), the VAE encoder stem (3->128) and the HED stem (3->64).  These layers have K = 9 * Cin <= 288 and
// a few output channels: on the 128 x BN x 64 tcgen05 tiles three quarters of every k-block and most of the N tile were padding
// (14 TF/s on the 16->16 layer, 3 TF/s + an im2col buffer on the 3->16 one: 7.4 ms per 32 images, profiles/r2_shapes_mb32_*.txt), while the
// layers themselves are HBM-bound (537 MB for 16->16 at 32 x 512^2: 83 us at the copy rate).  Here a CTA owns a 16-pixel-wide strip of output
// rows: the input halo tile and the whole weight tensor sit in shared memory, each warp runs bf16 mma.sync.m16n8k16 (fp32 accumulate) over
// the nine taps with ldmatrix fragments, and bias + activation + bf16 rounding happen on the accumulator fragments, which are stored
// straight to NHWC global memory.  Persistent over tiles: the weights are staged once per CTA.
#include "common.cuh"
#include "../../include/saspa_b200.h"

namespace {

constexpr int CS_THREADS = 256;
constexpr int TW = 16;  // output pixels per mma row tile

__device__ __forceinline__ void cs_ldmatrix_x4(uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3, const void* p) {
  const uint32_t s = (uint32_t)__cvta_generic_to_shared(p);
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(s));
}
__device__ __forceinline__ void cs_mma(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ float cs_act(float v, int act) {
  switch (act) {
    case SASPA_ACT_SILU: return silu_f(v);
    case SASPA_ACT_RELU: return fmaxf(v, 0.0f);
    case SASPA_ACT_GELU: return gelu_erf_f(v);
    case SASPA_ACT_QUICKGELU: return quick_gelu_f(v);
    default: return v;
  }
}

struct ConvSmallParams {
  const __nv_bfloat16* x;
  int ldx, cin, H, W, n_img;
  const __nv_bfloat16* w;  // [cout][kpad], k = (ky, kx, cin)
  int kpad;
  const float* bias;
  int act;
  const __nv_bfloat16* residual;  // [n, OH, OW, cout] with pixel stride ld_res, added before the activation (or nullptr)
  int ld_res;
  __nv_bfloat16* out;
  int ldo, cout, OH, OW;
  int pad;
  int tiles_x, tiles_y;
  long long total_tiles;
};

// CIN_PAD: channels per pixel in shared memory (16 | 32, zero filled above cin); NT: 8-wide output-channel tiles held in registers;
// STRIDE 1 | 2.  Pixel / weight rows are (CIN_PAD + 8) bf16 apart: 48 B / 80 B strides keep the eight 16-byte rows of every ldmatrix
// phase on distinct bank groups.
template <int CIN_PAD, int NT, int STRIDE>
__global__ void __launch_bounds__(CS_THREADS) conv3x3_small_kernel(const ConvSmallParams p) {
  constexpr int TR = STRIDE == 1 ? 16 : 8;  // output rows per tile (8 warps: 2 rows | 1 row each)
  constexpr int IN_W = (TW - 1) * STRIDE + 3, IN_H = (TR - 1) * STRIDE + 3;
  constexpr int PITCH = (CIN_PAD + 8) * 2;  // bytes
  constexpr int CH = CIN_PAD / 8;           // 16-byte chunks of real channels per pixel
  constexpr int ROWS_PER_WARP = TR / (CS_THREADS / 32);
  extern __shared__ __align__(16) uint8_t cs_smem[];
  uint8_t* sW = cs_smem;                         // [9][NT * 8][PITCH]
  uint8_t* sIn = cs_smem + 9 * NT * 8 * PITCH;   // [IN_H][IN_W][PITCH]
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  // ---- weights: once per CTA ----
  for (int i = tid; i < 9 * NT * 8 * CH; i += CS_THREADS) {
    const int ch = i % CH, co = (i / CH) % (NT * 8), tap = i / (CH * NT * 8);
    uint4 v = make_uint4(0u, 0u, 0u, 0u);
    if (co < p.cout) {
      const __nv_bfloat16* src = p.w + (size_t)co * p.kpad + tap * p.cin + ch * 8;
      if ((p.cin & 7) == 0) {
        if (ch * 8 < p.cin) v = __ldg(reinterpret_cast<const uint4*>(src));
      } else {
        __align__(16) __nv_bfloat16 e[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) e[j] = (ch * 8 + j < p.cin) ? src[j] : __float2bfloat16(0.0f);
        v = *reinterpret_cast<const uint4*>(e);
      }
    }
    *reinterpret_cast<uint4*>(sW + (size_t)(tap * NT * 8 + co) * PITCH + ch * 16) = v;
  }

  float bz[NT][2];  // this lane's bias pair of every output-channel tile
#pragma unroll
  for (int j = 0; j < NT; ++j) {
    const int c = j * 8 + (lane & 3) * 2;
    bz[j][0] = (p.bias && c < p.cout) ? __ldg(p.bias + c) : 0.0f;
    bz[j][1] = (p.bias && c + 1 < p.cout) ? __ldg(p.bias + c + 1) : 0.0f;
  }

  for (long long tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
    const int tx = (int)(tile % p.tiles_x);
    const long long t2 = tile / p.tiles_x;
    const int ty = (int)(t2 % p.tiles_y), img = (int)(t2 / p.tiles_y);
    const int ox0 = tx * TW, oy0 = ty * TR;
    const int ix0 = ox0 * STRIDE - p.pad, iy0 = oy0 * STRIDE - p.pad;
    __syncthreads();  // previous tile's fragments are out of sIn (and the weights are in place)
    // ---- input halo tile, zero outside the image (= the convolution's padding) ----
    const __nv_bfloat16* xi = p.x + (size_t)img * p.H * p.W * p.ldx;
    for (int i = tid; i < IN_H * IN_W * CH; i += CS_THREADS) {
      const int ch = i % CH, px = (i / CH) % IN_W, py = i / (CH * IN_W);
      const int y = iy0 + py, x = ix0 + px;
      uint4 v = make_uint4(0u, 0u, 0u, 0u);
      if (y >= 0 && y < p.H && x >= 0 && x < p.W) {
        const __nv_bfloat16* src = xi + ((size_t)y * p.W + x) * p.ldx + ch * 8;
        if ((p.cin & 7) == 0 && (p.ldx & 7) == 0) {
          if (ch * 8 < p.cin) v = __ldg(reinterpret_cast<const uint4*>(src));
        } else {
          __align__(16) __nv_bfloat16 e[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) e[j] = (ch * 8 + j < p.cin) ? src[j] : __float2bfloat16(0.0f);
          v = *reinterpret_cast<const uint4*>(e);
        }
      }
      *reinterpret_cast<uint4*>(sIn + (size_t)(py * IN_W + px) * PITCH + ch * 16) = v;
    }
    __syncthreads();

#pragma unroll 1
    for (int rr = 0; rr < ROWS_PER_WARP; ++rr) {
      const int ly = warp * ROWS_PER_WARP + rr;  // output row inside the tile
      float acc[NT][4];
#pragma unroll
      for (int j = 0; j < NT; ++j) acc[j][0] = acc[j][1] = acc[j][2] = acc[j][3] = 0.0f;
#pragma unroll
      for (int tap = 0; tap < 9; ++tap) {
        const int ky = tap / 3, kx = tap - 3 * ky;
#pragma unroll
        for (int ks = 0; ks < CIN_PAD / 16; ++ks) {
          uint32_t a[4];
          {  // A: 16 output pixels x 16 channels; lane -> (pixel lane % 16, channel chunk lane / 16)
            const int m = lane & 15, kc = lane >> 4;
            cs_ldmatrix_x4(a[0], a[1], a[2], a[3], sIn + (size_t)((ly * STRIDE + ky) * IN_W + m * STRIDE + kx) * PITCH + (ks * 16 + kc * 8) * 2);
          }
#pragma unroll
          for (int j = 0; j < NT; j += 2) {
            uint32_t b0, b1, b2, b3;  // B: (cout tile j: k 0-7, k 8-15), (cout tile j + 1: k 0-7, k 8-15)
            const int co = j * 8 + (lane & 7) + ((lane >> 4) << 3), kc = (lane >> 3) & 1;
            cs_ldmatrix_x4(b0, b1, b2, b3, sW + (size_t)(tap * NT * 8 + co) * PITCH + (ks * 16 + kc * 8) * 2);
            cs_mma(acc[j], a, b0, b1);
            cs_mma(acc[j + 1], a, b2, b3);
          }
        }
      }
      // ---- epilogue on the fragments: c0,c1 -> pixel lane / 4, channels 2 * (lane % 4) + {0, 1}; c2,c3 -> pixel + 8 ----
      const int oy = oy0 + ly;
      if (oy < p.OH) {
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          const int ox = ox0 + (lane >> 2) + half * 8;
          if (ox >= p.OW) continue;
          const size_t pix = ((size_t)img * p.OH + oy) * p.OW + ox;
          __nv_bfloat16* op = p.out + pix * p.ldo;
          const __nv_bfloat16* rp = p.residual ? p.residual + pix * p.ld_res : nullptr;
#pragma unroll
          for (int j = 0; j < NT; ++j) {
            const int c = j * 8 + (lane & 3) * 2;
            if (c >= p.cout) continue;
            float v0 = acc[j][half * 2] + bz[j][0], v1 = acc[j][half * 2 + 1] + bz[j][1];
            if (rp) {
              v0 += __bfloat162float(rp[c]);
              if (c + 1 < p.cout) v1 += __bfloat162float(rp[c + 1]);
            }
            v0 = cs_act(v0, p.act);
            v1 = cs_act(v1, p.act);
            if (c + 1 < p.cout && (p.ldo & 1) == 0) {
              *reinterpret_cast<__nv_bfloat162*>(op + c) = __floats2bfloat162_rn(v0, v1);
            } else {
              op[c] = __float2bfloat16(v0);
              if (c + 1 < p.cout) op[c + 1] = __float2bfloat16(v1);
            }
          }
        }
      }
    }
  }
}

template <int CIN_PAD, int NT, int STRIDE>
int launch_small(const ConvSmallParams& p, cudaStream_t stream) {
  constexpr int TR = STRIDE == 1 ? 16 : 8;
  constexpr int IN_W = (TW - 1) * STRIDE + 3, IN_H = (TR - 1) * STRIDE + 3;
  constexpr int PITCH = (CIN_PAD + 8) * 2;
  constexpr int SMEM = (9 * NT * 8 + IN_H * IN_W) * PITCH;
  static_assert(SMEM <= 200 * 1024, "shared memory budget");
  static int per_sm = 0;  // co-resident CTAs per SM (registers and shared memory): the persistent grid is exactly one wave
  if (per_sm == 0) {
    if (SMEM > 48 * 1024) SASPA_CUDA(cudaFuncSetAttribute(conv3x3_small_kernel<CIN_PAD, NT, STRIDE>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
    int occ = 0;
    SASPA_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, conv3x3_small_kernel<CIN_PAD, NT, STRIDE>, CS_THREADS, SMEM));
    per_sm = occ < 1 ? 1 : occ;
  }
  ConvSmallParams q = p;
  q.tiles_x = ceil_div(p.OW, TW);
  q.tiles_y = ceil_div(p.OH, TR);
  q.total_tiles = (long long)q.tiles_x * q.tiles_y * p.n_img;
  const long long cap = (long long)saspa_num_sms() * per_sm;
  conv3x3_small_kernel<CIN_PAD, NT, STRIDE><<<(int)(q.total_tiles < cap ? q.total_tiles : cap), CS_THREADS, SMEM, stream>>>(q);
  SASPA_LAUNCH_CHECK();
  return SASPA_OK;
}

template <int CIN_PAD, int STRIDE>
int dispatch_nt(const ConvSmallParams& p, cudaStream_t stream) {
  if (p.cout <= 16) return launch_small<CIN_PAD, 2, STRIDE>(p, stream);
  if (p.cout <= 32) return launch_small<CIN_PAD, 4, STRIDE>(p, stream);
  if (p.cout <= 64) return launch_small<CIN_PAD, 8, STRIDE>(p, stream);
  if (p.cout <= 96) return launch_small<CIN_PAD, 12, STRIDE>(p, stream);
  return launch_small<CIN_PAD, 16, STRIDE>(p, stream);
}

}  // namespace

extern "C" int saspa_conv3x3_small_supported(int cin, int cout, int stride, int pad) {
  return (cin >= 1 && cin <= 32 && cout >= 1 && cout <= 128 && (stride == 1 || stride == 2) && pad == 1) ? 1 : 0;
}

extern "C" int saspa_conv3x3_small_bf16(const void* x, int ldx, int cin, int n, int h, int w, const void* weight, int kpad, const float* bias, int act,
                                        const void* residual, int ld_res, int stride, int pad, void* out, int ldo, int cout, int oh, int ow,
                                        cudaStream_t stream) {
  SASPA_CHECK_ARG(n >= 0 && h >= 0 && w >= 0 && oh >= 0 && ow >= 0, "saspa_conv3x3_small_bf16: negative dims");
  if (n == 0 || oh == 0 || ow == 0) return SASPA_OK;
  SASPA_CHECK_ARG(x && weight && out, "saspa_conv3x3_small_bf16: null pointer");
  SASPA_CHECK_ARG(saspa_conv3x3_small_supported(cin, cout, stride, pad), "saspa_conv3x3_small_bf16: needs cin <= 32, cout <= 128, stride 1 | 2, pad 1 (cin=%d cout=%d stride=%d pad=%d)",
                  cin, cout, stride, pad);
  SASPA_CHECK_ARG(ldx >= cin && ldo >= cout && kpad >= 9 * cin && (!residual || ld_res >= cout),
                  "saspa_conv3x3_small_bf16: ldx >= cin, ldo >= cout, ld_res >= cout, kpad >= 9 * cin");
  SASPA_CHECK_ARG((cin % 8 != 0) || (kpad % 8 == 0 && (reinterpret_cast<uintptr_t>(weight) & 15) == 0 && (reinterpret_cast<uintptr_t>(x) & 15) == 0),
                  "saspa_conv3x3_small_bf16: cin %% 8 == 0 takes 16-byte loads (aligned x / weight, kpad %% 8 == 0)");
  SASPA_CHECK_ARG(act == SASPA_ACT_NONE || act == SASPA_ACT_SILU || act == SASPA_ACT_RELU || act == SASPA_ACT_GELU || act == SASPA_ACT_QUICKGELU,
                  "saspa_conv3x3_small_bf16: unsupported activation %d", act);
  ConvSmallParams p = {};
  p.x = static_cast<const __nv_bfloat16*>(x);
  p.ldx = ldx;
  p.cin = cin;
  p.H = h;
  p.W = w;
  p.n_img = n;
  p.w = static_cast<const __nv_bfloat16*>(weight);
  p.kpad = kpad;
  p.bias = bias;
  p.act = act;
  p.residual = static_cast<const __nv_bfloat16*>(residual);
  p.ld_res = ld_res;
  p.out = static_cast<__nv_bfloat16*>(out);
  p.ldo = ldo;
  p.cout = cout;
  p.OH = oh;
  p.OW = ow;
  p.pad = pad;
  if (cin <= 16) return stride == 1 ? dispatch_nt<16, 1>(p, stream) : dispatch_nt<16, 2>(p, stream);
  return stride == 1 ? dispatch_nt<32, 1>(p, stream) : dispatch_nt<32, 2>(p, stream);
}
