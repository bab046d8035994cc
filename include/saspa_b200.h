/* saspa_b200 -- C ABI of the B200-native (sm_100a) SaSPA augmentation-generation hot path.
 *
 * The reference (EyalMichaeli/SaSPA-Aug) is pure Python and has no FFI of its own: the
 * boundary this library sits behind is the set of third-party calls the reference makes on
 * the path  run_aug/run_aug.py -> cv2.Canny -> diffusers pipeline -> filter -> aug JSON.
 * Each entry point below cites the reference call site (file:line under /root/reference)
 * whose arithmetic it replaces.  INTEGRATION.md shows the ctypes binding.
 *
 * Conventions
 *   - extern "C", plain pointers and sizes, no torch types.  Every pointer is a DEVICE pointer
 *     unless its name ends in _host.  The caller owns all buffers, including workspaces
 *     (size from the matching *_workspace_bytes query).
 *   - Every function returns 0 on success, a negative SASPA_ERR_* for argument errors, or a
 *     positive cudaError_t.  Nothing throws, exits or synchronises; work is enqueued on `stream`.
 *     saspa_last_error_string() returns the calling thread's last message.
 *   - bf16 tensors are passed as void*; activations are NHWC ("token major"): a 2-D view
 *     [pixels, channels] with a row stride in elements.  GEMM/conv weights are K-major
 *     [Cout, K] with K = kh*kw*Cin ordered (kh, kw, cin).
 *   - No mutable global state: every entry point is a function of its arguments (a per-device once-initialised cache of function
 *     attributes aside), re-entrant across streams and threads, CUDA-graph capturable.  The kernel-selection overrides used by tests and
 *     A/B timing are NOT part of this ABI; they live in saspa_aug_b200/csrc/tuning_hooks.h.
 *   - No CPU fallback exists: on a machine without an sm_100 device every compute entry point
 *     returns a CUDA error.
 */
#ifndef SASPA_B200_H_
#define SASPA_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SASPA_B200_VERSION 100

#ifndef __CUDA_RUNTIME_H__
typedef struct CUstream_st* cudaStream_t;
#endif

int saspa_version(void);
const char* saspa_last_error_string(void);

/* ------------------------------------------------------------------------------------------
 * Canny edge conditioning.  Replaces cv2.Canny(img, low, high) as called by
 * all_utils/utils.py:81-85 (CannyDetector.__call__) via generate_canny (:102-109), driven from
 * run_aug/run_aug.py:436-437.  Bit-exact (aperture 3, L1 gradient, per-pixel max-magnitude
 * channel, zero-bordered NMS, 8-connected hysteresis).
 *   img        u8 [n,h,w,c], c in {1,3}
 *   out_u8     u8 [n,h,w,out_channels] in {0,255}; out_channels 1 (cv2.Canny) or 3 (HWC3,
 *              all_utils/utils.py:95); may be NULL if out_ctrl_bf16 is given
 *   out_ctrl_bf16  optional bf16 [n,h,w,3] in {0,1}: the ControlNet conditioning tensor the
 *              diffusers image processor would build from the PIL image (/255, no normalise)
 * ------------------------------------------------------------------------------------------ */
size_t saspa_canny_workspace_bytes(int n, int h, int w);
int saspa_canny_u8(const uint8_t* img, int n, int h, int w, int c, int low_threshold, int high_threshold,
                   uint8_t* out_u8, int out_channels, void* out_ctrl_bf16, void* workspace, size_t workspace_bytes,
                   cudaStream_t stream);

/* ------------------------------------------------------------------------------------------
 * PIL-exact antialiased resize (+ crop + normalise) used by the filter's preprocessing:
 * all_utils/dataset_utils.py:78-85 (Resize((256,256)) bilinear -> CenterCrop(224) -> ToTensor
 * -> Normalize) at all_utils/utils.py:360, and openai-clip's _transform (Resize(224, bicubic)
 * -> CenterCrop(224) -> Normalize) at all_utils/utils.py:171.  The u8 resize is bit-exact
 * (22-bit fixed-point coefficients, uint8 intermediate between the two passes).
 *   filter: 0 bilinear, 1 bicubic.  coefficient tables are built on the host by
 *   saspa_pil_coeffs_host (same doubles as Pillow) and uploaded by the caller.
 * ------------------------------------------------------------------------------------------ */
int saspa_pil_coeffs_host(int in_size, int out_size, int filter, int* ksize_out, int32_t* bounds_host, int32_t* coeffs_host,
                          int coeffs_capacity);
int saspa_pil_ksize(int in_size, int out_size, int filter);
int saspa_resize_pil_u8(const uint8_t* img, int n, int h, int w, int c, uint8_t* tmp /* [n,h,out_w,c] */, uint8_t* out, int out_h,
                        int out_w, const int32_t* bounds_x, const int32_t* coeffs_x, int ksize_x, const int32_t* bounds_y,
                        const int32_t* coeffs_y, int ksize_y, cudaStream_t stream);
/* cv2.resize(u8 HWC, (dw, dh), interpolation = cv2.INTER_AREA) of ONE image on the device, bit for bit on OpenCV's three 8-bit code paths
 * (integer block sums / float area tables / fixed-point bilinear with area coefficients when one axis is up-scaled): the resize of
 * utils.resize_image for sources at least `resolution` px on their short side (all_utils/utils.py:58-79, k <= 1; called at
 * run_aug/run_aug.py:373 and again inside preprocess_canny, all_utils/utils.py:93-94).  src [sh, sw, c], dst [dh, dw, c], c <= 4, both
 * dense.  The per-axis tables are built on the host as OpenCV builds them and staged in `workspace`
 * (saspa_resize_area_workspace_bytes, 256-byte aligned; unused when both scale factors are integers or the sizes are equal). */
size_t saspa_resize_area_workspace_bytes(int sh, int sw, int dh, int dw);
int saspa_resize_area_u8(const uint8_t* src, int sh, int sw, int c, uint8_t* dst, int dh, int dw, void* workspace, size_t ws_bytes,
                         cudaStream_t stream);

/* cv2.resize(u8 HWC, (dw, dh), interpolation = cv2.INTER_LANCZOS4) of ONE image on the device, bit for bit: the k > 1 branch of
 * utils.resize_image (sources under `resolution` px on their short side; all_utils/utils.py:77).  interpolateLanczos4 weights built on
 * the host (double sin / cos, float normalisation, 11-bit fixed point) as OpenCV builds them, 8 x 8 taps with a replicated border in
 * int32, FixedPtCast<int, uchar, 22>.  Same buffer conventions as saspa_resize_area_u8. */
size_t saspa_resize_lanczos4_workspace_bytes(int dh, int dw);
int saspa_resize_lanczos4_u8(const uint8_t* src, int sh, int sw, int c, uint8_t* dst, int dh, int dw, void* workspace, size_t ws_bytes,
                             cudaStream_t stream);

/* LPIPS distance of the optional lpips_min / lpips_max filter (all_utils/utils.py:269-270, :377-381, calc_lpips_distance :576-590;
 * arithmetic of the un-vendored `lpips` package, net='alex').
 *   saspa_rgb_to_luma3_u8: PIL Image.convert("L").convert("RGB") (ITU-R 601-2 luma in 16-bit fixed point, replicated), u8 [pixels,3].
 *   saspa_lpips_layer_accum: accum[i] += mean_pixels sum_c w[c] * (f0[i,p,c] / |f0[i,p,:]| - f1[i,p,c] / |f1[i,p,:]|)^2 for one AlexNet
 *   layer's NHWC bf16 feature maps [n, hw, c] (c % 8 == 0, c <= 512); the five layers are accumulated by five stream-ordered calls. */
int saspa_rgb_to_luma3_u8(const uint8_t* img, long long pixels, uint8_t* out, cudaStream_t stream);
int saspa_lpips_layer_accum(const void* f0, const void* f1, const float* w, int n, int hw, int c, float* accum, cudaStream_t stream);

/* HED conditioning (run_aug/run_aug.py:311-312, :438-439: controlnet_aux.HEDdetector.__call__, un-vendored): the detector's tail.
 * side[k] (k = 0..4): the network's fp32 side outputs [n, side_h[k], side_w[k]] with a pixel stride of side_ld[k] floats (the 1x1
 * projections' output rows).  Each is resized to H x W like cv2.resize(INTER_LINEAR) on a float map, the five are averaged (fp32,
 * numpy's reduction order), edge = sigmoid (fp64); safe != 0 applies controlnet_aux's safe_step(edge, 2);
 * out u8 [n, H, W, out_channels] (1 or 3 = HWC3 replication) = trunc(clip(edge * 255, 0, 255)).  `side` and the three int arrays are
 * HOST arrays of five entries. */
int saspa_hed_fuse_u8(const float* const* side, const int* side_h, const int* side_w, const int* side_ld, int n, int H, int W, int safe,
                      uint8_t* out, int out_channels, cudaStream_t stream);

/* u8 [n,h,w,3] -> crop -> (x/255 - mean[c]) / std[c] -> bf16 NHWC [n,crop_h,crop_w,out_c] (channels >= 3 zero padded) */
int saspa_crop_normalize_bf16(const uint8_t* img, int n, int h, int w, int crop_y, int crop_x, int crop_h, int crop_w, float mean0,
                              float mean1, float mean2, float std0, float std1, float std2, void* out, int out_c,
                              cudaStream_t stream);

/* ------------------------------------------------------------------------------------------
 * Dense contractions (tcgen05.mma, TMEM accumulators, TMA-fed, persistent, warp-specialised).
 * Replace the cuBLAS / cuDNN calls under diffusers' UNet2DConditionModel / ControlNetModel /
 * AutoencoderKL forward (pipe(**pipe_args), run_aug/run_aug.py:278) and under the filter nets
 * (all_utils/utils.py:361, :152-164).
 *
 * Epilogue (applied in this order, per output element (row r, col j)):
 *     folded LayerNorm (ln_stats): acc = rstd_r * (acc - mean_r * ln_colsum[j]) -- A holds the UN-normalised rows, B = W * gamma,
 *                      bias = b + W . beta; (mean_r, rstd_r) come from the partial row sums a producer GEMM wrote (row_stats_out)
 *     v = acc + bias[j] + row_bias[r / rows_per_group, j]
 *     v = act(v)                       (SASPA_ACT_*; skipped here when act_after_residual)
 *     GEGLU: out col j pairs value col j with gate col j + BN/2 of the same tile: v = val * gelu(gate)
 *     v = alpha * v + beta * residual[r, j]
 *     v = act(v)                       (only when act_after_residual)
 *     out[r, j] = (bf16 | fp32) v
 *     row_stats_out[r, 2 * n_tile + g] = (sum_j out[r, j], sum_j out[r, j]^2) over the columns epilogue group g of N tile n_tile
 *                      stored, taken on the bf16-ROUNDED values (what the next GEMM reads)
 * ------------------------------------------------------------------------------------------ */
enum { SASPA_ACT_NONE = 0, SASPA_ACT_SILU = 1, SASPA_ACT_GELU = 2, SASPA_ACT_RELU = 3, SASPA_ACT_QUICKGELU = 4, SASPA_ACT_GEGLU = 5 };

typedef struct saspa_epilogue {
  const float* bias;       /* [N] or NULL */
  const float* row_bias;   /* [groups, N] or NULL: e.g. the per-image time-embedding projection */
  int rows_per_group;      /* rows (pixels) per row_bias group; ignored when row_bias is NULL */
  int ld_row_bias;         /* row stride of row_bias in elements (0 -> N): lets one GEMM produce the time-embedding
                              projections of every ResnetBlock and each conv read its column slice */
  int act;                 /* SASPA_ACT_* */
  float alpha;             /* scale on the activated value (1/output_scale_factor, conditioning_scale, ...) */
  const void* residual;    /* bf16 [M, ld_res] or NULL */
  int ld_res;
  float beta;              /* scale on the residual */
  int out_fp32;            /* 0: bf16 output, 1: fp32 output */
  int act_after_residual;  /* 0: act before the residual add (diffusers blocks); 1: after it (ResNet bottleneck ReLU) */
  /* LayerNorm folded into the GEMMs around it (diffusers BasicTransformerBlock norm1/2/3 -> to_q|k|v / to_q / ff.net.0.proj,
   * models/attention.py): the producer of the residual stream emits per-row partial sums, the consumer applies the statistics
   * to its accumulator -- the normalised tensor never exists in HBM.  GEMM entry point only. */
  void* row_stats_out;     /* float2 [M, row_stats_slots] or NULL */
  int row_stats_slots;     /* = saspa_gemm_row_stats_slots(N) */
  const void* ln_stats;    /* float2 [M, ln_slots] (a producer's row_stats_out for this GEMM's A rows) or NULL */
  int ln_slots;
  const float* ln_colsum;  /* fp32 [N]: sum_k B[j, k] of the bf16 weights */
  float ln_eps;
} saspa_epilogue;

/* D[M,N] = epilogue(A[M,K] . B[N,K]^T).  A, B bf16 row-major (K contiguous); lda/ldb/ldd in elements,
 * lda % 8 == ldb % 8 == 0, 16-byte aligned bases.  For SASPA_ACT_GEGLU, B must be tile-interleaved by
 * saspa_geglu_interleave_rows semantics and D is [M, N/2]. */
int saspa_gemm_bf16(const void* A, int lda, const void* B, int ldb, void* D, int ldd, int M, int N, int K,
                    const saspa_epilogue* ep_host, cudaStream_t stream);

/* Direct 3x3 convolution for small channel counts (cin <= 32, cout <= 128, stride 1 | 2, padding 1), NHWC bf16 -> NHWC bf16 with
 * bias + activation (SASPA_ACT_NONE | SILU | RELU | GELU | QUICKGELU) in the epilogue: the ControlNet conditioning embedding's first five
 * layers (diffusers ControlNetConditioningEmbedding, models/controlnets/controlnet.py; once per image before the denoise loop of
 * run_aug/run_aug.py:278), the VAE encoder's conv_in and the HED stem.  x [n, h, w, cin] with pixel stride ldx, weight bf16
 * [cout, kpad] in (ky, kx, cin) order (the layout of saspa_conv2d_igemm_bf16 / saspa_im2col_bf16 weights), out [n, oh, ow, cout] with
 * pixel stride ldo; residual (bf16 [n, oh, ow, cout], pixel stride ld_res, or NULL) is added before the activation (the ControlNet adds its
 * conditioning embedding to conv_in's output).  Wider layers run as slices of output channels (the host side uses 64: two CTAs per SM) (weight rows / out / residual offset by
 * the caller: the UNet's and ControlNet's conv_in 4 -> 320).  Replaces im2col + GEMM (cin = 3) and the 64-channel-granular implicit GEMM (cin = 16 / 32) on these layers. */
int saspa_conv3x3_small_supported(int cin, int cout, int stride, int pad);
int saspa_conv3x3_small_bf16(const void* x, int ldx, int cin, int n, int h, int w, const void* weight, int kpad, const float* bias, int act,
                             const void* residual, int ld_res, int stride, int pad, void* out, int ldo, int cout, int oh, int ow,
                             cudaStream_t stream);

/* Slots per row of the partial statistics a GEMM with N output columns writes to row_stats_out (a function of N only, so that a
 * row's statistics do not depend on how many rows share the launch). */
int saspa_gemm_row_stats_slots(int N);

/* Implicit-GEMM convolution, stride 1, "same" zero padding, ksize in {1,3}, NHWC:
 *   x0 [n,h,w,c0] (pixel stride ldx0) and optionally x1 [n,h,w,c1] (pixel stride ldx1) are read as the
 *   channel concatenation [x0, x1] (skip-concat of the UNet up path without materialising it);
 *   weight bf16 [cout, ksize*ksize*(c0+c1)] K-major (kh, kw, cin); out [n,h,w,cout] (pixel stride ldo).
 *   c0 % 64 == 0 when x1 is given; c0 % 8 == 0, c1 % 8 == 0. */
int saspa_conv2d_igemm_bf16(const void* x0, int ldx0, int c0, const void* x1, int ldx1, int c1, int n, int h, int w,
                            const void* weight, int ksize, void* out, int ldo, int cout, const saspa_epilogue* ep_host,
                            cudaStream_t stream);

/* The same implicit GEMM with stride in {1,2} and an explicit top/left zero padding `pad` (0 .. ksize/2): x [n,ih,iw,c] -> out
 * [n,oh,ow,cout], out(y,x) = sum_taps w(ky,kx) . x(y*stride + ky - pad, x*stride + kx - pad); taps outside the input read zeros
 * (TMA out-of-bounds fill), which also supplies any bottom/right padding -- diffusers' Downsample2D (stride 2, padding 1;
 * UNet/ControlNet, models/downsampling.py) and the VAE encoder's asymmetric (0,1,0,1) pad + stride-2 conv need no im2col buffer. */
int saspa_conv2d_igemm_strided_bf16(const void* x0, int ldx0, int c0, const void* x1, int ldx1, int c1, int n, int ih, int iw,
                                    const void* weight, int ksize, int stride, int pad, int oh, int ow, void* out, int ldo, int cout,
                                    const saspa_epilogue* ep_host, cudaStream_t stream);



/* General im2col for the few strided / odd-channel convolutions (conv_in Cin=4, ControlNet cond embedding,
 * stride-2 downsamplers, ResNet stems): x bf16 NHWC [n,h,w,c] (pixel stride ldx) ->
 * cols bf16 [n*oh*ow, kpad], kpad >= kh*kw*c, zero padded; (kh,kw,c) ordering. */
int saspa_im2col_bf16(const void* x, int ldx, int n, int h, int w, int c, int kh, int kw, int stride, int pad_top, int pad_left,
                      int oh, int ow, void* cols, int kpad, cudaStream_t stream);

/* ------------------------------------------------------------------------------------------
 * Normalisation / elementwise (HBM-bound, vectorised, warp-shuffle reductions).
 * ------------------------------------------------------------------------------------------ */
/* GroupNorm over NHWC: x [n, hw, c] (pixel stride ldx), groups | c, groups <= 64; y = act((x-mean)*rstd*gamma+beta).
 * act: SASPA_ACT_NONE or SASPA_ACT_SILU.  One launch; results are deterministic and independent of which other
 * images share the batch (fixed-order reductions, no float atomics).  stats_ws: saspa_groupnorm_workspace_bytes(n, hw,
 * groups) bytes, 256-byte aligned, contents irrelevant on entry. */
size_t saspa_groupnorm_workspace_bytes(int n, int hw, int groups);
int saspa_groupnorm_nhwc_bf16(const void* x, int ldx, int n, int hw, int c, int groups, float eps, const float* gamma,
                              const float* beta, int act, void* y, int ldy, void* stats_ws, size_t ws_bytes, cudaStream_t stream);
/* LayerNorm over the last dim: x [rows, c] (row stride ldx) -> y bf16 [rows, c] (row stride ldy). */
int saspa_layernorm_bf16(const void* x, int ldx, int rows, int c, float eps, const float* gamma, const float* beta, void* y,
                         int ldy, cudaStream_t stream);
/* y = act(x) elementwise over n bf16 values. */
int saspa_act_bf16(const void* x, void* y, size_t count, int act, cudaStream_t stream);
/* y[r, :] = a[r, :] + b[r, :] (bf16, row strides) */
int saspa_add_bf16(const void* a, int lda, const void* b, int ldb, void* y, int ldy, int rows, int c, cudaStream_t stream);
/* nearest-neighbour x2 upsample NHWC: [n,h,w,c] -> [n,2h,2w,c] (diffusers Upsample2D) */
int saspa_upsample_nearest2x_bf16(const void* x, int n, int h, int w, int c, void* y, cudaStream_t stream);
/* fp32 <-> bf16 casts with optional NCHW<->NHWC transposition of [n,c,h,w] fp32 latents */
int saspa_nchw_f32_to_nhwc_bf16(const float* x, int n, int c, int h, int w, void* y, int ldy, float scale, cudaStream_t stream);
int saspa_nhwc_to_nchw_f32(const void* x, int ldx, int x_is_fp32, int n, int c, int h, int w, float* y, cudaStream_t stream);
/* max / average pooling NHWC (filter networks) */
int saspa_pool2d_nhwc_bf16(const void* x, int n, int h, int w, int c, int k, int stride, int pad, int is_max, void* y, int oh,
                           int ow, cudaStream_t stream);

/* ------------------------------------------------------------------------------------------
 * Attention: O = softmax(Q K^T * scale) V per (batch, head).  Q [b, tq, heads*d] with row stride ldq
 * (so a fused QKV buffer works), K/V [b, tkv, heads*d] (ldk, ldv), O [b, tq, heads*d] (ldo).
 * d % 8 == 0, d <= 160 (SD v1.5: 40/80/160; SDXL: 64) or d == 512 (VAE mid block, heads == 1).
 * Replaces F.scaled_dot_product_attention under diffusers' AttnProcessor2_0.
 * ------------------------------------------------------------------------------------------ */
int saspa_attention_bf16(const void* q, int ldq, const void* k, int ldk, const void* v, int ldv, void* o, int ldo, int batch,
                         int heads, int tq, int tkv, int d, float scale, int causal /* CLIP text tower */, cudaStream_t stream);

/* Row softmax y = softmax(x * scale) (bf16, fp32 math) and batched 2-D transpose: the d = 512 single-head VAE
 * mid-block attention (diffusers models/autoencoders/vae.py UNetMidBlock2D) runs GEMM -> softmax -> GEMM. */
int saspa_softmax_rows_bf16(const void* x, int ldx, void* y, int ldy, long long rows, int cols, float scale, cudaStream_t stream);
int saspa_transpose_bf16(const void* x, int ldx, long long batch_stride_x, void* y, int ldy, long long batch_stride_y, int batch, int rows,
                         int cols, cudaStream_t stream);

/* ------------------------------------------------------------------------------------------
 * Denoising-loop glue.
 * ------------------------------------------------------------------------------------------ */
/* Sinusoidal timestep projection (diffusers Timesteps, flip_sin_to_cos=True, freq_shift=0):
 * out bf16 [rows, dim] = [cos(t*f_i), sin(t*f_i)], f_i = exp(-ln(1e4) * i / (dim/2)). */
int saspa_timestep_sinusoid_bf16(const float* t, int rows, int dim, int flip_sin_to_cos, float freq_shift, void* out,
                                 cudaStream_t stream);
/* Classifier-free-guidance combine fused with the scheduler update (diffusers pipeline __call__ +
 * DDIMScheduler/UniPCMultistepScheduler/PNDMScheduler.step).  All buffers fp32 [count].
 *   e      = eps_uncond + guidance * (eps_cond - eps_uncond)      (eps_uncond NULL: e = eps_cond)
 *   in[0]  = x (current sample), in[1] = e, in[2..n_in) = history buffers supplied by the caller
 *   out[j] = sum_i coef[j*n_in + i] * in[i]     for j < n_out   (coefficients by value, host side)
 * n_in <= 8, n_out <= 4.  Outputs may alias inputs elementwise. */
typedef struct saspa_lincomb {
  const float* in[8];
  float* out[4];
  float coef[32];
  int n_in, n_out;
} saspa_lincomb;
int saspa_cfg_sched_step(const float* eps_uncond, const float* eps_cond, float guidance, const saspa_lincomb* lc_host, size_t count,
                         cudaStream_t stream);
/* Same, plus ONE clamped term (DDIMScheduler clip_sample=True -- the DDIM defaults SD-XL-turbo's scheduler config inherits
 * through DDIMScheduler.from_config at run_aug/run_aug.py:228: pred_original_sample.clamp(-clip_sample_range, clip_sample_range)):
 *   c      = clamp(sum_i pre[i] * in[i], -clip_range, clip_range)
 *   out[j] = sum_i coef[j*n_in + i] * in[i] + post[j] * c
 * pre_host [n_in], post_host [n_out]; clip_range <= 0 disables the term (pre/post may then be NULL). */
int saspa_cfg_sched_step_clip(const float* eps_uncond, const float* eps_cond, float guidance, const saspa_lincomb* lc_host,
                              const float* pre_host, const float* post_host, float clip_range, size_t count, cudaStream_t stream);
/* img2img start (diffusers prepare_latents of StableDiffusionControlNetImg2ImgPipeline): posterior sample of the VAE
 * encoder moments (fp32 NHWC [n,h,w,2*lc]: mean | logvar) times scaling_factor, then scheduler.add_noise:
 *   z0 = (mean + exp(0.5*clamp(logvar,-30,20)) * noise_posterior) * scaling;  latents = alpha*z0 + sigma*noise_diffusion
 * noises / latents / z0_out (optional) are fp32 NCHW [n,lc,h,w]. */
int saspa_vae_sample_add_noise(const float* moments, const float* noise_posterior, const float* noise_diffusion, float scaling,
                               float alpha, float sigma, int n, int h, int w, int latent_channels, float* latents, float* z0_out,
                               cudaStream_t stream);
/* VAE postprocess (diffusers VaeImageProcessor.postprocess): u8 = round_half_even(clamp(x/2+0.5,0,1)*255).
 * x bf16 or fp32 NHWC [pixels, ldx] (first 3 channels) -> u8 [pixels,3]. */
int saspa_vae_quantize_u8(const void* x, int ldx, int x_is_fp32, size_t pixels, uint8_t* out, cudaStream_t stream);

/* ------------------------------------------------------------------------------------------
 * Filter heads (all_utils/utils.py:357-365 and :169-177).
 * ------------------------------------------------------------------------------------------ */
/* WS-DAN bilinear attention pooling head (fgvc/models/cal.py:53-86 BAP, :184-213 forward):
 * feat bf16 [n, hw, c] , att bf16 [n, hw, m]  ->  fm fp32 [n, m*c] = l2norm(sign(p)*sqrt(|p|+1e-6)) * 100,
 * p[i,j] = sum_hw att[hw,i]*feat[hw,j] / hw. */
int saspa_bap_head(const void* feat, int ldf, const void* att, int lda, int n, int hw, int c, int m, float* fm, float* sq_ws /* [n] */,
                   cudaStream_t stream);
/* logits fp32 [n, classes] = fm [n, k] . W[classes, k]^T + b  (skinny, HBM-bound on the fp32 W; the reference runs
 * this classifier in fp32, all_utils/dataset_utils.py:111) */
int saspa_fc_f32(const float* x, const float* w, const float* bias, int n, int k, int classes, float* logits, cudaStream_t stream);
/* keep[i] = (label[i] in top-k(logits[i, :]))  -- torch.topk tie order: lower index wins */
int saspa_topk_contains(const float* logits, int n, int classes, const int32_t* label, int k, uint8_t* keep, float* margin,
                        cudaStream_t stream);
/* prob[i] = softmax(logits[i, :])[idx[i]], max_logit[i] = max_j logits[i, j], argmax[i] = its first index (each output optional): the
 * confidence tests of the reference's optional filters (all_utils/utils.py:186-191 clip_filtering, :370-376
 * filter_confidence_higher_than, :411-418 alia_conf_filtering).  idx values must lie in [0, classes). */
int saspa_softmax_at_f32(const float* logits, int n, int classes, const int32_t* idx, float* prob, float* max_logit, int32_t* argmax,
                         cudaStream_t stream);
/* CLIP cosine scoring (all_utils/utils.py:152-166): logits = scale * norm(img) @ norm(txt)^T; argmax over prompts.
 * img fp32 [n, d], txt fp32 [p, d] -> logits fp32 [n, p] (optional), argmax int32 [n]. */
int saspa_clip_score_argmax(const float* img, const float* txt, int n, int p, int d, float logit_scale, float* logits,
                            int32_t* argmax, cudaStream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* SASPA_B200_H_ */
