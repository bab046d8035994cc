/* Kernel-selection overrides for tests and A/B timing.  NOT part of the product ABI (include/saspa_b200.h): the library selects kernels
 * from the problem shape alone and a product caller never needs these.  Each hook sets one process-wide atomic integer (0 = automatic
 * selection, the state the library starts in) and returns the previous value; a negative argument only queries.  They exist so that a
 * test can run the same shape through every kernel variant and a tool can time one variant against another. */
#ifndef SASPA_B200_TUNING_HOOKS_H_
#define SASPA_B200_TUNING_HOOKS_H_
#ifdef __cplusplus
extern "C" {
#endif
/* 3x3 implicit-GEMM main loop: 0 auto, 1 one TMA box per tap, 2 halo tile only (error when the shape is not eligible) */
int saspa_conv_impl(int impl);
/* CTAs per output tile of the tcgen05 GEMM / conv kernel: 0 auto, 1 single-CTA tiles, 2 cta_group::2 pairs */
int saspa_gemm_force_ctas(int ctas);
/* GEMM row tiles walked last to first (1) instead of first to last (0): an A operand written front to back by the previous launch is
 * then read starting with the rows still in L2.  Same results, different tile order. */
int saspa_gemm_reverse_m(int on);
/* N tile width of the non-GEGLU GEMM kernels: 0 auto, or 32 / 64 / 128 / 160 / 256 */
int saspa_gemm_force_bn(int bn);
/* GroupNorm: 0 auto, 1 two-pass kernel only */
int saspa_groupnorm_impl(int impl);
/* attention: 0 auto, 1 mma.sync flash kernel, 2 tcgen05 kernels only, 3 the mma.sync K/V-resident cross-attention kernel */
int saspa_attention_impl(int impl);
#ifdef __cplusplus
}
#endif
#endif
