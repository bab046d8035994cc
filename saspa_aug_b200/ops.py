"""Thin torch-tensor wrappers over the C ABI (include/saspa_b200.h).

torch is plumbing only: it owns device memory and the current stream.  Every function
here enqueues hand-written sm_100a kernels from libsaspa_b200.so on
``torch.cuda.current_stream()`` and raises ``SaspaError`` on failure -- there is no
eager/PyTorch fallback.  ``LAUNCHES`` counts kernel launches issued through this module
(bench.py reports it as ``gpu_launches``)."""
from __future__ import annotations

import ctypes
from typing import Optional

import torch

from . import _lib
from ._lib import ACT_GEGLU, ACT_GELU, ACT_NONE, ACT_QUICKGELU, ACT_RELU, ACT_SILU, Epilogue, LinComb, SaspaError, check

LAUNCHES = 0  # kernels launched via this module (host-side count)
PROFILE = None  # set to a list to record (kind, flops, start_event, end_event) around every tcgen05 launch (bench.py roofline leg)

BF16 = torch.bfloat16


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else t.data_ptr()


def _need_cuda(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise SaspaError("saspa_aug_b200.ops: tensors must live on a CUDA device (no CPU fallback exists)")


def _count(n: int = 1):
    global LAUNCHES
    LAUNCHES += n


def _prof_begin():
    if PROFILE is None:
        return None
    ev = torch.cuda.Event(enable_timing=True)
    ev.record()
    return ev


def _prof_end(ev, kind: str, flops: float, shape=None):
    if ev is None:
        return
    end = torch.cuda.Event(enable_timing=True)
    end.record()
    PROFILE.append((kind, flops, ev, end, shape))


# ------------------------------------------------------------------------------------------------
# Canny
# ------------------------------------------------------------------------------------------------
def canny(img: torch.Tensor, low: int, high: int, out_channels: int = 1, want_ctrl: bool = False):
    """img u8 [n,h,w,c] (c in {1,3}) -> (edges u8 [n,h,w] or [n,h,w,3], ctrl bf16 [n,h,w,3] | None)."""
    _need_cuda(img)
    assert img.dtype == torch.uint8 and img.dim() == 4 and img.is_contiguous()
    n, h, w, c = img.shape
    lib = _lib.load()
    ws_bytes = lib.saspa_canny_workspace_bytes(n, h, w)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=img.device)
    shape = (n, h, w) if out_channels == 1 else (n, h, w, out_channels)
    out = torch.empty(shape, dtype=torch.uint8, device=img.device)
    ctrl = torch.empty((n, h, w, 3), dtype=BF16, device=img.device) if want_ctrl else None
    check(
        lib.saspa_canny_u8(_ptr(img), n, h, w, c, int(low), int(high), _ptr(out), out_channels, _ptr(ctrl), _ptr(ws), ws_bytes, _stream()),
        "saspa_canny_u8",
    )
    _count(3)
    return out, ctrl


# ------------------------------------------------------------------------------------------------
# GEMM / conv
# ------------------------------------------------------------------------------------------------
def make_epilogue(bias=None, row_bias=None, rows_per_group=1, act=ACT_NONE, alpha=1.0, residual=None, ld_res=0, beta=1.0, out_fp32=False,
                  act_after_residual=False, row_stats_out=None, ln_stats=None, ln_colsum=None, ln_eps=1e-5):
    ep = Epilogue()
    ep.row_stats_out = _ptr(row_stats_out)
    ep.row_stats_slots = int(row_stats_out.shape[1]) if row_stats_out is not None else 0
    ep.ln_stats = _ptr(ln_stats)
    ep.ln_slots = int(ln_stats.shape[1]) if ln_stats is not None else 0
    ep.ln_colsum = _ptr(ln_colsum)
    ep.ln_eps = float(ln_eps)
    ep.bias = _ptr(bias)
    ep.row_bias = _ptr(row_bias)
    ep.rows_per_group = int(rows_per_group)
    ep.ld_row_bias = int(row_bias.stride(0)) if row_bias is not None else 0
    ep.act_after_residual = 1 if act_after_residual else 0
    ep.act = int(act)
    ep.alpha = float(alpha)
    ep.residual = _ptr(residual)
    ep.ld_res = int(ld_res) if residual is not None else 0
    ep.beta = float(beta)
    ep.out_fp32 = 1 if out_fp32 else 0
    return ep


def row_stats_slots(n: int) -> int:
    """Partial-statistics slots per row a GEMM with ``n`` output columns writes (see ``gemm(row_stats=...)``)."""
    return int(_lib.load().saspa_gemm_row_stats_slots(int(n)))


def gemm(a: torch.Tensor, b: torch.Tensor, out: Optional[torch.Tensor] = None, *, bias=None, row_bias=None, rows_per_group=1,
         act=ACT_NONE, alpha=1.0, residual=None, beta=1.0, out_fp32=False, act_after_residual=False, row_stats: bool = False,
         ln_stats: Optional[torch.Tensor] = None, ln_colsum: Optional[torch.Tensor] = None, ln_eps: float = 1e-5):
    """out[M,N'] = epilogue(a[M,K] @ b[N,K]^T); a, b bf16 2-D views with unit inner stride.
    row_stats=True additionally returns fp32 [M, slots, 2] partial (sum, sum of squares) of every OUTPUT row -> (out, stats);
    ln_stats (such a tensor for the rows of ``a``) + ln_colsum fold a LayerNorm of ``a`` into this GEMM (b = W * gamma,
    bias = b + W . beta, ln_colsum = b.float().sum(1))."""
    _need_cuda(a, b)
    assert a.dtype == BF16 and b.dtype == BF16 and a.dim() == 2 and b.dim() == 2
    assert a.stride(1) == 1 and b.stride(1) == 1
    M, K = a.shape
    N, Kb = b.shape
    assert K == Kb, (a.shape, b.shape)
    n_out = N // 2 if act == ACT_GEGLU else N
    if out is None:
        out = torch.empty((M, n_out), dtype=torch.float32 if out_fp32 else BF16, device=a.device)
    assert out.stride(1) == 1 and out.shape == (M, n_out)
    if residual is not None:
        assert residual.dtype == BF16 and residual.stride(1) == 1 and residual.shape == (M, n_out)
    stats = torch.empty((M, row_stats_slots(n_out), 2), dtype=torch.float32, device=a.device) if row_stats else None
    if ln_stats is not None:
        assert ln_stats.dtype == torch.float32 and ln_stats.is_contiguous() and ln_stats.shape[0] == M and ln_stats.shape[2] == 2
        assert ln_colsum is not None and ln_colsum.dtype == torch.float32 and ln_colsum.numel() == N
    ep = make_epilogue(bias, row_bias, rows_per_group, act, alpha, residual, residual.stride(0) if residual is not None else 0, beta,
                       out.dtype == torch.float32, act_after_residual, stats, ln_stats, ln_colsum, ln_eps)
    ev = _prof_begin()
    check(
        _lib.load().saspa_gemm_bf16(_ptr(a), a.stride(0), _ptr(b), b.stride(0), _ptr(out), out.stride(0), M, N, K, ctypes.byref(ep), _stream()),
        "saspa_gemm_bf16",
    )
    _prof_end(ev, "gemm", 2.0 * M * N * K, (M, N, K, int(act), residual is not None))
    _count()
    return (out, stats) if row_stats else out


def conv2d_igemm(x: torch.Tensor, weight: torch.Tensor, ksize: int, out: Optional[torch.Tensor] = None, *, x1: Optional[torch.Tensor] = None,
                 bias=None, row_bias=None, act=ACT_NONE, alpha=1.0, residual=None, beta=1.0, out_fp32=False, act_after_residual=False,
                 stride: int = 1, pad: Optional[int] = None, out_hw: Optional[tuple] = None) -> torch.Tensor:
    """Implicit-GEMM conv on NHWC bf16 views [n,h,w,c] (channel stride 1, dense n/h/w strides); weight bf16 [cout, ksize*ksize*(c0+c1)].
    Default: stride 1, "same" padding.  stride 2 / explicit top-left ``pad`` / ``out_hw`` = (oh, ow): taps outside the input read zeros."""
    _need_cuda(x, weight)
    n, h, w, c0 = x.shape
    assert x.dtype == BF16 and x.stride(3) == 1 and x.stride(1) == w * x.stride(2) and x.stride(0) == h * x.stride(1)
    c1 = 0
    if x1 is not None:
        assert x1.shape[:3] == x.shape[:3] and x1.stride(3) == 1 and x1.stride(1) == w * x1.stride(2) and x1.stride(0) == h * x1.stride(1)
        c1 = x1.shape[3]
    cout = weight.shape[0]
    assert weight.dtype == BF16 and weight.is_contiguous() and weight.shape[1] == ksize * ksize * (c0 + c1)
    pad = ksize // 2 if pad is None else pad
    oh, ow = out_hw if out_hw is not None else ((h + 2 * pad - ksize) // stride + 1, (w + 2 * pad - ksize) // stride + 1)
    if out is None:
        out = torch.empty((n, oh, ow, cout), dtype=torch.float32 if out_fp32 else BF16, device=x.device)
    assert out.shape[1:3] == (oh, ow) and out.stride(3) == 1 and out.stride(1) == ow * out.stride(2) and out.stride(0) == oh * out.stride(1)
    ld_res = 0
    if residual is not None:
        assert residual.dtype == BF16 and residual.stride(3) == 1 and residual.stride(1) == ow * residual.stride(2)
        ld_res = residual.stride(2)
    ep = make_epilogue(bias, row_bias, oh * ow, act, alpha, residual, ld_res, beta, out.dtype == torch.float32, act_after_residual)
    ev = _prof_begin()
    check(
        _lib.load().saspa_conv2d_igemm_strided_bf16(_ptr(x), x.stride(2), c0, _ptr(x1), x1.stride(2) if x1 is not None else 0, c1, n, h, w,
                                                    _ptr(weight), ksize, stride, pad, oh, ow, _ptr(out), out.stride(2), cout, ctypes.byref(ep),
                                                    _stream()),
        "saspa_conv2d_igemm_strided_bf16",
    )
    _prof_end(ev, "conv", 2.0 * n * oh * ow * cout * ksize * ksize * (c0 + c1), (n, oh, ow, c0 + c1, cout, ksize) if stride == 1 else (n, oh, ow, c0 + c1, cout, ksize, stride))
    _count()
    return out


def conv3x3_small_supported(cin: int, cout: int, stride: int, pad: int) -> bool:
    return bool(_lib.load().saspa_conv3x3_small_supported(int(cin), int(cout), int(stride), int(pad)))


def conv3x3_small(x: torch.Tensor, weight: torch.Tensor, bias=None, act=ACT_NONE, stride: int = 1, out: Optional[torch.Tensor] = None,
                  residual: Optional[torch.Tensor] = None) -> torch.Tensor:
    """3x3 / padding 1 convolution for cin <= 32 on NHWC bf16 [n,h,w,cin] (dense pixel strides); weight bf16 [cout, kpad] in
    (ky, kx, cin) order.  Bias (+ residual [n,oh,ow,cout]) + activation in the epilogue, bf16 out [n,oh,ow,cout] (any pixel stride).
    cout > 96 runs as slices of 64 output channels."""
    _need_cuda(x, weight)
    n, h, w, cin = x.shape
    assert x.dtype == BF16 and x.stride(3) == 1 and x.stride(1) == w * x.stride(2) and x.stride(0) == h * x.stride(1)
    cout, kpad = weight.shape
    assert weight.dtype == BF16 and weight.is_contiguous() and kpad >= 9 * cin
    oh, ow = (h + 2 - 3) // stride + 1, (w + 2 - 3) // stride + 1
    if out is None:
        out = torch.empty((n, oh, ow, cout), dtype=BF16, device=x.device)
    for t in (out, residual):
        assert t is None or (t.dtype == BF16 and t.shape == (n, oh, ow, cout) and t.stride(3) == 1 and t.stride(1) == ow * t.stride(2) and t.stride(0) == oh * t.stride(1))
    lib = _lib.load()
    step = 128 if cout <= 96 else 64  # wide layers: 64-channel slices keep two CTAs per SM (the 128-channel kernel holds 179 registers)
    for c0 in range(0, cout, step):
        c1 = min(cout, c0 + step)
        check(lib.saspa_conv3x3_small_bf16(_ptr(x), x.stride(2), cin, n, h, w, _ptr(weight[c0:c1]), kpad, _ptr(bias[c0:c1]) if bias is not None else None, int(act),
                                           _ptr(residual[..., c0:c1]) if residual is not None else None, residual.stride(2) if residual is not None else 0,
                                           int(stride), 1, _ptr(out[..., c0:c1]), out.stride(2), c1 - c0, oh, ow, _stream()), "saspa_conv3x3_small_bf16")
        _count()
    return out


def im2col(x: torch.Tensor, kh: int, kw: int, stride: int, pad_top: int, pad_left: int, oh: int, ow: int, kpad: int) -> torch.Tensor:
    _need_cuda(x)
    n, h, w, c = x.shape
    assert x.dtype == BF16 and x.stride(3) == 1 and x.stride(1) == w * x.stride(2) and x.stride(0) == h * x.stride(1)
    cols = torch.empty((n * oh * ow, kpad), dtype=BF16, device=x.device)
    check(
        _lib.load().saspa_im2col_bf16(_ptr(x), x.stride(2), n, h, w, c, kh, kw, stride, pad_top, pad_left, oh, ow, _ptr(cols), kpad, _stream()),
        "saspa_im2col_bf16",
    )
    _count()
    return cols


# ------------------------------------------------------------------------------------------------
# normalisation / elementwise
# ------------------------------------------------------------------------------------------------
def groupnorm(x: torch.Tensor, groups: int, eps: float, gamma, beta, act=ACT_NONE, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """x bf16 [n, hw, c] view (channel stride 1, dense hw) -> same shape."""
    _need_cuda(x)
    n, hw, c = x.shape
    assert x.dtype == BF16 and x.stride(2) == 1 and x.stride(0) == hw * x.stride(1)
    if out is None:
        out = torch.empty((n, hw, c), dtype=BF16, device=x.device)
    assert out.stride(2) == 1 and out.stride(0) == hw * out.stride(1)
    lib = _lib.load()
    ws_bytes = lib.saspa_groupnorm_workspace_bytes(n, hw, groups)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=x.device)
    check(
        lib.saspa_groupnorm_nhwc_bf16(_ptr(x), x.stride(1), n, hw, c, groups, float(eps), _ptr(gamma), _ptr(beta), int(act), _ptr(out),
                                      out.stride(1), _ptr(ws), ws_bytes, _stream()),
        "saspa_groupnorm_nhwc_bf16",
    )
    _count(1)
    return out


def layernorm(x: torch.Tensor, eps: float, gamma, beta, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    _need_cuda(x)
    rows, c = x.shape
    assert x.dtype == BF16 and x.stride(1) == 1
    if out is None:
        out = torch.empty((rows, c), dtype=BF16, device=x.device)
    check(_lib.load().saspa_layernorm_bf16(_ptr(x), x.stride(0), rows, c, float(eps), _ptr(gamma), _ptr(beta), _ptr(out), out.stride(0), _stream()),
          "saspa_layernorm_bf16")
    _count()
    return out


def act(x: torch.Tensor, kind: int, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    _need_cuda(x)
    assert x.dtype == BF16 and x.is_contiguous()
    if out is None:
        out = torch.empty_like(x)
    check(_lib.load().saspa_act_bf16(_ptr(x), _ptr(out), x.numel(), int(kind), _stream()), "saspa_act_bf16")
    _count()
    return out


def add(a: torch.Tensor, b: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    _need_cuda(a, b)
    rows, c = a.shape
    assert a.dtype == BF16 and b.dtype == BF16 and a.stride(1) == 1 and b.stride(1) == 1 and b.shape == a.shape
    if out is None:
        out = torch.empty((rows, c), dtype=BF16, device=a.device)
    check(_lib.load().saspa_add_bf16(_ptr(a), a.stride(0), _ptr(b), b.stride(0), _ptr(out), out.stride(0), rows, c, _stream()), "saspa_add_bf16")
    _count()
    return out


def upsample_nearest2x(x: torch.Tensor) -> torch.Tensor:
    _need_cuda(x)
    n, h, w, c = x.shape
    assert x.dtype == BF16 and x.is_contiguous()
    y = torch.empty((n, 2 * h, 2 * w, c), dtype=BF16, device=x.device)
    check(_lib.load().saspa_upsample_nearest2x_bf16(_ptr(x), n, h, w, c, _ptr(y), _stream()), "saspa_upsample_nearest2x_bf16")
    _count()
    return y


def nchw_f32_to_nhwc_bf16(x: torch.Tensor, scale: float = 1.0, out: Optional[torch.Tensor] = None, pad_c: Optional[int] = None) -> torch.Tensor:
    _need_cuda(x)
    n, c, h, w = x.shape
    assert x.dtype == torch.float32 and x.is_contiguous()
    if out is None:
        cc = pad_c or c
        out = torch.zeros((n, h, w, cc), dtype=BF16, device=x.device) if cc != c else torch.empty((n, h, w, c), dtype=BF16, device=x.device)
    check(_lib.load().saspa_nchw_f32_to_nhwc_bf16(_ptr(x), n, c, h, w, _ptr(out), out.stride(2), float(scale), _stream()), "saspa_nchw_f32_to_nhwc_bf16")
    _count()
    return out


def nhwc_to_nchw_f32(x: torch.Tensor, c: Optional[int] = None) -> torch.Tensor:
    _need_cuda(x)
    n, h, w, cc = x.shape
    c = c or cc
    assert x.stride(3) == 1 and x.stride(1) == w * x.stride(2) and x.stride(0) == h * x.stride(1)
    y = torch.empty((n, c, h, w), dtype=torch.float32, device=x.device)
    check(_lib.load().saspa_nhwc_to_nchw_f32(_ptr(x), x.stride(2), 1 if x.dtype == torch.float32 else 0, n, c, h, w, _ptr(y), _stream()),
          "saspa_nhwc_to_nchw_f32")
    _count()
    return y


def pool2d(x: torch.Tensor, k: int, stride: int, pad: int, is_max: bool) -> torch.Tensor:
    _need_cuda(x)
    n, h, w, c = x.shape
    assert x.dtype == BF16 and x.is_contiguous()
    oh = (h + 2 * pad - k) // stride + 1
    ow = (w + 2 * pad - k) // stride + 1
    y = torch.empty((n, oh, ow, c), dtype=BF16, device=x.device)
    check(_lib.load().saspa_pool2d_nhwc_bf16(_ptr(x), n, h, w, c, k, stride, pad, 1 if is_max else 0, _ptr(y), oh, ow, _stream()), "saspa_pool2d_nhwc_bf16")
    _count()
    return y


# ------------------------------------------------------------------------------------------------
# attention
# ------------------------------------------------------------------------------------------------
def attention(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, heads: int, scale: Optional[float] = None, out: Optional[torch.Tensor] = None,
              causal: bool = False):
    """q [b,tq,heads*d], k/v [b,tkv,heads*d] bf16 views (inner stride 1, dense token stride) -> [b,tq,heads*d]."""
    _need_cuda(q, k, v)
    b, tq, hd = q.shape
    tkv = k.shape[1]
    d = hd // heads
    assert q.stride(2) == 1 and k.stride(2) == 1 and v.stride(2) == 1
    assert q.stride(0) == tq * q.stride(1) and k.stride(0) == tkv * k.stride(1) and v.stride(0) == tkv * v.stride(1)
    if scale is None:
        scale = d ** -0.5
    if out is None:
        out = torch.empty((b, tq, hd), dtype=BF16, device=q.device)
    assert out.stride(2) == 1 and out.stride(0) == tq * out.stride(1)
    ev = _prof_begin()
    check(
        _lib.load().saspa_attention_bf16(_ptr(q), q.stride(1), _ptr(k), k.stride(1), _ptr(v), v.stride(1), _ptr(out), out.stride(1), b, heads, tq,
                                         tkv, d, float(scale), 1 if causal else 0, _stream()),
        "saspa_attention_bf16",
    )
    _prof_end(ev, "attention", 4.0 * b * heads * tq * tkv * d, (b, heads, tq, tkv, d))
    _count()
    return out


def softmax_rows(x: torch.Tensor, scale: float = 1.0, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    _need_cuda(x)
    rows, cols = x.shape
    assert x.dtype == BF16 and x.stride(1) == 1
    if out is None:
        out = torch.empty((rows, cols), dtype=BF16, device=x.device)
    check(_lib.load().saspa_softmax_rows_bf16(_ptr(x), x.stride(0), _ptr(out), out.stride(0), rows, cols, float(scale), _stream()), "saspa_softmax_rows_bf16")
    _count()
    return out


def transpose(x: torch.Tensor) -> torch.Tensor:
    """x bf16 [batch, rows, cols] (inner stride 1) -> contiguous [batch, cols, rows]."""
    _need_cuda(x)
    b, rows, cols = x.shape
    assert x.dtype == BF16 and x.stride(2) == 1
    y = torch.empty((b, cols, rows), dtype=BF16, device=x.device)
    check(_lib.load().saspa_transpose_bf16(_ptr(x), x.stride(1), x.stride(0), _ptr(y), rows, cols * rows, b, rows, cols, _stream()), "saspa_transpose_bf16")
    _count()
    return y


# ------------------------------------------------------------------------------------------------
# denoising-loop glue
# ------------------------------------------------------------------------------------------------
def timestep_sinusoid(t: torch.Tensor, dim: int, flip_sin_to_cos: bool = True, freq_shift: float = 0.0) -> torch.Tensor:
    _need_cuda(t)
    assert t.dtype == torch.float32 and t.dim() == 1
    out = torch.empty((t.shape[0], dim), dtype=BF16, device=t.device)
    check(_lib.load().saspa_timestep_sinusoid_bf16(_ptr(t), t.shape[0], dim, 1 if flip_sin_to_cos else 0, float(freq_shift), _ptr(out), _stream()),
          "saspa_timestep_sinusoid_bf16")
    _count()
    return out


def cfg_sched_step(eps_uncond, eps_cond, guidance: float, inputs, outputs, coef, clip_pre=None, clip_post=None, clip_range: float = 0.0) -> None:
    """outputs[j] = sum_i coef[j][i] * in_i  (+ clip_post[j] * clamp(sum_i clip_pre[i] * in_i, +-clip_range) when clip_range > 0)
    with in = [x, cfg(eps), *history]; ``inputs`` = [x, None, hist...] (slot 1 is the CFG-combined eps, produced inside the kernel)."""
    _need_cuda(eps_cond)
    lc = LinComb()
    n_in, n_out = len(inputs), len(outputs)
    assert 2 <= n_in <= 8 and 1 <= n_out <= 4 and len(coef) == n_out and all(len(r) == n_in for r in coef)
    count = eps_cond.numel()
    for i, t in enumerate(inputs):
        if t is not None:
            assert t.dtype == torch.float32 and t.is_contiguous() and t.numel() == count
        lc.inp[i] = _ptr(t)
    for j, t in enumerate(outputs):
        assert t.dtype == torch.float32 and t.is_contiguous() and t.numel() == count
        lc.out[j] = _ptr(t)
    for j in range(n_out):
        for i in range(n_in):
            lc.coef[j * n_in + i] = float(coef[j][i])
    lc.n_in, lc.n_out = n_in, n_out
    assert eps_cond.dtype == torch.float32 and eps_cond.is_contiguous()
    if eps_uncond is not None:
        assert eps_uncond.dtype == torch.float32 and eps_uncond.is_contiguous() and eps_uncond.numel() == count
    if clip_range > 0.0:
        assert len(clip_pre) == n_in and len(clip_post) == n_out
        pre = (ctypes.c_float * n_in)(*[float(v) for v in clip_pre])
        post = (ctypes.c_float * n_out)(*[float(v) for v in clip_post])
        check(_lib.load().saspa_cfg_sched_step_clip(_ptr(eps_uncond), _ptr(eps_cond), float(guidance), ctypes.byref(lc), pre, post, float(clip_range), count,
                                                    _stream()), "saspa_cfg_sched_step_clip")
    else:
        check(_lib.load().saspa_cfg_sched_step(_ptr(eps_uncond), _ptr(eps_cond), float(guidance), ctypes.byref(lc), count, _stream()), "saspa_cfg_sched_step")
    _count()


def vae_sample_add_noise(moments: torch.Tensor, noise_post, noise_diff, scaling: float, alpha: float, sigma: float):
    """moments fp32 NHWC [n,h,w,2*lc] -> (latents, z0) fp32 NCHW [n,lc,h,w]."""
    _need_cuda(moments)
    n, h, w, c2 = moments.shape
    lc = c2 // 2
    assert moments.dtype == torch.float32 and moments.is_contiguous()
    lat = torch.empty((n, lc, h, w), dtype=torch.float32, device=moments.device)
    z0 = torch.empty_like(lat)
    for t in (noise_post, noise_diff):
        assert t is None or (t.dtype == torch.float32 and t.is_contiguous() and t.shape == lat.shape)
    check(_lib.load().saspa_vae_sample_add_noise(_ptr(moments), _ptr(noise_post), _ptr(noise_diff), float(scaling), float(alpha), float(sigma), n, h, w, lc,
                                                 _ptr(lat), _ptr(z0), _stream()), "saspa_vae_sample_add_noise")
    _count()
    return lat, z0


def vae_quantize_u8(x: torch.Tensor) -> torch.Tensor:
    """x [n,h,w,c>=3] (bf16|fp32, channel stride 1) -> u8 [n,h,w,3]."""
    _need_cuda(x)
    n, h, w, c = x.shape
    assert x.stride(3) == 1 and x.stride(1) == w * x.stride(2) and x.stride(0) == h * x.stride(1)
    out = torch.empty((n, h, w, 3), dtype=torch.uint8, device=x.device)
    check(_lib.load().saspa_vae_quantize_u8(_ptr(x), x.stride(2), 1 if x.dtype == torch.float32 else 0, n * h * w, _ptr(out), _stream()), "saspa_vae_quantize_u8")
    _count()
    return out


# ------------------------------------------------------------------------------------------------
# filter side
# ------------------------------------------------------------------------------------------------
_COEFF_CACHE: dict = {}


def _pil_coeffs(in_size: int, out_size: int, filt: int, device):
    key = (in_size, out_size, filt, str(device))
    if key not in _COEFF_CACHE:
        lib = _lib.load()
        ks = lib.saspa_pil_ksize(in_size, out_size, filt)
        bounds = (ctypes.c_int32 * (2 * out_size))()
        coeffs = (ctypes.c_int32 * (ks * out_size))()
        kso = ctypes.c_int(0)
        check(lib.saspa_pil_coeffs_host(in_size, out_size, filt, ctypes.byref(kso), bounds, coeffs, ks * out_size), "saspa_pil_coeffs_host")
        b = torch.tensor(list(bounds), dtype=torch.int32).to(device)
        c = torch.tensor(list(coeffs), dtype=torch.int32).to(device)
        _COEFF_CACHE[key] = (b, c, kso.value)
    return _COEFF_CACHE[key]


def resize_pil(img: torch.Tensor, out_h: int, out_w: int, filt: str) -> torch.Tensor:
    """img u8 [n,h,w,c] -> u8 [n,out_h,out_w,c]; bit-exact PIL Image.resize(bilinear|bicubic)."""
    _need_cuda(img)
    n, h, w, c = img.shape
    assert img.dtype == torch.uint8 and img.is_contiguous()
    f = {"bilinear": 0, "bicubic": 1}[filt]
    bx, cx, kx = _pil_coeffs(w, out_w, f, img.device)
    by, cy, ky = _pil_coeffs(h, out_h, f, img.device)
    tmp = torch.empty((n, h, out_w, c), dtype=torch.uint8, device=img.device)
    out = torch.empty((n, out_h, out_w, c), dtype=torch.uint8, device=img.device)
    check(_lib.load().saspa_resize_pil_u8(_ptr(img), n, h, w, c, _ptr(tmp), _ptr(out), out_h, out_w, _ptr(bx), _ptr(cx), kx, _ptr(by), _ptr(cy), ky, _stream()),
          "saspa_resize_pil_u8")
    _count(2)
    return out


def crop_normalize(img: torch.Tensor, crop_y: int, crop_x: int, crop_h: int, crop_w: int, mean, std, out_c: int = 8) -> torch.Tensor:
    _need_cuda(img)
    n, h, w, c = img.shape
    assert img.dtype == torch.uint8 and c == 3 and img.is_contiguous()
    out = torch.empty((n, crop_h, crop_w, out_c), dtype=BF16, device=img.device)
    check(_lib.load().saspa_crop_normalize_bf16(_ptr(img), n, h, w, crop_y, crop_x, crop_h, crop_w, *[float(m) for m in mean], *[float(s) for s in std],
                                                _ptr(out), out_c, _stream()), "saspa_crop_normalize_bf16")
    _count()
    return out


def bap_head(feat: torch.Tensor, att: torch.Tensor) -> torch.Tensor:
    """feat bf16 [n,hw,c], att bf16 [n,hw,m] -> fp32 [n, m*c] (sign-sqrt, L2-normalised, x100)."""
    _need_cuda(feat, att)
    n, hw, c = feat.shape
    m = att.shape[2]
    assert feat.stride(2) == 1 and att.stride(2) == 1 and feat.stride(0) == hw * feat.stride(1) and att.stride(0) == hw * att.stride(1)
    fm = torch.empty((n, m * c), dtype=torch.float32, device=feat.device)
    sq = torch.empty((n,), dtype=torch.float32, device=feat.device)
    check(_lib.load().saspa_bap_head(_ptr(feat), feat.stride(1), _ptr(att), att.stride(1), n, hw, c, m, _ptr(fm), _ptr(sq), _stream()), "saspa_bap_head")
    _count(3)
    return fm


def fc_f32(x: torch.Tensor, w: torch.Tensor, bias) -> torch.Tensor:
    _need_cuda(x, w)
    n, k = x.shape
    classes = w.shape[0]
    assert x.dtype == torch.float32 and w.dtype == torch.float32 and x.is_contiguous() and w.is_contiguous() and w.shape[1] == k
    out = torch.empty((n, classes), dtype=torch.float32, device=x.device)
    check(_lib.load().saspa_fc_f32(_ptr(x), _ptr(w), _ptr(bias), n, k, classes, _ptr(out), _stream()), "saspa_fc_f32")
    _count((n + 7) // 8)
    return out


def topk_contains(logits: torch.Tensor, labels: torch.Tensor, k: int):
    _need_cuda(logits, labels)
    n, classes = logits.shape
    assert logits.dtype == torch.float32 and logits.is_contiguous() and labels.dtype == torch.int32
    keep = torch.empty((n,), dtype=torch.uint8, device=logits.device)
    margin = torch.empty((n,), dtype=torch.float32, device=logits.device)
    check(_lib.load().saspa_topk_contains(_ptr(logits), n, classes, _ptr(labels), int(k), _ptr(keep), _ptr(margin), _stream()), "saspa_topk_contains")
    _count()
    return keep, margin


def softmax_at(logits: torch.Tensor, idx: torch.Tensor):
    """logits fp32 [n, C], idx int32 [n] -> (softmax(logits)[i, idx[i]] fp32 [n], max logit fp32 [n], argmax int32 [n])."""
    _need_cuda(logits, idx)
    n, classes = logits.shape
    assert logits.dtype == torch.float32 and logits.is_contiguous() and idx.dtype == torch.int32 and idx.numel() == n
    prob = torch.empty((n,), dtype=torch.float32, device=logits.device)
    mx = torch.empty((n,), dtype=torch.float32, device=logits.device)
    arg = torch.empty((n,), dtype=torch.int32, device=logits.device)
    check(_lib.load().saspa_softmax_at_f32(_ptr(logits), n, classes, _ptr(idx), _ptr(prob), _ptr(mx), _ptr(arg), _stream()), "saspa_softmax_at_f32")
    _count()
    return prob, mx, arg


def clip_score_argmax(img: torch.Tensor, txt: torch.Tensor, logit_scale: float):
    _need_cuda(img, txt)
    n, d = img.shape
    p = txt.shape[0]
    assert img.dtype == torch.float32 and txt.dtype == torch.float32 and img.is_contiguous() and txt.is_contiguous()
    logits = torch.empty((n, p), dtype=torch.float32, device=img.device)
    arg = torch.empty((n,), dtype=torch.int32, device=img.device)
    check(_lib.load().saspa_clip_score_argmax(_ptr(img), _ptr(txt), n, p, d, float(logit_scale), _ptr(logits), _ptr(arg), _stream()), "saspa_clip_score_argmax")
    _count()
    return logits, arg


def rgb_to_luma3(img: torch.Tensor) -> torch.Tensor:
    """u8 [..., 3] -> u8 same shape: PIL ``convert("L").convert("RGB")`` (bit-exact)."""
    _need_cuda(img)
    assert img.dtype == torch.uint8 and img.shape[-1] == 3 and img.is_contiguous()
    out = torch.empty_like(img)
    check(_lib.load().saspa_rgb_to_luma3_u8(_ptr(img), img.numel() // 3, _ptr(out), _stream()), "saspa_rgb_to_luma3_u8")
    _count()
    return out


def lpips_layer_accum(f0: torch.Tensor, f1: torch.Tensor, w: torch.Tensor, accum: torch.Tensor) -> None:
    """accum[n] += spatial mean of sum_c w_c (unit(f0) - unit(f1))^2 for NHWC bf16 feature maps [n,h,w,c]."""
    _need_cuda(f0, f1, w, accum)
    n, h, ww, c = f0.shape
    assert f0.dtype == BF16 and f1.dtype == BF16 and f0.is_contiguous() and f1.is_contiguous() and f1.shape == f0.shape
    assert w.dtype == torch.float32 and w.numel() == c and accum.dtype == torch.float32 and accum.numel() == n
    check(_lib.load().saspa_lpips_layer_accum(_ptr(f0), _ptr(f1), _ptr(w), n, h * ww, c, _ptr(accum), _stream()), "saspa_lpips_layer_accum")
    _count()


def hed_fuse(sides, H: int, W: int, safe: bool = False, out_channels: int = 3) -> torch.Tensor:
    """Five fp32 side outputs [n,h_k,w_k,1] (the HED projections) -> u8 control map [n,H,W,out_channels]: cv2-style bilinear resize of
    each to H x W, mean, sigmoid, (safe_step,) * 255 truncated (controlnet_aux HEDdetector.__call__)."""
    import ctypes

    assert len(sides) == 5
    _need_cuda(*sides)
    n = sides[0].shape[0]
    for s in sides:
        assert s.dtype == torch.float32 and s.dim() == 4 and s.shape[0] == n and s.shape[3] == 1 and s.is_contiguous()
    out = torch.empty((n, H, W, out_channels), dtype=torch.uint8, device=sides[0].device)
    ptrs = (ctypes.c_void_p * 5)(*[s.data_ptr() for s in sides])
    hs = (ctypes.c_int * 5)(*[s.shape[1] for s in sides])
    ws = (ctypes.c_int * 5)(*[s.shape[2] for s in sides])
    lds = (ctypes.c_int * 5)(*[1] * 5)
    check(_lib.load().saspa_hed_fuse_u8(ptrs, hs, ws, lds, n, int(H), int(W), 1 if safe else 0, _ptr(out), int(out_channels), _stream()), "saspa_hed_fuse_u8")
    _count()
    return out


def resize_area(img: torch.Tensor, dh: int, dw: int) -> torch.Tensor:
    """cv2.resize(img, (dw, dh), interpolation=cv2.INTER_AREA) for ONE u8 HWC image on the device (bit-exact, see csrc/resize_area.cu)."""
    _need_cuda(img)
    assert img.dtype == torch.uint8 and img.dim() == 3 and img.is_contiguous()
    sh, sw, c = img.shape
    lib = _lib.load()
    ws_bytes = lib.saspa_resize_area_workspace_bytes(sh, sw, int(dh), int(dw))
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=img.device)
    out = torch.empty((int(dh), int(dw), c), dtype=torch.uint8, device=img.device)
    check(lib.saspa_resize_area_u8(_ptr(img), sh, sw, c, _ptr(out), int(dh), int(dw), _ptr(ws), ws_bytes, _stream()), "saspa_resize_area_u8")
    _count()
    return out


def resize_lanczos4(img: torch.Tensor, dh: int, dw: int) -> torch.Tensor:
    """cv2.resize(img, (dw, dh), interpolation=cv2.INTER_LANCZOS4) for ONE u8 HWC image on the device (bit-exact, csrc/resize_area.cu)."""
    _need_cuda(img)
    assert img.dtype == torch.uint8 and img.dim() == 3 and img.is_contiguous()
    sh, sw, c = img.shape
    lib = _lib.load()
    ws_bytes = lib.saspa_resize_lanczos4_workspace_bytes(int(dh), int(dw))
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=img.device)
    out = torch.empty((int(dh), int(dw), c), dtype=torch.uint8, device=img.device)
    check(lib.saspa_resize_lanczos4_u8(_ptr(img), sh, sw, c, _ptr(out), int(dh), int(dw), _ptr(ws), ws_bytes, _stream()), "saspa_resize_lanczos4_u8")
    _count()
    return out
