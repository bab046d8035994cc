"""GPU parity: tcgen05 GEMM / implicit-GEMM conv (through the C ABI) vs a plain PyTorch fp32 reference of the
same op on the same bf16-rounded inputs.  Tolerance: bf16 output rounding (rel 2^-8) + fp32 accumulation
order => |err| <= 2e-2 * max|ref| is generous; typical observed error is ~4e-3 relative."""
import math

import pytest
import torch
import torch.nn.functional as F

from saspa_aug_b200 import ops

pytestmark = pytest.mark.gpu


def _rand(shape, seed, scale=1.0):
    g = torch.Generator(device="cpu").manual_seed(seed)
    return (torch.randn(shape, generator=g) * scale).to(torch.bfloat16).cuda()


def _check(got, ref, tol=2e-2):
    got = got.float()
    denom = ref.abs().max().clamp_min(1e-6)
    err = (got - ref).abs().max() / denom
    assert torch.isfinite(got).all()
    assert err.item() < tol, f"rel max err {err.item():.4g}"


@pytest.mark.parametrize("mnk", [(128, 128, 64), (128, 256, 128), (256, 160, 320), (4096, 320, 320), (1000, 320, 2880), (8192, 640, 640),
                                 (77, 768, 768), (300, 96, 200), (64, 1280, 1280), (130, 40, 72), (2048, 1280, 5120), (512, 512, 4608)])
def test_gemm_plain(cuda_device, mnk):
    M, N, K = mnk
    a, b = _rand((M, K), 1), _rand((N, K), 2, 1.0 / math.sqrt(K))
    got = ops.gemm(a, b)
    _check(got, a.float() @ b.float().t())


def test_gemm_strided_views_and_fp32_out(cuda_device):
    big_a, big_b = _rand((500, 704), 3), _rand((320, 640), 4, 0.05)
    a, b = big_a[:, 64:64 + 320], big_b[:, :320]
    out = torch.zeros((500, 512), dtype=torch.float32, device="cuda")
    ops.gemm(a, b, out=out[:, 128:128 + 320])
    _check(out[:, 128:448], a.float() @ b.float().t(), 1e-3)
    assert out[:, :128].abs().max() == 0 and out[:, 448:].abs().max() == 0


@pytest.mark.parametrize("act,fn", [(ops.ACT_SILU, F.silu), (ops.ACT_GELU, F.gelu), (ops.ACT_RELU, F.relu),
                                    (ops.ACT_QUICKGELU, lambda x: x * torch.sigmoid(1.702 * x))])
def test_gemm_epilogue_bias_act_residual(cuda_device, act, fn):
    M, N, K = 1024, 640, 320
    a, b = _rand((M, K), 5), _rand((N, K), 6, 1.0 / math.sqrt(K))
    bias = torch.randn(N, device="cuda")
    rows_per_group = 256
    rb = torch.randn(M // rows_per_group, N, device="cuda")
    res = _rand((M, N), 7)
    got = ops.gemm(a, b, bias=bias, row_bias=rb, rows_per_group=rows_per_group, act=act, alpha=0.75, residual=res, beta=0.5)
    ref = a.float() @ b.float().t() + bias + rb.repeat_interleave(rows_per_group, 0)
    ref = 0.75 * fn(ref) + 0.5 * res.float()
    _check(got, ref)


def test_gemm_geglu(cuda_device):
    from saspa_aug_b200.layout import geglu_interleave

    M, K, inner = 512, 320, 1280
    a = _rand((M, K), 8)
    w = _rand((2 * inner, K), 9, 1.0 / math.sqrt(K))  # diffusers GEGLU proj: [value; gate]
    bias = torch.randn(2 * inner, device="cuda")
    wi, bi = geglu_interleave(w, bias)
    got = ops.gemm(a, wi, bias=bi, act=ops.ACT_GEGLU)
    h = a.float() @ w.float().t() + bias
    ref = h[:, :inner] * F.gelu(h[:, inner:])
    assert got.shape == (M, inner)
    _check(got, ref)


@pytest.fixture(params=[0, 1], ids=["ctas-auto-pair", "ctas-single"], autouse=True)
def gemm_ctas(request):
    """Every test of this file runs with two-CTA (cta_group::2) tiles where eligible and with single-CTA tiles only."""
    from saspa_aug_b200 import _lib

    prev = _lib.load().saspa_gemm_force_ctas(request.param)
    yield request.param
    _lib.load().saspa_gemm_force_ctas(prev)


@pytest.fixture(params=[0, 1], ids=["conv-auto-halo", "conv-per-tap"])
def conv_impl(request):
    """Runs a conv test once per 3x3 main loop (saspa_conv_impl: 0 = auto / halo tile where eligible, 1 = one TMA box per tap)."""
    from saspa_aug_b200 import _lib

    prev = _lib.load().saspa_conv_impl(request.param)
    yield request.param
    _lib.load().saspa_conv_impl(prev)


@pytest.mark.parametrize("cfg", [  # n, h, w, cin, cout, ksize
    (1, 17, 9, 64, 64, 3), (3, 40, 24, 72, 40, 3),
    (2, 64, 64, 320, 320, 3), (4, 32, 32, 640, 640, 3), (4, 16, 16, 1280, 1280, 3), (6, 8, 8, 1280, 1280, 3),
    (2, 64, 88, 320, 320, 3), (1, 128, 128, 128, 128, 3), (3, 16, 16, 64, 96, 3), (2, 32, 32, 320, 640, 1), (1, 24, 40, 128, 256, 3),
    (2, 8, 8, 2560, 1280, 3), (2, 64, 64, 16, 16, 3), (2, 32, 32, 32, 32, 3), (1, 64, 64, 96, 96, 3), (2, 16, 16, 8, 24, 3), (2, 64, 64, 320, 4, 3)])
def test_conv_igemm(cuda_device, conv_impl, cfg):
    n, h, w, cin, cout, ks = cfg
    x = _rand((n, h, w, cin), 10)
    wt = _rand((cout, cin, ks, ks), 11, 1.0 / math.sqrt(cin * ks * ks))
    bias = torch.randn(cout, device="cuda")
    wk = wt.permute(0, 2, 3, 1).reshape(cout, ks * ks * cin).contiguous()
    got = ops.conv2d_igemm(x, wk, ks, bias=bias)
    ref = F.conv2d(x.float().permute(0, 3, 1, 2), wt.float(), bias, padding=ks // 2).permute(0, 2, 3, 1)
    _check(got, ref)


@pytest.mark.parametrize("cfg", [  # n, h, w, cin, cout, pad (0 = the VAE encoder's bottom/right-only padding)
    (2, 64, 64, 320, 320, 1), (4, 32, 32, 640, 640, 1), (4, 16, 16, 1280, 1280, 1), (2, 33, 47, 64, 96, 1), (1, 17, 9, 72, 40, 1),
    (2, 64, 64, 128, 128, 0), (1, 128, 96, 256, 256, 0), (3, 10, 14, 64, 64, 0), (1, 512, 512, 128, 128, 0)])
def test_conv_igemm_stride2(cuda_device, cfg):
    """Stride-2 3x3 implicit GEMM (TMA element strides; no im2col buffer): diffusers Downsample2D (padding 1) and the VAE encoder's
    F.pad(x, (0,1,0,1)) + conv(stride 2, padding 0); with a per-image row bias and activation in the epilogue."""
    n, h, w, cin, cout, pad = cfg
    x = _rand((n, h, w, cin), 20)
    wt = _rand((cout, cin, 3, 3), 21, 1.0 / math.sqrt(cin * 9))
    bias = torch.randn(cout, device="cuda")
    wk = wt.permute(0, 2, 3, 1).reshape(cout, 9 * cin).contiguous()
    xin = x.float().permute(0, 3, 1, 2)
    if pad == 0:
        ref = F.conv2d(F.pad(xin, (0, 1, 0, 1)), wt.float(), bias, stride=2).permute(0, 2, 3, 1)
    else:
        ref = F.conv2d(xin, wt.float(), bias, stride=2, padding=1).permute(0, 2, 3, 1)
    oh, ow = ref.shape[1:3]
    got = ops.conv2d_igemm(x, wk, 3, bias=bias, stride=2, pad=pad, out_hw=(oh, ow))
    _check(got, ref)
    rb = torch.randn(n, cout, device="cuda")
    got2 = ops.conv2d_igemm(x, wk, 3, bias=bias, row_bias=rb, act=ops.ACT_SILU, stride=2, pad=pad, out_hw=(oh, ow))
    _check(got2, F.silu(ref + rb[:, None, None, :]))


def test_conv_igemm_two_sources_rowbias_residual(cuda_device, conv_impl):
    n, h, w, c0, c1, cout = 2, 32, 32, 640, 320, 640
    x0, x1 = _rand((n, h, w, c0), 12), _rand((n, h, w, c1), 13)
    wt = _rand((cout, c0 + c1, 3, 3), 14, 1.0 / math.sqrt(9 * (c0 + c1)))
    wk = wt.permute(0, 2, 3, 1).reshape(cout, -1).contiguous()
    bias = torch.randn(cout, device="cuda")
    temb = torch.randn(n, cout, device="cuda")
    res = _rand((n, h, w, cout), 15)
    got = ops.conv2d_igemm(x0, wk, 3, x1=x1, bias=bias, row_bias=temb, residual=res, beta=1.0)
    xin = torch.cat([x0, x1], dim=3).float().permute(0, 3, 1, 2)
    ref = (F.conv2d(xin, wt.float(), bias, padding=1) + temb[:, :, None, None]).permute(0, 2, 3, 1) + res.float()
    _check(got, ref)


def test_conv_igemm_channel_slice_views(cuda_device, conv_impl):
    """Inputs / outputs living inside wider concat buffers (pixel stride > channels)."""
    n, h, w = 2, 16, 16
    buf = _rand((n, h, w, 1280 + 640), 16)
    x = buf[..., 1280:]
    wt = _rand((320, 640, 3, 3), 17, 0.02)
    wk = wt.permute(0, 2, 3, 1).reshape(320, -1).contiguous()
    outbuf = torch.zeros((n, h, w, 960), dtype=torch.bfloat16, device="cuda")
    ops.conv2d_igemm(x, wk, 3, out=outbuf[..., 640:])
    ref = F.conv2d(x.float().permute(0, 3, 1, 2), wt.float(), None, padding=1).permute(0, 2, 3, 1)
    _check(outbuf[..., 640:], ref)
    assert outbuf[..., :640].abs().max() == 0


@pytest.mark.parametrize("cfg", [(2, 64, 64, 4, 320, 3, 1, 1), (2, 64, 64, 320, 320, 3, 2, 1), (2, 224, 224, 3, 64, 7, 2, 3), (1, 512, 512, 3, 16, 3, 1, 1),
                                 (2, 65, 65, 128, 128, 3, 2, 0)])
def test_im2col_gemm_conv(cuda_device, cfg):
    n, h, w, cin, cout, ks, stride, pad = cfg
    x = _rand((n, h, w, cin), 18)
    wt = _rand((cout, cin, ks, ks), 19, 1.0 / math.sqrt(cin * ks * ks))
    oh = (h + 2 * pad - ks) // stride + 1
    ow = (w + 2 * pad - ks) // stride + 1
    K = ks * ks * cin
    kpad = (K + 7) // 8 * 8
    wk = torch.zeros((cout, kpad), dtype=torch.bfloat16, device="cuda")
    wk[:, :K] = wt.permute(0, 2, 3, 1).reshape(cout, K)
    cols = ops.im2col(x, ks, ks, stride, pad, pad, oh, ow, kpad)
    got = ops.gemm(cols, wk).view(n, oh, ow, cout)
    ref = F.conv2d(x.float().permute(0, 3, 1, 2), wt.float(), None, stride=stride, padding=pad).permute(0, 2, 3, 1)
    _check(got, ref)


def test_gemm_full_size_linearity(cuda_device):
    """BASELINE config-2 sized GEMM (32 CFG rows x 4096 tokens): size-independent property
    gemm(a1 + a2, b) == gemm(a1, b) + gemm(a2, b) within rounding, plus a sampled exact check."""
    M, N, K = 32 * 4096, 320, 2880
    a1, a2 = _rand((M, K), 20, 0.5), _rand((M, K), 21, 0.5)
    b = _rand((N, K), 22, 1.0 / math.sqrt(K))
    s = (a1.float() + a2.float()).to(torch.bfloat16)
    g1, g2, gs = ops.gemm(a1, b, out_fp32=True), ops.gemm(a2, b, out_fp32=True), ops.gemm(s, b, out_fp32=True)
    # s was re-rounded to bf16, so compare against the reference of s on sampled rows instead of g1+g2 exactly
    idx = torch.randint(0, M, (256,), device="cuda")
    _check(gs[idx], s[idx].float() @ b.float().t(), 2e-3)
    assert ((g1 + g2) - gs).abs().max() / gs.abs().max() < 2e-2


def test_empty_inputs_and_argument_errors(cuda_device):
    """C-ABI error behaviour: zero-sized problems are no-ops that return 0; malformed arguments return a negative code that the
    binding turns into SaspaError (a RuntimeError, so the reference's `except RuntimeError`, run_aug.py:493, still catches it) with
    the entry point's message -- never a crash, never a silent fallback."""
    from saspa_aug_b200.ops import SaspaError

    a0 = torch.empty((0, 64), dtype=torch.bfloat16, device="cuda")
    b = _rand((32, 64), 1)
    assert ops.gemm(a0, b).shape == (0, 32)
    x0 = torch.empty((0, 16, 16, 64), dtype=torch.bfloat16, device="cuda")
    wk = _rand((64, 9 * 64), 2)
    assert ops.conv2d_igemm(x0, wk, 3).shape == (0, 16, 16, 64)
    q0 = torch.empty((0, 128, 64), dtype=torch.bfloat16, device="cuda")
    assert ops.attention(q0, q0, q0, 1).shape == (0, 128, 64)
    x = _rand((1, 16, 16, 64), 3)
    with pytest.raises(SaspaError, match="ksize"):
        ops.conv2d_igemm(x, _rand((64, 25 * 64), 4), 5)
    with pytest.raises(SaspaError, match="stride"):
        ops.conv2d_igemm(x, wk, 3, stride=3, pad=1, out_hw=(6, 6))
    with pytest.raises((SaspaError, AssertionError)):
        ops.gemm(_rand((8, 60), 5), _rand((8, 60), 6))  # K = 60: row stride not a multiple of 8 elements
    assert isinstance(SaspaError("x"), RuntimeError)
    with pytest.raises((SaspaError, AssertionError)):
        ops.gemm(torch.zeros((8, 64), dtype=torch.bfloat16), torch.zeros((8, 64), dtype=torch.bfloat16))  # CPU tensors: no CPU fallback


@pytest.mark.parametrize("M,C,N,act", [(300, 320, 960, 0), (4096 + 77, 640, 640, 0), (1000, 1280, 2560, 5), (257, 320, 2560, 5), (129, 1280, 3840, 0)])
def test_layernorm_folded_into_gemm(cuda_device, M, C, N, act):
    """LN(x) W^T + b computed as rstd * (x W'^T - mean * colsum(W')) + (b + W beta) in the consumer's epilogue, with the row
    statistics emitted by the PRODUCER GEMM's epilogue (diffusers BasicTransformerBlock norm -> projection, models/attention.py).
    Reference: torch fp32 LayerNorm + matmul on the same bf16 inputs."""
    import torch.nn.functional as F

    from saspa_aug_b200.layout import geglu_interleave
    from saspa_aug_b200.ops import ACT_GEGLU

    g = torch.Generator().manual_seed(M + N)
    # producer: h = a @ wp^T + residual, bf16, with row statistics of the stored values
    a = torch.randn((M, 256), generator=g).to(torch.bfloat16).cuda()
    wp = (torch.randn((C, 256), generator=g) / 16).to(torch.bfloat16).cuda()
    res = (torch.randn((M, C), generator=g) * 2 + 0.7).to(torch.bfloat16).cuda()  # non-zero mean: exercises the mean * colsum term
    h, st = ops.gemm(a, wp, residual=res, beta=1.0, row_stats=True)
    assert st.shape == (M, ops.row_stats_slots(C), 2) and st.dtype == torch.float32
    hf = h.float()
    # the sums are taken on the fp32 values just before the bf16 rounding of the store: each term is within 2^-9 relative of the stored one
    assert ((st[..., 0].sum(1) - hf.sum(1)).abs() <= 2.0 ** -8 * hf.abs().sum(1) + 1e-3).all()
    assert ((st[..., 1].sum(1) - (hf * hf).sum(1)).abs() <= 2.0 ** -7 * (hf * hf).sum(1) + 1e-2).all()
    # a row's statistics do not depend on how many rows share the launch
    h2, st2 = ops.gemm(a[:130], wp, residual=res[:130], beta=1.0, row_stats=True)
    assert torch.equal(h2, h[:130]) and torch.equal(st2, st[:130])
    # consumer
    gamma = 1.0 + 0.2 * torch.randn(C, generator=g)
    beta = 0.1 * torch.randn(C, generator=g)
    w = torch.randn((N, C), generator=g) / C ** 0.5
    b = 0.1 * torch.randn(N, generator=g)
    want = F.linear(F.layer_norm(hf.cpu(), (C,), gamma, beta, 1e-5), w, b)
    if act == ACT_GEGLU:
        val, gate = want.chunk(2, dim=-1)
        want = val * F.gelu(gate)
        w, b = geglu_interleave(w, b)
    wf = (w * gamma[None, :]).to(torch.bfloat16).cuda()
    bias = (w @ beta + b).float().cuda()
    got = ops.gemm(h, wf, bias=bias, act=act, ln_stats=st, ln_colsum=wf.float().sum(1).contiguous(), ln_eps=1e-5).float().cpu()
    # the unfused kernels on the same inputs (LayerNorm rounds its output to bf16 first): the folded path must be at least as close
    y = ops.layernorm(h, 1e-5, gamma.cuda(), beta.cuda())
    unfused = ops.gemm(y, w.to(torch.bfloat16).cuda(), bias=b.float().cuda(), act=act).float().cpu()
    e_fold, e_unf = (got - want).abs().max().item(), (unfused - want).abs().max().item()
    print(f"LN fold M={M} C={C} N={N} act={act}: max err folded {e_fold:.4g}, unfused {e_unf:.4g}, max|ref| {want.abs().max().item():.3g}")
    assert e_fold <= max(1.5 * e_unf, 2e-2 * want.abs().max().item())


@pytest.mark.parametrize("n,h,w,cin,cout,stride,act,sliced", [
    (2, 40, 56, 3, 16, 1, "silu", False), (2, 64, 48, 16, 16, 1, "silu", False), (3, 33, 47, 16, 32, 2, "silu", False),
    (1, 64, 64, 32, 32, 1, "none", False), (2, 32, 32, 32, 96, 2, "silu", False), (1, 48, 48, 3, 128, 1, "none", False),
    (1, 48, 40, 3, 64, 1, "relu", False), (2, 17, 19, 4, 20, 1, "gelu", False), (2, 24, 24, 16, 16, 1, "silu", True),
    (4, 128, 128, 16, 16, 1, "silu", False)])
def test_conv3x3_small_channels(cuda_device, n, h, w, cin, cout, stride, act, sliced):
    """saspa_conv3x3_small_bf16 (ControlNet conditioning embedding 3->16->16->32->32->96, VAE / HED stems) against torch conv2d in fp32 on
    the same bf16 inputs and weights: odd map sizes (partial tiles, image borders = zero padding), stride 2, Cin 3 / 4 (scalar staging),
    Cout not a multiple of 8, a channel-sliced input view, every epilogue activation.  Also equal to the im2col + GEMM path it replaces
    up to the bf16 rounding of the output."""
    import torch.nn.functional as F

    from saspa_aug_b200.layout import conv_weight_kmajor

    g = torch.Generator().manual_seed(n * 100 + cin + cout)
    acts = {"none": (ops.ACT_NONE, lambda t: t), "silu": (ops.ACT_SILU, F.silu), "relu": (ops.ACT_RELU, F.relu), "gelu": (ops.ACT_GELU, F.gelu)}
    code, fn = acts[act]
    xs = torch.randn((n, h, w, cin + (8 if sliced else 0)), generator=g).to(torch.bfloat16).cuda()
    x = xs[..., :cin] if sliced else xs
    wt = torch.randn((cout, cin, 3, 3), generator=g) / (9 * cin) ** 0.5
    bias = (0.2 * torch.randn(cout, generator=g)).cuda()
    kpad = (9 * cin + 7) // 8 * 8
    wk = conv_weight_kmajor(wt, kpad).to(torch.bfloat16).cuda()
    assert ops.conv3x3_small_supported(cin, cout, stride, 1)
    got = ops.conv3x3_small(x, wk, bias, code, stride).float()
    want = fn(F.conv2d(x.float().permute(0, 3, 1, 2), wt.to(torch.bfloat16).float().cuda(), bias, stride=stride, padding=1)).permute(0, 2, 3, 1)
    assert got.shape == want.shape
    err = (got - want).abs().max().item()
    assert err <= 1e-2 * want.abs().max().item() + 1e-3, (err, want.abs().max().item())
    # the path it replaces: im2col + tcgen05 GEMM
    oh, ow = want.shape[1:3]
    cols = ops.im2col(x.contiguous(), 3, 3, stride, 1, 1, oh, ow, kpad)
    old = ops.gemm(cols, wk, bias=bias, act=code).view(n, oh, ow, cout).float()
    assert (got - old).abs().max().item() <= 1.6e-2 * want.abs().max().item() + 1e-3
    # per-image results do not depend on the batch
    one = ops.conv3x3_small(x[:1], wk, bias, code, stride).float()
    assert torch.equal(one, got[:1])


def test_conv3x3_small_wide_output_with_residual(cuda_device):
    """conv_in 4 -> 320 of the UNet / ControlNet (the ControlNet adds its conditioning embedding as a residual) on the small-channel kernel:
    slices of 128 output channels written into a channel slice of a wider buffer."""
    import torch.nn.functional as F

    from saspa_aug_b200.layout import conv_weight_kmajor

    g = torch.Generator().manual_seed(5)
    n, h, w, cin, cout = 3, 24, 40, 4, 320
    x = torch.randn((n, h, w, cin), generator=g).to(torch.bfloat16).cuda()
    wt = torch.randn((cout, cin, 3, 3), generator=g) / 6.0
    bias = (0.2 * torch.randn(cout, generator=g)).cuda()
    res = torch.randn((n, h, w, cout), generator=g).to(torch.bfloat16).cuda()
    wk = conv_weight_kmajor(wt, 40).to(torch.bfloat16).cuda()
    wide = torch.zeros((n, h, w, cout + 64), dtype=torch.bfloat16, device="cuda")
    got = ops.conv3x3_small(x, wk, bias, ops.ACT_NONE, 1, out=wide[..., 32 : 32 + cout], residual=res)
    want = (F.conv2d(x.float().permute(0, 3, 1, 2), wt.to(torch.bfloat16).float().cuda(), bias, padding=1).permute(0, 2, 3, 1) + res.float())
    assert (got.float() - want).abs().max().item() <= 1e-2 * want.abs().max().item()
    assert wide[..., :32].abs().max().item() == 0 and wide[..., 32 + cout :].abs().max().item() == 0
