"""GPU parity: fused attention vs torch fp32 softmax(QK^T/sqrt(d))V on the same bf16 inputs.
Tolerance 2e-2 absolute on outputs of O(1) magnitude (P is rounded to bf16 before the PV product)."""
import pytest
import torch

from saspa_aug_b200 import ops

pytestmark = pytest.mark.gpu


def _rand(shape, seed, scale=1.0):
    g = torch.Generator(device="cpu").manual_seed(seed)
    return (torch.randn(shape, generator=g) * scale).to(torch.bfloat16).cuda()


def _ref(q, k, v, heads, scale=None, causal=False):
    b, tq, hd = q.shape
    d = hd // heads
    qf, kf, vf = (t.float().view(b, -1, heads, d).transpose(1, 2) for t in (q, k, v))
    s = (qf @ kf.transpose(-1, -2)) * (scale if scale is not None else d ** -0.5)
    if causal:
        s = s.masked_fill(torch.ones(s.shape[-2:], dtype=torch.bool, device=s.device).triu(1), float("-inf"))
    return (torch.softmax(s, dim=-1) @ vf).transpose(1, 2).reshape(b, tq, hd)


@pytest.fixture(params=[0, 1, 2, 3], ids=["auto", "mma_sync", "tcgen05", "resident_mma_sync"])
def impl(request):
    """Runs a test once per attention kernel (saspa_attention_impl: 0 = auto, incl. the persistent tcgen05 cross-attention kernel
    for tkv <= 128; 1 = mma.sync flash; 2 = tcgen05/TMEM only; 3 = the older mma.sync K/V-resident cross-attention kernel)."""
    from saspa_aug_b200 import _lib

    prev = _lib.load().saspa_attention_impl(request.param)
    yield request.param
    _lib.load().saspa_attention_impl(prev)


@pytest.mark.parametrize("cfg", [  # b, heads, tq, tkv, d
    (2, 8, 4096, 4096, 40), (2, 8, 1024, 1024, 80), (4, 8, 256, 256, 160), (4, 8, 64, 64, 160), (2, 8, 4096, 77, 40), (2, 8, 1024, 77, 80),
    (3, 8, 256, 77, 160), (2, 8, 64, 77, 160), (2, 12, 77, 77, 64), (1, 8, 5632, 5632, 40), (2, 5, 100, 131, 64), (2, 16, 257, 257, 64),
    (7, 8, 50, 50, 64), (2, 4, 200, 200, 128), (1, 8, 1000, 77, 40), (2, 5, 300, 128, 64), (1, 8, 5632, 77, 40), (2, 8, 320, 1, 80),
    # cross-attention, persistent kernel: many head switches per CTA, SDXL head dims, 81..128 keys, one-deep K/V ring (d 128, 128 keys)
    (40, 8, 256, 77, 40), (4, 20, 1024, 77, 64), (3, 10, 4096, 77, 64), (2, 8, 256, 77, 128), (9, 8, 384, 100, 80), (5, 4, 640, 128, 128),
    (64, 8, 4096, 77, 40)])
def test_attention(cuda_device, cfg, impl):
    b, heads, tq, tkv, d = cfg
    q, k, v = _rand((b, tq, heads * d), 1), _rand((b, tkv, heads * d), 2), _rand((b, tkv, heads * d), 3)
    got = ops.attention(q, k, v, heads)
    ref = _ref(q, k, v, heads)
    err = (got.float() - ref).abs().max().item()
    assert torch.isfinite(got.float()).all() and err < 2e-2, err


@pytest.mark.parametrize("cfg", [(3, 12, 77, 64), (2, 8, 300, 64), (1, 8, 1024, 40), (2, 4, 515, 128)])  # b, heads, t, d
def test_attention_causal(cuda_device, cfg, impl):
    b, heads, t, d = cfg
    q, k, v = _rand((b, t, heads * d), 11), _rand((b, t, heads * d), 12), _rand((b, t, heads * d), 13)
    got = ops.attention(q, k, v, heads, causal=True)
    ref = _ref(q, k, v, heads, causal=True)
    err = (got.float() - ref).abs().max().item()
    assert torch.isfinite(got.float()).all() and err < 2e-2, err


@pytest.mark.parametrize("d", [40, 64, 80, 160])
def test_attention_growing_logits_rescale(cuda_device, d, impl):
    """Keys whose scores grow along the sequence: the running max rises in every key tile, which drives the
    tcgen05 kernel's lazy O-rescale path (max growth > 2^8) many times per row."""
    b, heads, t = 1, 8, 2048
    q = _rand((b, t, heads * d), 21).abs()
    k = _rand((b, t, heads * d), 22).abs() * torch.linspace(0.05, 6.0, t, device="cuda").view(1, t, 1).to(torch.bfloat16)
    v = _rand((b, t, heads * d), 23)
    got = ops.attention(q, k, v, heads)
    ref = _ref(q, k, v, heads)
    err = (got.float() - ref).abs().max().item()
    assert torch.isfinite(got.float()).all() and err < 3e-2, err


@pytest.mark.parametrize("d", [80, 40])
def test_attention_fused_qkv_view_and_peaked_scores(cuda_device, impl, d):
    b, t, heads = 2, 1024, 8
    qkv = _rand((b, t, 3 * heads * d), 4, 3.0)  # large logits -> peaked softmax
    q, k, v = qkv[..., : heads * d], qkv[..., heads * d : 2 * heads * d], qkv[..., 2 * heads * d :]
    out = torch.zeros((b, t, 2 * heads * d), dtype=torch.bfloat16, device="cuda")
    ops.attention(q, k, v, heads, out=out[..., : heads * d])
    ref = _ref(q, k, v, heads)
    assert (out[..., : heads * d].float() - ref).abs().max().item() < 6e-2
    assert out[..., heads * d :].abs().max() == 0


def test_softmax_rows_and_transpose(cuda_device):
    x = _rand((300, 4096), 5, 4.0)
    got = ops.softmax_rows(x, 0.044)
    ref = torch.softmax(x.float() * 0.044, dim=-1)
    assert (got.float() - ref).abs().max().item() < 2e-3
    t = _rand((3, 100, 72), 6)
    assert torch.equal(ops.transpose(t), t.transpose(1, 2).contiguous())
