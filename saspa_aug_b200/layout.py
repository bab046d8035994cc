"""Weight re-layouts from diffusers / torch checkpoint layouts (OIHW convs, [out,in] linears) to the
K-major layouts the kernels read."""
from __future__ import annotations

import torch

GEGLU_TILE = 256  # BN of the GEGLU GEMM tile: 128 value columns followed by their 128 gate columns


def conv_weight_kmajor(w: torch.Tensor, kpad: int | None = None) -> torch.Tensor:
    """OIHW [cout, cin, kh, kw] -> [cout, kh*kw*cin] (kh, kw, cin) order, optionally zero padded along K."""
    cout, cin, kh, kw = w.shape
    k = kh * kw * cin
    out = w.permute(0, 2, 3, 1).reshape(cout, k)
    if kpad is not None and kpad != k:
        p = torch.zeros((cout, kpad), dtype=w.dtype, device=w.device)
        p[:, :k] = out
        out = p
    return out.contiguous()


def geglu_interleave(w: torch.Tensor, bias: torch.Tensor | None):
    """diffusers GEGLU: proj = Linear(dim, 2*inner); hidden, gate = proj(x).chunk(2); out = hidden * gelu(gate)
    (models/activations.py).  The fused epilogue needs value and gate of the same output feature in one
    256-wide tile: rows are regrouped as [v[0:128], g[0:128], v[128:256], g[128:256], ...]."""
    two_inner = w.shape[0]
    inner = two_inner // 2
    half = GEGLU_TILE // 2
    assert inner % half == 0, f"GEGLU inner dim {inner} must be a multiple of {half}"
    v, g = w[:inner], w[inner:]
    wi = torch.stack([v.reshape(inner // half, half, -1), g.reshape(inner // half, half, -1)], dim=1).reshape(two_inner, -1).contiguous()
    bi = None
    if bias is not None:
        bv, bg = bias[:inner], bias[inner:]
        bi = torch.stack([bv.reshape(-1, half), bg.reshape(-1, half)], dim=1).reshape(two_inner).contiguous()
    return wi, bi
