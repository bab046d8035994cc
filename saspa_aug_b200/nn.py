"""Model graphs of the denoising path as launch lists over the sm_100a kernels (saspa_aug_b200.ops).

Host-side mirror of the diffusers modules the reference drives through ``pipe(**pipe_args)``
(run_aug/run_aug.py:278; built at :184-211): UNet2DConditionModel, ControlNetModel, AutoencoderKL,
CLIPTextModel.  Weights come from a diffusers-keyed state dict and are re-laid out once
(conv OIHW -> [Cout, kh*kw*Cin] bf16 K-major, fused QKV, tile-interleaved GEGLU).  Activations are NHWC
bf16; everything below is kernel launches on the current stream -- CUDA-graph capturable, no torch math.

B200-first choices (vs. the diffusers graph):
  * skip connections are never concatenated: producers write straight into the channel slice of the
    consumer's concat buffer (all kernels take row strides);
  * ControlNet residuals are accumulated into the UNet skip tensors by the zero-conv GEMM epilogue
    (alpha = conditioning_scale, beta = 1) -- the UNet encoder + mid run first, then the ControlNet;
  * every ResnetBlock time-embedding projection of a model is ONE GEMM per step; each conv1 reads its
    column slice as a per-image row bias in the epilogue;
  * the step-invariant work is hoisted: ControlNet conditioning embedding and the cross-attention K/V
    projections of the text are computed once per image, not once per step.
"""
from __future__ import annotations

import os
from typing import Dict, List, Optional, Tuple

import torch

from . import ops
from .layout import conv_weight_kmajor, geglu_interleave
from .ops import ACT_GEGLU, ACT_NONE, ACT_QUICKGELU, ACT_SILU, BF16

SD = Dict[str, torch.Tensor]


def _f32(t, dev):
    return None if t is None else t.detach().to(device=dev, dtype=torch.float32).contiguous()


def _bf(t, dev):
    return t.detach().to(device=dev, dtype=BF16).contiguous()


def pix2d(x: torch.Tensor) -> torch.Tensor:
    """NHWC view [n,h,w,c] with dense pixel strides -> 2-D [n*h*w, c] view (row stride = pixel stride)."""
    n, h, w, c = x.shape
    return x.as_strided((n * h * w, c), (x.stride(2), 1))


def pix3d(x: torch.Tensor) -> torch.Tensor:
    n, h, w, c = x.shape
    return x.as_strided((n, h * w, c), (x.stride(0), x.stride(2), 1))


class Conv:
    """Conv2d.  1x1/3x3 with Cin % 8 == 0 and stride 1 ("same") or stride 2 (padding 1, or the VAE encoder's bottom/right-only
    padding) run as TMA implicit GEMM; 3x3 / padding-1 layers with Cin <= 32 and Cout <= 128 (conditioning embedding, image stems) on
    the direct small-channel kernel (conv_in 4 -> 320 as slices of 64 output channels); everything else (7x7 / 14x14 stems, strides > 2)
    as im2col + GEMM."""

    def __init__(self, sd: SD, prefix: str, dev, stride: int = 1, padding: Optional[int] = None, weight=None, bias=None):
        w = sd[prefix + ".weight"] if weight is None else weight
        b = sd.get(prefix + ".bias") if bias is None else bias
        if w.dim() == 2:  # Linear used as a 1x1 conv
            w = w[:, :, None, None]
        self.cout, self.cin, self.k, _ = w.shape
        self.stride = stride
        self.pad = self.k // 2 if padding is None else padding
        self.bias = _f32(b, dev)
        # 3x3 layers with a handful of channels (conditioning embedding, image stems): direct mma.sync kernel, no im2col buffer and no
        # 64-channel k-block padding (csrc/conv_small.cu)
        self.small = self.k == 3 and self.pad == 1 and stride in (1, 2) and self.cin <= 32 and (self.cout <= 128 or self.cin <= 8)
        self.direct = not self.small and stride == 1 and self.pad == self.k // 2 and self.k in (1, 3) and self.cin % 8 == 0
        self.direct_strided = not self.small and stride == 2 and self.k == 3 and self.pad in (0, 1) and self.cin % 8 == 0
        if self.direct or self.direct_strided:
            self.kpad = self.k * self.k * self.cin
            self.w = _bf(conv_weight_kmajor(w), dev)
        else:
            self.kpad = (self.k * self.k * self.cin + 7) // 8 * 8
            self.w = _bf(conv_weight_kmajor(w, self.kpad), dev)

    def __call__(self, x: torch.Tensor, out: Optional[torch.Tensor] = None, pad_extra_br: int = 0, **epi) -> torch.Tensor:
        n, h, w, c = x.shape
        assert c == self.cin, (c, self.cin)
        if self.small and pad_extra_br == 0 and set(epi) <= {"act", "residual", "beta"} and epi.get("beta", 1.0) == 1.0:
            return ops.conv3x3_small(x, self.w, self.bias, epi.get("act", ACT_NONE), self.stride, out=out, residual=epi.get("residual"))
        if self.direct and self.k == 1:
            # a 1x1 convolution on NHWC is the GEMM [pixels, Cin] x [Cout, Cin]^T: the 2-D TMA boxes of the GEMM entry point move the
            # same bytes as the 4-D pixel boxes of the convolution entry point but run 1.3-1.9x faster on these shapes
            # (profiles/r2_conv1x1_as_gemm.txt), and K >= 1024 gets the CTA-pair tiles
            if out is None:
                out = torch.empty((n, h, w, self.cout), dtype=torch.float32 if epi.get("out_fp32") else BF16, device=x.device)
            if epi.get("residual") is not None:
                epi["residual"] = pix2d(epi["residual"])
            if epi.get("row_bias") is not None:
                epi["rows_per_group"] = h * w
            ops.gemm(pix2d(x), self.w, out=pix2d(out), bias=self.bias, **epi)
            return out
        if self.direct:
            return ops.conv2d_igemm(x, self.w, self.k, out=out, bias=self.bias, **epi)
        oh = (h + 2 * self.pad + pad_extra_br - self.k) // self.stride + 1
        ow = (w + 2 * self.pad + pad_extra_br - self.k) // self.stride + 1
        if self.direct_strided:  # taps past the bottom/right edge read TMA zero fill: covers padding 1 and the (0,1,0,1) VAE pad
            return ops.conv2d_igemm(x, self.w, self.k, out=out, bias=self.bias, stride=self.stride, pad=self.pad, out_hw=(oh, ow), **epi)
        cols = ops.im2col(x, self.k, self.k, self.stride, self.pad, self.pad, oh, ow, self.kpad)
        if out is None:
            out = torch.empty((n, oh, ow, self.cout), dtype=torch.float32 if epi.get("out_fp32") else BF16, device=x.device)
        if "residual" in epi and epi["residual"] is not None:
            epi["residual"] = pix2d(epi["residual"])
        if epi.get("row_bias") is not None:
            epi["rows_per_group"] = oh * ow
        ops.gemm(cols, self.w, out=pix2d(out), bias=self.bias, **epi)
        return out


class Linear:
    def __init__(self, sd: SD, prefix: str, dev, weight=None, bias=None):
        w = sd[prefix + ".weight"] if weight is None else weight
        b = sd.get(prefix + ".bias") if bias is None else bias
        self.w = _bf(w, dev)
        self.bias = _f32(b, dev)

    def __call__(self, x2d: torch.Tensor, out=None, **epi) -> torch.Tensor:
        return ops.gemm(x2d, self.w, out=out, bias=self.bias, **epi)


class Norm:
    def __init__(self, sd: SD, prefix: str, dev):
        self.g = _f32(sd[prefix + ".weight"], dev)
        self.b = _f32(sd[prefix + ".bias"], dev)


class ResnetBlock:
    """diffusers ResnetBlock2D: GN-SiLU-conv3x3 (+time bias) - GN-SiLU-conv3x3 + shortcut."""

    def __init__(self, sd: SD, prefix: str, dev, groups: int, eps: float, temb_registry: Optional[list]):
        self.groups, self.eps = groups, eps
        self.norm1 = Norm(sd, prefix + ".norm1", dev)
        self.conv1 = Conv(sd, prefix + ".conv1", dev)
        self.norm2 = Norm(sd, prefix + ".norm2", dev)
        self.conv2 = Conv(sd, prefix + ".conv2", dev)
        self.shortcut = Conv(sd, prefix + ".conv_shortcut", dev) if prefix + ".conv_shortcut.weight" in sd else None
        self.temb_slice = None
        if temb_registry is not None and prefix + ".time_emb_proj.weight" in sd:
            off = sum(w.shape[0] for w, _ in temb_registry)
            temb_registry.append((sd[prefix + ".time_emb_proj.weight"], sd[prefix + ".time_emb_proj.bias"]))
            self.temb_slice = (off, off + self.conv1.cout)

    def _norm_conv(self, x: torch.Tensor, norm: "Norm", conv: "Conv", **epi) -> torch.Tensor:
        n, h, w, c = x.shape
        a = ops.groupnorm(pix3d(x), self.groups, self.eps, norm.g, norm.b, ACT_SILU).view(n, h, w, c)
        return conv(a, **epi)

    def __call__(self, x: torch.Tensor, temb_all: Optional[torch.Tensor], out: Optional[torch.Tensor] = None) -> torch.Tensor:
        rb = temb_all[:, self.temb_slice[0] : self.temb_slice[1]] if (self.temb_slice is not None and temb_all is not None) else None
        h1 = self._norm_conv(x, self.norm1, self.conv1, row_bias=rb)
        res = self.shortcut(x) if self.shortcut is not None else x
        return self._norm_conv(h1, self.norm2, self.conv2, out=out, residual=res, beta=1.0)


FOLD_LAYERNORM = os.environ.get("SASPA_FOLD_LN", "1") != "0"  # A/B switch for tools and tests; the product default is the folded path


class TransformerBlock:
    """diffusers BasicTransformerBlock (self-attn, cross-attn, GEGLU feed-forward; pre-LayerNorm).

    The three LayerNorms never run as kernels: each is folded into the GEMM that consumes it.  With W' = W * gamma (bf16),
    LN(x) W^T + b = rstd * (x W'^T - mean * colsum(W')) + (b + W beta): the GEMM multiplies the UN-normalised residual stream and its
    epilogue applies the row statistics, which the GEMM that PRODUCED the stream (proj_in / to_out / ff.net.2, all with the residual
    add in their epilogue) emitted as per-row partial sums of its bf16 output.  The normalised tensor never exists in HBM."""

    def __init__(self, sd: SD, p: str, dev, heads: int):
        self.heads = heads
        self.fold = FOLD_LAYERNORM
        self.ln1, self.ln2, self.ln3 = Norm(sd, p + ".norm1", dev), Norm(sd, p + ".norm2", dev), Norm(sd, p + ".norm3", dev)
        wqkv = torch.cat([sd[p + ".attn1.to_q.weight"], sd[p + ".attn1.to_k.weight"], sd[p + ".attn1.to_v.weight"]], 0)
        self.o1 = Linear(sd, p + ".attn1.to_out.0", dev)
        wq2 = sd[p + ".attn2.to_q.weight"]
        self.wkv2 = _bf(torch.cat([sd[p + ".attn2.to_k.weight"], sd[p + ".attn2.to_v.weight"]], 0), dev)
        self.o2 = Linear(sd, p + ".attn2.to_out.0", dev)
        wi, bi = geglu_interleave(sd[p + ".ff.net.0.proj.weight"], sd[p + ".ff.net.0.proj.bias"])
        self.ff2 = Linear(sd, p + ".ff.net.2", dev)
        if self.fold:
            self.qkv = self._folded(wqkv, None, sd[p + ".norm1.weight"], sd[p + ".norm1.bias"], dev)
            self.q2 = self._folded(wq2, sd.get(p + ".attn2.to_q.bias"), sd[p + ".norm2.weight"], sd[p + ".norm2.bias"], dev)
            self.ff1 = self._folded(wi, bi, sd[p + ".norm3.weight"], sd[p + ".norm3.bias"], dev)
        else:
            self.wqkv = _bf(wqkv, dev)
            self.q2 = Linear(sd, p + ".attn2.to_q", dev)
            self.ff1 = Linear(sd, "", dev, weight=wi, bias=bi)

    @staticmethod
    def _folded(w, b, gamma, beta, dev):
        """(W' = bf16(W * gamma), fp32 colsum of W' as the MMA sees it, fp32 bias b + W beta).  The two reductions run in float64 and
        are rounded once: a float32 matmul / sum would depend on the summation order (OpenMP thread count, GPU reduction tree), and two
        processes of one job would then build models that differ in the last bit -- enough for visibly different pixels."""
        w64, g64, be64 = w.detach().double().cpu(), gamma.detach().double().cpu(), beta.detach().double().cpu()
        wp = _bf((w64 * g64[None, :]).float(), dev)
        bias = w64 @ be64 + (b.detach().double().cpu() if b is not None else 0.0)
        colsum = wp.detach().cpu().double().sum(dim=1)
        return wp, _f32(colsum.float(), dev), _f32(bias.float(), dev)

    def text_kv(self, text2d: torch.Tensor, n: int) -> Tuple[torch.Tensor, torch.Tensor]:
        """Step-invariant K/V projections of the text states [n*77, cross] -> two [n,77,c] views."""
        kv = ops.gemm(text2d, self.wkv2)
        c = kv.shape[1] // 2
        t = kv.shape[0] // n
        kv3 = kv.view(n, t, 2 * c)
        return kv3[..., :c], kv3[..., c:]

    def __call__(self, h: torch.Tensor, n: int, kv: Tuple[torch.Tensor, torch.Tensor], stats: Optional[torch.Tensor] = None, want_stats: bool = False):
        """h [rows, c] is updated in place.  Folded path: ``stats`` = row statistics of h from its producer; returns (h, stats of the
        new h when ``want_stats``)."""
        rows, c = h.shape
        t = rows // n
        if not self.fold:
            y = ops.layernorm(h, 1e-5, self.ln1.g, self.ln1.b)
            qkv = ops.gemm(y, self.wqkv).view(n, t, 3 * c)
            a = ops.attention(qkv[..., :c], qkv[..., c : 2 * c], qkv[..., 2 * c :], self.heads)
            self.o1(a.view(rows, c), out=h, residual=h, beta=1.0)
            y = ops.layernorm(h, 1e-5, self.ln2.g, self.ln2.b)
            q = self.q2(y).view(n, t, c)
            a = ops.attention(q, kv[0], kv[1], self.heads)
            self.o2(a.view(rows, c), out=h, residual=h, beta=1.0)
            y = ops.layernorm(h, 1e-5, self.ln3.g, self.ln3.b)
            f = self.ff1(y, act=ACT_GEGLU)
            self.ff2(f, out=h, residual=h, beta=1.0)
            return h, None
        w, cs, b = self.qkv
        qkv = ops.gemm(h, w, bias=b, ln_stats=stats, ln_colsum=cs, ln_eps=1e-5).view(n, t, 3 * c)
        a = ops.attention(qkv[..., :c], qkv[..., c : 2 * c], qkv[..., 2 * c :], self.heads)
        _, st = self.o1(a.view(rows, c), out=h, residual=h, beta=1.0, row_stats=True)
        w, cs, b = self.q2
        q = ops.gemm(h, w, bias=b, ln_stats=st, ln_colsum=cs, ln_eps=1e-5).view(n, t, c)
        a = ops.attention(q, kv[0], kv[1], self.heads)
        _, st = self.o2(a.view(rows, c), out=h, residual=h, beta=1.0, row_stats=True)
        w, cs, b = self.ff1
        f = ops.gemm(h, w, bias=b, act=ACT_GEGLU, ln_stats=st, ln_colsum=cs, ln_eps=1e-5)
        if want_stats:
            _, st = self.ff2(f, out=h, residual=h, beta=1.0, row_stats=True)
            return h, st
        self.ff2(f, out=h, residual=h, beta=1.0)
        return h, None


class Transformer2D:
    def __init__(self, sd: SD, p: str, dev, heads: int, groups: int):
        self.groups = groups
        self.norm = Norm(sd, p + ".norm", dev)
        self.proj_in = Linear(sd, "", dev, weight=sd[p + ".proj_in.weight"].reshape(sd[p + ".proj_in.weight"].shape[0], -1), bias=sd[p + ".proj_in.bias"])
        self.proj_out = Linear(sd, "", dev, weight=sd[p + ".proj_out.weight"].reshape(sd[p + ".proj_out.weight"].shape[0], -1), bias=sd[p + ".proj_out.bias"])
        depth = 0
        while f"{p}.transformer_blocks.{depth}.norm1.weight" in sd:
            depth += 1
        self.blocks = [TransformerBlock(sd, f"{p}.transformer_blocks.{i}", dev, heads) for i in range(depth)]

    def text_kv(self, text2d, n):
        return [b.text_kv(text2d, n) for b in self.blocks]

    def __call__(self, x: torch.Tensor, kvs, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        n, hh, ww, c = x.shape
        g = ops.groupnorm(pix3d(x), self.groups, 1e-6, self.norm.g, self.norm.b, ACT_NONE)
        fold = bool(self.blocks) and self.blocks[0].fold
        h, st = self.proj_in(g.view(n * hh * ww, c), row_stats=True) if fold else (self.proj_in(g.view(n * hh * ww, c)), None)
        for bi, (blk, kv) in enumerate(zip(self.blocks, kvs)):
            h, st = blk(h, n, kv, st, want_stats=bi + 1 < len(self.blocks))
        if out is None:
            out = torch.empty((n, hh, ww, c), dtype=BF16, device=x.device)
        self.proj_out(h, out=pix2d(out), residual=pix2d(x), beta=1.0)
        return out


class Downsample:
    def __init__(self, sd: SD, p: str, dev, padding: int = 1):
        self.conv = Conv(sd, p + ".conv", dev, stride=2, padding=padding)
        self.extra = 1 if padding == 0 else 0  # VAE encoder: F.pad(x, (0,1,0,1)) then valid conv

    def __call__(self, x, out=None):
        return self.conv(x, out=out, pad_extra_br=self.extra)


class Upsample:
    def __init__(self, sd: SD, p: str, dev):
        self.conv = Conv(sd, p + ".conv", dev)

    def __call__(self, x, out=None):
        if not x.is_contiguous():
            raise ops.SaspaError("Upsample expects a contiguous NHWC tensor")
        return self.conv(ops.upsample_nearest2x(x), out=out)


# ------------------------------------------------------------------------------------------------
# UNet / ControlNet encoder (shared structure)
# ------------------------------------------------------------------------------------------------
class _EncoderBase:
    def _build_encoder(self, sd: SD, dev, cfg):
        self.cfg, self.dev = cfg, dev
        self.temb_registry: list = []
        g, eps = cfg.norm_num_groups, cfg.norm_eps
        self.conv_in = Conv(sd, "conv_in", dev)
        self.t_lin1 = Linear(sd, "time_embedding.linear_1", dev)
        self.t_lin2 = Linear(sd, "time_embedding.linear_2", dev)
        self.add_emb = None
        if cfg.addition_embed_type == "text_time":
            self.add_emb = (Linear(sd, "add_embedding.linear_1", dev), Linear(sd, "add_embedding.linear_2", dev))
        self.down = []
        nlev = len(cfg.block_out_channels)
        for i, t in enumerate(cfg.down_block_types):
            blk = {"resnets": [], "attns": [], "down": None}
            for j in range(cfg.layers_per_block):
                blk["resnets"].append(ResnetBlock(sd, f"down_blocks.{i}.resnets.{j}", dev, g, eps, self.temb_registry))
                if t.startswith("CrossAttn"):
                    blk["attns"].append(Transformer2D(sd, f"down_blocks.{i}.attentions.{j}", dev, cfg.num_attention_heads[i], g))
            if i != nlev - 1:
                blk["down"] = Downsample(sd, f"down_blocks.{i}.downsamplers.0", dev)
            self.down.append(blk)
        self.mid_res = [ResnetBlock(sd, "mid_block.resnets.0", dev, g, eps, self.temb_registry),
                        ResnetBlock(sd, "mid_block.resnets.1", dev, g, eps, self.temb_registry)]
        self.mid_attn = Transformer2D(sd, "mid_block.attentions.0", dev, cfg.num_attention_heads[-1], g)

    def _finish_temb(self, dev):
        self.temb_w = _bf(torch.cat([w for w, _ in self.temb_registry], 0), dev)
        self.temb_b = _f32(torch.cat([b for _, b in self.temb_registry], 0), dev)

    def skip_channels(self) -> List[Tuple[int, int]]:
        """(channels, level) of every skip tensor in production order (conv_in first)."""
        cfg = self.cfg
        out = [(cfg.block_out_channels[0], 0)]
        n = len(cfg.block_out_channels)
        for i, c in enumerate(cfg.block_out_channels):
            out += [(c, i)] * cfg.layers_per_block
            if i != n - 1:
                out.append((c, i + 1))
        return out

    def added_embed(self, added: Optional[dict]) -> Optional[torch.Tensor]:
        """SDXL "text_time" (diffusers unet_2d_condition.get_aug_embed): add_embedding(cat[pooled text embeds, sinusoid(time_ids)])
        -> bf16 [n, temb].  Step-invariant: computed once per image, consumed as the residual of ``time_embed``."""
        if self.add_emb is None:
            return None
        cfg = self.cfg
        te = added["text_embeds"]  # bf16 [n, proj]
        n = te.shape[0]
        tid = ops.timestep_sinusoid(added["time_ids"].reshape(-1).float().contiguous(), cfg.addition_time_embed_dim, cfg.flip_sin_to_cos, cfg.freq_shift)
        buf = torch.empty((n, te.shape[1] + tid.numel() // n), dtype=BF16, device=te.device)
        buf[:, : te.shape[1]].copy_(te)  # concatenation = data placement only
        buf[:, te.shape[1] :].copy_(tid.view(n, -1))
        return self.add_emb[1](self.add_emb[0](buf, act=ACT_SILU))

    def time_embed(self, t: torch.Tensor, aug: Optional[torch.Tensor] = None) -> torch.Tensor:
        """t fp32 [n] -> fp32 [n, sum(cout of every resnet)]: all time_emb_proj(silu(emb)) in one GEMM.
        ``aug`` = ``added_embed(...)`` for SDXL (emb = time_embedding(t) + aug)."""
        cfg = self.cfg
        s = ops.timestep_sinusoid(t, cfg.block_out_channels[0], cfg.flip_sin_to_cos, cfg.freq_shift)
        h = self.t_lin1(s, act=ACT_SILU)
        emb = self.t_lin2(h, residual=aug, beta=1.0) if aug is not None else self.t_lin2(h)
        return ops.gemm(ops.act(emb, ACT_SILU), self.temb_w, bias=self.temb_b, out_fp32=True)

    def text_kv(self, text: torch.Tensor):
        """text bf16 [n, 77, cross] -> per-attention K/V views (step-invariant)."""
        n = text.shape[0]
        t2 = text.reshape(n * text.shape[1], text.shape[2])
        kv = {"down": [[a.text_kv(t2, n) for a in blk["attns"]] for blk in self.down], "mid": self.mid_attn.text_kv(t2, n)}
        return kv

    def run_encoder(self, x0: torch.Tensor, temb_all, kv, skip_out: Optional[List[torch.Tensor]], add_after_conv_in=None):
        """conv_in -> down blocks.  Each produced skip j is written to skip_out[j] (a strided view into the
        consumer's concat buffer) when given.  Returns (last hidden, list of skip tensors)."""
        skips = []

        def dst(j):
            return skip_out[j] if skip_out is not None else None

        x = self.conv_in(x0, out=dst(0), residual=add_after_conv_in, beta=1.0) if add_after_conv_in is not None else self.conv_in(x0, out=dst(0))
        skips.append(x)
        for bi, blk in enumerate(self.down):
            for j, r in enumerate(blk["resnets"]):
                has_attn = len(blk["attns"]) > 0
                x = r(x, temb_all, out=None if has_attn else dst(len(skips)))
                if has_attn:
                    x = blk["attns"][j](x, kv["down"][bi][j], out=dst(len(skips)))
                skips.append(x)
            if blk["down"] is not None:
                x = blk["down"](x, out=dst(len(skips)))
                skips.append(x)
        return x, skips

    def run_mid(self, x, temb_all, kv, out=None):
        x = self.mid_res[0](x, temb_all)
        x = self.mid_attn(x, kv["mid"])
        return self.mid_res[1](x, temb_all, out=out)


class UNet(_EncoderBase):
    def __init__(self, sd: SD, cfg, dev):
        self._build_encoder(sd, dev, cfg)
        g, eps = cfg.norm_num_groups, cfg.norm_eps
        rev_heads = list(reversed(cfg.num_attention_heads))
        self.up = []
        nlev = len(cfg.block_out_channels)
        for i, t in enumerate(cfg.up_block_types):
            blk = {"resnets": [], "attns": [], "up": None}
            for j in range(cfg.layers_per_block + 1):
                blk["resnets"].append(ResnetBlock(sd, f"up_blocks.{i}.resnets.{j}", dev, g, eps, self.temb_registry))
                if t.startswith("CrossAttn"):
                    blk["attns"].append(Transformer2D(sd, f"up_blocks.{i}.attentions.{j}", dev, rev_heads[i], g))
            if i != nlev - 1:
                blk["up"] = Upsample(sd, f"up_blocks.{i}.upsamplers.0", dev)
            self.up.append(blk)
        self.norm_out = Norm(sd, "conv_norm_out", dev)
        self.conv_out = Conv(sd, "conv_out", dev)
        self._finish_temb(dev)
        # concat plan: up resnet r consumes skip[-1-r]; hidden channels = resnet input - skip channels
        sk = self.skip_channels()
        self.cat_plan = []  # per up resnet: (hidden_ch, skip_ch, level)
        r = 0
        for blk in self.up:
            for res in blk["resnets"]:
                sc, lvl = sk[len(sk) - 1 - r]
                self.cat_plan.append((res.conv1.cin - sc, sc, lvl))
                r += 1

    def text_kv(self, text):
        kv = super().text_kv(text)
        n = text.shape[0]
        t2 = text.reshape(n * text.shape[1], text.shape[2])
        kv["up"] = [[a.text_kv(t2, n) for a in blk["attns"]] for blk in self.up]
        return kv

    def alloc_cat(self, n: int, h: int, w: int):
        """Concat buffers of the up path; returns (buffers, skip destination views, hidden destination views)."""
        bufs, skip_dst, hid_dst = [], [None] * len(self.cat_plan), []
        nsk = len(self.cat_plan)
        for r, (hc, sc, lvl) in enumerate(self.cat_plan):
            b = torch.empty((n, h >> lvl, w >> lvl, hc + sc), dtype=BF16, device=self.dev)
            bufs.append(b)
            hid_dst.append(b[..., :hc])
            skip_dst[nsk - 1 - r] = b[..., hc:]
        return bufs, skip_dst, hid_dst

    def encode(self, x0, temb_all, kv, n, h, w):
        """UNet encoder + mid block.  Returns state consumed by ``decode`` (and by ControlNet.inject)."""
        bufs, skip_dst, hid_dst = self.alloc_cat(n, h, w)
        x, skips = self.run_encoder(x0, temb_all, kv, skip_dst)
        mid = self.run_mid(x, temb_all, kv, out=hid_dst[0])
        return {"bufs": bufs, "skips": skips, "hid": hid_dst, "mid": mid}

    def decode(self, st, temb_all, kv) -> torch.Tensor:
        """Up path + conv_out.  Returns eps fp32 NHWC [n,h,w,out_channels]."""
        bufs, hid = st["bufs"], st["hid"]
        r = 0
        x = None
        for bi, blk in enumerate(self.up):
            for j, res in enumerate(blk["resnets"]):
                last_in_model = (r == len(self.cat_plan) - 1)
                has_attn = len(blk["attns"]) > 0
                last_in_block = j == len(blk["resnets"]) - 1
                nxt = None if (last_in_model or (last_in_block and blk["up"] is not None)) else hid[r + 1]
                x = res(bufs[r], temb_all, out=None if has_attn else nxt)
                if has_attn:
                    x = blk["attns"][j](x, kv["up"][bi][j], out=nxt)
                r += 1
            if blk["up"] is not None:
                x = blk["up"](x, out=hid[r])
        n, h, w, c = x.shape
        a = ops.groupnorm(pix3d(x), self.cfg.norm_num_groups, self.cfg.norm_eps, self.norm_out.g, self.norm_out.b, ACT_SILU).view(n, h, w, c)
        return self.conv_out(a, out_fp32=True)


class ControlNet(_EncoderBase):
    def __init__(self, sd: SD, cfg, dev):
        self._build_encoder(sd, dev, cfg)
        self._finish_temb(dev)
        p = "controlnet_cond_embedding"
        self.ce_in = Conv(sd, p + ".conv_in", dev)
        self.ce_blocks = []
        i = 0
        while f"{p}.blocks.{i}.weight" in sd:
            self.ce_blocks.append(Conv(sd, f"{p}.blocks.{i}", dev, stride=2 if i % 2 == 1 else 1))
            i += 1
        self.ce_out = Conv(sd, p + ".conv_out", dev)
        self.zero = []
        i = 0
        while f"controlnet_down_blocks.{i}.weight" in sd:
            self.zero.append(Conv(sd, f"controlnet_down_blocks.{i}", dev))
            i += 1
        self.zero_mid = Conv(sd, "controlnet_mid_block", dev)

    def cond_embedding(self, cond: torch.Tensor) -> torch.Tensor:
        """cond bf16 NHWC [n,H,W,3] in {0,1} -> [n,H/8,W/8,C0]; step-invariant, computed once per image."""
        x = self.ce_in(cond, act=ACT_SILU)
        for c in self.ce_blocks:
            x = c(x, act=ACT_SILU)
        return self.ce_out(x)

    def trunk(self, x0, temb_all, kv, cond_emb):
        """The ControlNet's own encoder + mid block -> (mid hidden, per-skip hiddens); independent of the UNet until ``accumulate``."""
        x, outs = self.run_encoder(x0, temb_all, kv, None, add_after_conv_in=cond_emb)
        return self.run_mid(x, temb_all, kv), outs

    def accumulate(self, x, outs, scale: float, unet_state):
        """Every zero-conv accumulates `scale * residual` into the UNet's skip tensors / mid output in its epilogue (diffusers:
        down_block_additional_residuals / mid_block_additional_residual)."""
        for conv, o, dst in zip(self.zero, outs, unet_state["skips"]):
            conv(o, out=dst, alpha=scale, residual=dst, beta=1.0)
        m = unet_state["mid"]
        self.zero_mid(x, out=m, alpha=scale, residual=m, beta=1.0)

    def inject(self, x0, temb_all, kv, cond_emb, scale: float, unet_state):
        """ControlNet forward + residual injection on the current stream."""
        x, outs = self.trunk(x0, temb_all, kv, cond_emb)
        self.accumulate(x, outs, scale, unet_state)


# ------------------------------------------------------------------------------------------------
# VAE
# ------------------------------------------------------------------------------------------------
class VAEAttention:
    """Single-head d = C attention of the VAE mid block: GEMM -> row softmax -> GEMM per image."""

    def __init__(self, sd: SD, p: str, dev, groups: int):
        self.groups = groups
        self.norm = Norm(sd, p + ".group_norm", dev)
        self.wqkv = _bf(torch.cat([sd[p + ".to_q.weight"], sd[p + ".to_k.weight"], sd[p + ".to_v.weight"]], 0), dev)
        self.bqkv = _f32(torch.cat([sd[p + ".to_q.bias"], sd[p + ".to_k.bias"], sd[p + ".to_v.bias"]], 0), dev)
        self.out = Linear(sd, p + ".to_out.0", dev)

    def __call__(self, x: torch.Tensor) -> torch.Tensor:
        n, h, w, c = x.shape
        t = h * w
        g = ops.groupnorm(pix3d(x), self.groups, 1e-6, self.norm.g, self.norm.b, ACT_NONE)
        qkv = ops.gemm(g.view(n * t, c), self.wqkv, bias=self.bqkv).view(n, t, 3 * c)
        o = torch.empty((n, t, c), dtype=BF16, device=x.device)
        scale = c ** -0.5
        if c <= 160:
            ops.attention(qkv[..., :c], qkv[..., c : 2 * c], qkv[..., 2 * c :], 1, out=o)
        else:
            vt = ops.transpose(qkv[..., 2 * c :])  # [n, c, t]
            for i in range(n):
                s = ops.gemm(qkv[i, :, :c], qkv[i, :, c : 2 * c])  # [t, t]
                ops.softmax_rows(s, scale, out=s)
                ops.gemm(s, vt[i], out=o[i])
        y = torch.empty_like(x) if x.is_contiguous() else torch.empty((n, h, w, c), dtype=BF16, device=x.device)
        self.out(o.view(n * t, c), out=pix2d(y), residual=pix2d(x), beta=1.0)
        return y


class _VAEStage:
    def _mid(self, sd, p, dev, groups):
        return ([ResnetBlock(sd, p + ".resnets.0", dev, groups, 1e-6, None), ResnetBlock(sd, p + ".resnets.1", dev, groups, 1e-6, None)],
                VAEAttention(sd, p + ".attentions.0", dev, groups))


class VAEDecoder(_VAEStage):
    def __init__(self, sd: SD, cfg, dev):
        g = cfg.norm_num_groups
        self.cfg = cfg
        self.post_quant = Conv(sd, "post_quant_conv", dev)
        self.conv_in = Conv(sd, "decoder.conv_in", dev)
        self.mid_res, self.mid_attn = self._mid(sd, "decoder.mid_block", dev, g)
        self.blocks = []
        for i in range(len(cfg.block_out_channels)):
            res = [ResnetBlock(sd, f"decoder.up_blocks.{i}.resnets.{j}", dev, g, 1e-6, None) for j in range(cfg.layers_per_block + 1)]
            up = Upsample(sd, f"decoder.up_blocks.{i}.upsamplers.0", dev) if f"decoder.up_blocks.{i}.upsamplers.0.conv.weight" in sd else None
            self.blocks.append((res, up))
        self.norm_out = Norm(sd, "decoder.conv_norm_out", dev)
        self.conv_out = Conv(sd, "decoder.conv_out", dev)

    def __call__(self, z: torch.Tensor) -> torch.Tensor:
        """z bf16 NHWC [n,h,w,4] (latents / scaling_factor) -> fp32 NHWC [n,8h,8w,3]."""
        x = self.conv_in(self.post_quant(z))
        x = self.mid_res[1](self.mid_attn(self.mid_res[0](x, None)), None)
        for res, up in self.blocks:
            for r in res:
                x = r(x, None)
            if up is not None:
                x = up(x)
        n, h, w, c = x.shape
        a = ops.groupnorm(pix3d(x), self.cfg.norm_num_groups, 1e-6, self.norm_out.g, self.norm_out.b, ACT_SILU).view(n, h, w, c)
        return self.conv_out(a, out_fp32=True)


class VAEEncoder(_VAEStage):
    def __init__(self, sd: SD, cfg, dev):
        g = cfg.norm_num_groups
        self.cfg = cfg
        self.conv_in = Conv(sd, "encoder.conv_in", dev)
        self.blocks = []
        for i in range(len(cfg.block_out_channels)):
            res = [ResnetBlock(sd, f"encoder.down_blocks.{i}.resnets.{j}", dev, g, 1e-6, None) for j in range(cfg.layers_per_block)]
            dn = Downsample(sd, f"encoder.down_blocks.{i}.downsamplers.0", dev, padding=0) if f"encoder.down_blocks.{i}.downsamplers.0.conv.weight" in sd else None
            self.blocks.append((res, dn))
        self.mid_res, self.mid_attn = self._mid(sd, "encoder.mid_block", dev, g)
        self.norm_out = Norm(sd, "encoder.conv_norm_out", dev)
        self.conv_out = Conv(sd, "encoder.conv_out", dev)
        self.quant = Conv(sd, "quant_conv", dev)

    def __call__(self, img: torch.Tensor) -> torch.Tensor:
        """img bf16 NHWC [n,H,W,3] in [-1,1] -> moments fp32 NHWC [n,H/8,W/8,8] (mean 0:4, logvar 4:8)."""
        x = self.conv_in(img)
        for res, dn in self.blocks:
            for r in res:
                x = r(x, None)
            if dn is not None:
                x = dn(x)
        x = self.mid_res[1](self.mid_attn(self.mid_res[0](x, None)), None)
        n, h, w, c = x.shape
        a = ops.groupnorm(pix3d(x), self.cfg.norm_num_groups, 1e-6, self.norm_out.g, self.norm_out.b, ACT_SILU).view(n, h, w, c)
        return self.quant(self.conv_out(a), out_fp32=True)


# ------------------------------------------------------------------------------------------------
# CLIP text encoder (transformers CLIPTextModel: causal pre-LN transformer, quick_gelu for ViT-L/14)
# ------------------------------------------------------------------------------------------------
class CLIPTextEncoder:
    def __init__(self, sd: SD, dev, heads: int, act: str = "quick_gelu", eps: float = 1e-5):
        p = "text_model."
        self.dev, self.heads, self.eps = dev, heads, eps
        self.act = ACT_QUICKGELU if act == "quick_gelu" else ops.ACT_GELU
        self.tok = _bf(sd[p + "embeddings.token_embedding.weight"], dev)
        self.pos = _bf(sd[p + "embeddings.position_embedding.weight"], dev)
        self.layers = []
        i = 0
        while f"{p}encoder.layers.{i}.layer_norm1.weight" in sd:
            q = f"{p}encoder.layers.{i}."
            wqkv = torch.cat([sd[q + "self_attn.q_proj.weight"], sd[q + "self_attn.k_proj.weight"], sd[q + "self_attn.v_proj.weight"]], 0)
            bqkv = torch.cat([sd[q + "self_attn.q_proj.bias"], sd[q + "self_attn.k_proj.bias"], sd[q + "self_attn.v_proj.bias"]], 0)
            self.layers.append({
                "ln1": Norm(sd, q + "layer_norm1", dev), "ln2": Norm(sd, q + "layer_norm2", dev),
                "qkv": Linear(sd, "", dev, weight=wqkv, bias=bqkv), "o": Linear(sd, q + "self_attn.out_proj", dev),
                "fc1": Linear(sd, q + "mlp.fc1", dev), "fc2": Linear(sd, q + "mlp.fc2", dev)})
            i += 1
        self.ln_final = Norm(sd, p + "final_layer_norm", dev)
        # CLIPTextModelWithProjection (SDXL's second encoder): text_embeds = text_projection(final_ln(last)[eos])
        self.proj = Linear(sd, "", dev, weight=sd["text_projection.weight"], bias=None) if "text_projection.weight" in sd else None

    def __call__(self, ids: torch.Tensor, penultimate: bool = False, pooled: bool = False, ctx_embeddings: Optional[torch.Tensor] = None,
                 ctx_begin_pos: int = 2):
        """ids int64 [n, 77] (device) -> last_hidden_state bf16 [n, 77, width].
        ctx_embeddings bf16 [n, q, width] (BLIP-Diffusion's ContextCLIPTextModel, diffusers modeling_ctx_clip.py): inserted after
        ``ctx_begin_pos`` token embeddings; ids are then [n, 77 - q] and positions run over the spliced sequence.
        penultimate=True returns hidden_states[-2] instead (the input of the last layer, no final LayerNorm: what the SDXL
        pipelines feed the UNet); pooled=True additionally returns text_embeds bf16 [n, proj] (EOS token = argmax id)."""
        n, t = ids.shape
        c = self.tok.shape[1]
        # embedding gather is index plumbing (no arithmetic): torch indexing, then our add kernel
        e = self.tok.index_select(0, ids.reshape(-1))
        if ctx_embeddings is not None:  # splice = data placement
            q = ctx_embeddings.shape[1]
            e3 = e.view(n, t, c)
            e = torch.cat([e3[:, :ctx_begin_pos], ctx_embeddings.to(BF16), e3[:, ctx_begin_pos:]], dim=1).reshape(n * (t + q), c)
            t = t + q
        pos = self.pos[:t].repeat(n, 1)
        h = ops.add(e, pos)
        pen = None
        for li, L in enumerate(self.layers):
            if penultimate and li == len(self.layers) - 1:
                pen = h.clone().view(n, t, c)  # placement: the residual stream is updated in place below
            y = ops.layernorm(h, self.eps, L["ln1"].g, L["ln1"].b)
            qkv = L["qkv"](y).view(n, t, 3 * c)
            a = ops.attention(qkv[..., :c], qkv[..., c : 2 * c], qkv[..., 2 * c :], self.heads, causal=True)
            L["o"](a.view(n * t, c), out=h, residual=h, beta=1.0)
            y = ops.layernorm(h, self.eps, L["ln2"].g, L["ln2"].b)
            f = L["fc1"](y, act=self.act)
            L["fc2"](f, out=h, residual=h, beta=1.0)
        last = ops.layernorm(h, self.eps, self.ln_final.g, self.ln_final.b).view(n, t, c)
        out = pen if penultimate else last
        if not pooled:
            return out
        eos = ids.argmax(dim=-1)  # index plumbing (transformers: input_ids.argmax(-1) for the legacy eos id)
        rows = last.reshape(n * t, c).index_select(0, torch.arange(n, device=ids.device) * t + eos)
        return out, (self.proj(rows) if self.proj is not None else rows)


# ------------------------------------------------------------------------------------------------
# ViT encoder (BLIP-2 vision tower of the Q-Former; CLIP ViT-L/14 image tower of the filter)
# ------------------------------------------------------------------------------------------------
class ViTEncoder:
    """Pre-LayerNorm ViT: patch conv (stride = kernel, no bias) + class token + learned positions -> pre-LN -> L x
    [LN, fused-QKV attention, +; LN, MLP, +] -> (post-LN).  Built from already-gathered tensors so one class serves the
    BLIP-2 key layout (``from_blip2``) and the CLIP vision layouts (filter_nets)."""

    def __init__(self, dev, *, patch_w, class_emb, pos_emb, pre_ln, layers, post_ln, heads: int, act: int, eps: float):
        self.dev, self.heads, self.act, self.eps = dev, heads, act, eps
        self.width, _, self.patch, _ = patch_w.shape
        self.patch_conv = Conv({}, "", dev, stride=self.patch, padding=0, weight=patch_w, bias=None)
        # token 0 = class embedding + position 0 (constant row); patch rows get position 1.. in the GEMM epilogue
        self.cls_row = _bf(class_emb.reshape(1, -1).float() + pos_emb.reshape(-1, self.width)[:1].float(), dev)
        self.pos_patches = _bf(pos_emb.reshape(-1, self.width)[1:], dev)
        self.pre_ln, self.post_ln, self.layers = pre_ln, post_ln, layers

    @staticmethod
    def _layer(sd, dev, wqkv, bqkv, o, ln1, ln2, fc1, fc2):
        return {"ln1": Norm(sd, ln1, dev), "ln2": Norm(sd, ln2, dev), "qkv": Linear(sd, "", dev, weight=wqkv, bias=bqkv), "o": Linear(sd, o, dev),
                "fc1": Linear(sd, fc1, dev), "fc2": Linear(sd, fc2, dev)}

    @classmethod
    def from_blip2(cls, sd: SD, p: str, dev, heads: int, eps: float = 1e-5):
        """diffusers modeling_blip2.Blip2VisionModel keys under prefix ``p`` (e.g. "visual_encoder.")."""
        layers = []
        i = 0
        while f"{p}encoder.layers.{i}.layer_norm1.weight" in sd:
            q = f"{p}encoder.layers.{i}."
            layers.append(cls._layer(sd, dev, sd[q + "self_attn.qkv.weight"], sd[q + "self_attn.qkv.bias"], q + "self_attn.projection", q + "layer_norm1",
                                     q + "layer_norm2", q + "mlp.fc1", q + "mlp.fc2"))
            i += 1
        return cls(dev, patch_w=sd[p + "embeddings.patch_embedding.weight"], class_emb=sd[p + "embeddings.class_embedding"],
                   pos_emb=sd[p + "embeddings.position_embedding"], pre_ln=Norm(sd, p + "pre_layernorm", dev), layers=layers,
                   post_ln=Norm(sd, p + "post_layernorm", dev), heads=heads, act=ACT_QUICKGELU, eps=eps)

    def __call__(self, x: torch.Tensor, apply_post_ln: bool = True) -> torch.Tensor:
        """x bf16 NHWC [n, S, S, 3] (normalised pixels) -> hidden states bf16 [n, 1 + (S/patch)^2, width]."""
        n = x.shape[0]
        c = self.width
        pt = self.patch_conv(x)  # [n, g, g, width]
        g2 = pt.shape[1] * pt.shape[2]
        assert g2 == self.pos_patches.shape[0], "image size does not match the position table"
        t = g2 + 1
        h3 = torch.empty((n, t, c), dtype=BF16, device=x.device)
        h3[:, 0] = self.cls_row  # placement
        h3[:, 1:] = ops.add(pt.view(n * g2, c), self.pos_patches.repeat(n, 1)).view(n, g2, c)
        h = ops.layernorm(h3.view(n * t, c), self.eps, self.pre_ln.g, self.pre_ln.b) if self.pre_ln is not None else h3.view(n * t, c)
        for L in self.layers:
            y = ops.layernorm(h, self.eps, L["ln1"].g, L["ln1"].b)
            qkv = L["qkv"](y).view(n, t, 3 * c)
            a = ops.attention(qkv[..., :c], qkv[..., c : 2 * c], qkv[..., 2 * c :], self.heads)
            L["o"](a.view(n * t, c), out=h, residual=h, beta=1.0)
            y = ops.layernorm(h, self.eps, L["ln2"].g, L["ln2"].b)
            L["fc2"](L["fc1"](y, act=self.act), out=h, residual=h, beta=1.0)
        if apply_post_ln and self.post_ln is not None:
            h = ops.layernorm(h, self.eps, self.post_ln.g, self.post_ln.b)
        return h.view(n, t, c)


# ------------------------------------------------------------------------------------------------
# BLIP-2 Q-Former (diffusers pipelines/blip_diffusion/modeling_blip2.py)
# ------------------------------------------------------------------------------------------------
class _BertAttn:
    """BERT attention block, post-LayerNorm: LN(dense(attn(x, ctx)) + x).  q/k/v fused where they share an input."""

    def __init__(self, sd: SD, p: str, dev, heads: int, eps: float, cross: bool):
        self.heads, self.eps, self.cross = heads, eps, cross
        a = p + ".attention."
        if cross:
            self.q = Linear(sd, a + "query", dev)
            self.kv = Linear(sd, "", dev, weight=torch.cat([sd[a + "key.weight"], sd[a + "value.weight"]], 0),
                             bias=torch.cat([sd[a + "key.bias"], sd[a + "value.bias"]], 0))
        else:
            self.qkv = Linear(sd, "", dev, weight=torch.cat([sd[a + "query.weight"], sd[a + "key.weight"], sd[a + "value.weight"]], 0),
                              bias=torch.cat([sd[a + "query.bias"], sd[a + "key.bias"], sd[a + "value.bias"]], 0))
        self.o = Linear(sd, p + ".output.dense", dev)
        self.ln = Norm(sd, p + ".output.LayerNorm", dev)

    def __call__(self, x2: torch.Tensor, n: int, ctx2: Optional[torch.Tensor] = None) -> torch.Tensor:
        rows, c = x2.shape
        t = rows // n
        if self.cross:
            q = self.q(x2).view(n, t, c)
            kv = self.kv(ctx2).view(n, ctx2.shape[0] // n, 2 * c)
            a = ops.attention(q, kv[..., :c], kv[..., c:], self.heads)
        else:
            qkv = self.qkv(x2).view(n, t, 3 * c)
            a = ops.attention(qkv[..., :c], qkv[..., c : 2 * c], qkv[..., 2 * c :], self.heads)
        y = self.o(a.view(rows, c), residual=x2, beta=1.0)
        return ops.layernorm(y, self.eps, self.ln.g, self.ln.b)


class _BertFFN:
    def __init__(self, sd: SD, p_int: str, p_out: str, dev, eps: float):
        self.fc1, self.fc2, self.ln, self.eps = Linear(sd, p_int + ".dense", dev), Linear(sd, p_out + ".dense", dev), Norm(sd, p_out + ".LayerNorm", dev), eps

    def __call__(self, x2):
        y = self.fc2(self.fc1(x2, act=ops.ACT_GELU), residual=x2, beta=1.0)
        return ops.layernorm(y, self.eps, self.ln.g, self.ln.b)


class QFormer:
    """Blip2QFormerModel.forward(image_input, text_input) -> proj_layer(sequence_output[:, :num_query_tokens]): the 16 subject
    embeddings BLIP-Diffusion splices into the CLIP prompt.  Step-invariant and prompt-invariant: computed once per
    (reference image, source subject)."""

    def __init__(self, sd: SD, cfg, dev):
        self.cfg, self.dev = cfg, dev
        e = cfg.layer_norm_eps
        self.word = _bf(sd["embeddings.word_embeddings.weight"], dev)
        self.pos = _bf(sd["embeddings.position_embeddings.weight"], dev)
        self.emb_ln = Norm(sd, "embeddings.LayerNorm", dev)
        self.query_tokens = _bf(sd["query_tokens"].reshape(cfg.num_query_tokens, cfg.hidden_size), dev)
        self.vision = ViTEncoder.from_blip2(sd, "visual_encoder.", dev, cfg.vision_num_attention_heads, cfg.vision_layer_norm_eps)
        self.layers = []
        for i in range(cfg.num_hidden_layers):
            p = f"encoder.layer.{i}"
            self.layers.append({
                "attn": _BertAttn(sd, p + ".attention", dev, cfg.num_attention_heads, e, cross=False),
                "cross": _BertAttn(sd, p + ".crossattention", dev, cfg.num_attention_heads, e, cross=True) if i % cfg.cross_attention_frequency == 0 else None,
                "ffn_text": _BertFFN(sd, p + ".intermediate", p + ".output", dev, e),
                "ffn_query": _BertFFN(sd, p + ".intermediate_query", p + ".output_query", dev, e)})
        self.proj_ln = Norm(sd, "proj_layer.LayerNorm", dev)
        self.proj1, self.proj2 = Linear(sd, "proj_layer.dense1", dev), Linear(sd, "proj_layer.dense2", dev)

    def __call__(self, image: torch.Tensor, subject_ids: torch.Tensor) -> torch.Tensor:
        """image bf16 NHWC [n,S,S,3] (BlipImageProcessor-normalised), subject_ids int64 [n,L] (no padding) -> bf16 [n,16,hidden]."""
        n, L = subject_ids.shape
        c, nq = self.cfg.hidden_size, self.cfg.num_query_tokens
        t = nq + L
        img = self.vision(image)
        img2 = img.reshape(n * img.shape[1], img.shape[2])
        txt = ops.add(self.word.index_select(0, subject_ids.reshape(-1).to(self.dev)), self.pos[:L].repeat(n, 1)).view(n, L, c)
        e = torch.cat([self.query_tokens[None].expand(n, -1, -1), txt], dim=1).reshape(n * t, c)  # placement
        h = ops.layernorm(e, self.cfg.layer_norm_eps, self.emb_ln.g, self.emb_ln.b)
        for Lr in self.layers:
            a = Lr["attn"](h, n).view(n, t, c)
            q = a[:, :nq].reshape(n * nq, c)
            if Lr["cross"] is not None:
                q = Lr["cross"](q, n, img2)
            q = Lr["ffn_query"](q)
            tx = Lr["ffn_text"](a[:, nq:].reshape(n * L, c))
            h = torch.cat([q.view(n, nq, c), tx.view(n, L, c)], dim=1).reshape(n * t, c)
        q = h.view(n, t, c)[:, :nq].reshape(n * nq, c)
        y = ops.layernorm(q, 1e-12, self.proj_ln.g, self.proj_ln.b)
        return self.proj2(self.proj1(y, act=ACT_QUICKGELU), residual=q, beta=1.0).view(n, nq, c)
