"""CPU fp32 restatement of the diffusers 0.32.2 model graphs the reference drives through
``pipe(**pipe_args)`` (run_aug/run_aug.py:278; models loaded at :184-211).

TEST INFRASTRUCTURE (see oracle/__init__.py).  **Parity unpinned**: diffusers==0.32.2
(environment.yml:17) is a third-party dependency that is neither vendored under /root/reference
nor installable here (no wheel, no network), and the reference holds no test or golden vector
at this boundary.  What is restated is the published architecture of
  models/unets/unet_2d_condition.py, unet_2d_blocks.py, models/controlnets/controlnet.py,
  models/resnet.py, attention.py, attention_processor.py, embeddings.py, downsampling.py,
  upsampling.py, activations.py, transformers/transformer_2d.py, autoencoders/{autoencoder_kl,vae}.py
with the public checkpoint configs of runwayml/stable-diffusion-v1-5,
lllyasviel/control_v11p_sd15_canny, stabilityai/sdxl-turbo, diffusers/controlnet-canny-sdxl-1.0
(model ids: run_aug/run_aug.py:53-72).  Module / parameter names follow the diffusers
state-dict keys so real checkpoints would load unchanged.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import List, Optional, Sequence, Tuple

import torch
import torch.nn as nn
import torch.nn.functional as F


# ------------------------------------------------------------------------------------------------
# configs
# ------------------------------------------------------------------------------------------------
@dataclass
class UNetConfig:
    in_channels: int = 4
    out_channels: int = 4
    block_out_channels: Tuple[int, ...] = (320, 640, 1280, 1280)
    down_block_types: Tuple[str, ...] = ("CrossAttnDownBlock2D", "CrossAttnDownBlock2D", "CrossAttnDownBlock2D", "DownBlock2D")
    up_block_types: Tuple[str, ...] = ("UpBlock2D", "CrossAttnUpBlock2D", "CrossAttnUpBlock2D", "CrossAttnUpBlock2D")
    layers_per_block: int = 2
    transformer_layers_per_block: Tuple[int, ...] = (1, 1, 1, 1)
    num_attention_heads: Tuple[int, ...] = (8, 8, 8, 8)  # SD1.5's `attention_head_dim: 8` is really the head COUNT
    cross_attention_dim: int = 768
    norm_num_groups: int = 32
    norm_eps: float = 1e-5
    use_linear_projection: bool = False
    flip_sin_to_cos: bool = True
    freq_shift: float = 0.0
    addition_embed_type: Optional[str] = None  # "text_time" for SDXL
    addition_time_embed_dim: int = 256
    projection_class_embeddings_input_dim: int = 2816
    conditioning_embedding_out_channels: Tuple[int, ...] = (16, 32, 96, 256)  # ControlNet only

    @staticmethod
    def sd15() -> "UNetConfig":
        return UNetConfig()

    @staticmethod
    def sdxl() -> "UNetConfig":
        return UNetConfig(
            block_out_channels=(320, 640, 1280),
            down_block_types=("DownBlock2D", "CrossAttnDownBlock2D", "CrossAttnDownBlock2D"),
            up_block_types=("CrossAttnUpBlock2D", "CrossAttnUpBlock2D", "UpBlock2D"),
            transformer_layers_per_block=(1, 2, 10),
            num_attention_heads=(5, 10, 20),
            cross_attention_dim=2048,
            use_linear_projection=True,
            addition_embed_type="text_time",
        )

    @staticmethod
    def tiny(cross_attention_dim: int = 64) -> "UNetConfig":
        """Small same-topology config for fast CPU parity tests (all structural features of SD1.5)."""
        return UNetConfig(block_out_channels=(64, 128, 128, 128), num_attention_heads=(4, 4, 4, 4), cross_attention_dim=cross_attention_dim,
                          conditioning_embedding_out_channels=(16, 32, 32, 64))


@dataclass
class VAEConfig:
    in_channels: int = 3
    out_channels: int = 3
    latent_channels: int = 4
    block_out_channels: Tuple[int, ...] = (128, 256, 512, 512)
    layers_per_block: int = 2
    norm_num_groups: int = 32
    scaling_factor: float = 0.18215

    @staticmethod
    def sd15() -> "VAEConfig":
        return VAEConfig()

    @staticmethod
    def sdxl() -> "VAEConfig":
        return VAEConfig(scaling_factor=0.13025)

    @staticmethod
    def tiny() -> "VAEConfig":
        return VAEConfig(block_out_channels=(32, 64, 64, 64))


# ------------------------------------------------------------------------------------------------
# building blocks
# ------------------------------------------------------------------------------------------------
def get_timestep_embedding(timesteps: torch.Tensor, dim: int, flip_sin_to_cos: bool, downscale_freq_shift: float, max_period: int = 10000):
    half = dim // 2
    exponent = -math.log(max_period) * torch.arange(half, dtype=torch.float32, device=timesteps.device)
    exponent = exponent / (half - downscale_freq_shift)
    emb = timesteps[:, None].float() * torch.exp(exponent)[None, :]
    emb = torch.cat([torch.sin(emb), torch.cos(emb)], dim=-1)
    if flip_sin_to_cos:
        emb = torch.cat([emb[:, half:], emb[:, :half]], dim=-1)
    return emb


class TimestepEmbedding(nn.Module):
    def __init__(self, in_dim: int, dim: int):
        super().__init__()
        self.linear_1 = nn.Linear(in_dim, dim)
        self.linear_2 = nn.Linear(dim, dim)

    def forward(self, x):
        return self.linear_2(F.silu(self.linear_1(x)))


class ResnetBlock2D(nn.Module):
    def __init__(self, cin: int, cout: int, temb_channels: Optional[int], groups: int, eps: float):
        super().__init__()
        self.norm1 = nn.GroupNorm(groups, cin, eps=eps)
        self.conv1 = nn.Conv2d(cin, cout, 3, padding=1)
        self.time_emb_proj = nn.Linear(temb_channels, cout) if temb_channels else None
        self.norm2 = nn.GroupNorm(groups, cout, eps=eps)
        self.conv2 = nn.Conv2d(cout, cout, 3, padding=1)
        self.conv_shortcut = nn.Conv2d(cin, cout, 1) if cin != cout else None

    def forward(self, x, temb=None):
        h = self.conv1(F.silu(self.norm1(x)))
        if self.time_emb_proj is not None and temb is not None:
            h = h + self.time_emb_proj(F.silu(temb))[:, :, None, None]
        h = self.conv2(F.silu(self.norm2(h)))
        if self.conv_shortcut is not None:
            x = self.conv_shortcut(x)
        return x + h  # output_scale_factor = 1.0


class Attention(nn.Module):
    def __init__(self, query_dim: int, cross_dim: Optional[int], heads: int, dim_head: int, qkv_bias: bool = False):
        super().__init__()
        inner = heads * dim_head
        self.heads = heads
        self.to_q = nn.Linear(query_dim, inner, bias=qkv_bias)
        self.to_k = nn.Linear(cross_dim or query_dim, inner, bias=qkv_bias)
        self.to_v = nn.Linear(cross_dim or query_dim, inner, bias=qkv_bias)
        self.to_out = nn.ModuleList([nn.Linear(inner, query_dim), nn.Dropout(0.0)])

    def forward(self, x, context=None):
        ctx = x if context is None else context
        b, t, _ = x.shape
        q, k, v = self.to_q(x), self.to_k(ctx), self.to_v(ctx)
        h = self.heads
        q, k, v = (z.view(b, -1, h, z.shape[-1] // h).transpose(1, 2) for z in (q, k, v))
        o = F.scaled_dot_product_attention(q, k, v)
        o = o.transpose(1, 2).reshape(b, t, -1)
        return self.to_out[0](o)


class GEGLU(nn.Module):
    def __init__(self, dim: int, inner: int):
        super().__init__()
        self.proj = nn.Linear(dim, inner * 2)

    def forward(self, x):
        h, gate = self.proj(x).chunk(2, dim=-1)
        return h * F.gelu(gate)


class FeedForward(nn.Module):
    def __init__(self, dim: int):
        super().__init__()
        self.net = nn.ModuleList([GEGLU(dim, dim * 4), nn.Dropout(0.0), nn.Linear(dim * 4, dim)])

    def forward(self, x):
        return self.net[2](self.net[0](x))


class BasicTransformerBlock(nn.Module):
    def __init__(self, dim: int, heads: int, dim_head: int, cross_dim: int):
        super().__init__()
        self.norm1 = nn.LayerNorm(dim)
        self.attn1 = Attention(dim, None, heads, dim_head)
        self.norm2 = nn.LayerNorm(dim)
        self.attn2 = Attention(dim, cross_dim, heads, dim_head)
        self.norm3 = nn.LayerNorm(dim)
        self.ff = FeedForward(dim)

    def forward(self, x, context):
        x = self.attn1(self.norm1(x)) + x
        x = self.attn2(self.norm2(x), context) + x
        x = self.ff(self.norm3(x)) + x
        return x


class Transformer2DModel(nn.Module):
    def __init__(self, channels: int, heads: int, depth: int, cross_dim: int, groups: int, linear_proj: bool):
        super().__init__()
        self.linear_proj = linear_proj
        self.norm = nn.GroupNorm(groups, channels, eps=1e-6)
        if linear_proj:
            self.proj_in = nn.Linear(channels, channels)
            self.proj_out = nn.Linear(channels, channels)
        else:
            self.proj_in = nn.Conv2d(channels, channels, 1)
            self.proj_out = nn.Conv2d(channels, channels, 1)
        self.transformer_blocks = nn.ModuleList([BasicTransformerBlock(channels, heads, channels // heads, cross_dim) for _ in range(depth)])

    def forward(self, x, context):
        b, c, h, w = x.shape
        res = x
        x = self.norm(x)
        if not self.linear_proj:
            x = self.proj_in(x).permute(0, 2, 3, 1).reshape(b, h * w, c)
        else:
            x = self.proj_in(x.permute(0, 2, 3, 1).reshape(b, h * w, c))
        for blk in self.transformer_blocks:
            x = blk(x, context)
        if not self.linear_proj:
            x = self.proj_out(x.reshape(b, h, w, c).permute(0, 3, 1, 2))
        else:
            x = self.proj_out(x).reshape(b, h, w, c).permute(0, 3, 1, 2)
        return x + res


class Downsample2D(nn.Module):
    def __init__(self, channels: int, padding: int = 1):
        super().__init__()
        self.padding = padding
        self.conv = nn.Conv2d(channels, channels, 3, stride=2, padding=padding)

    def forward(self, x):
        if self.padding == 0:
            x = F.pad(x, (0, 1, 0, 1))
        return self.conv(x)


class Upsample2D(nn.Module):
    def __init__(self, channels: int):
        super().__init__()
        self.conv = nn.Conv2d(channels, channels, 3, padding=1)

    def forward(self, x):
        return self.conv(F.interpolate(x, scale_factor=2.0, mode="nearest"))


class DownBlock(nn.Module):
    def __init__(self, cin, cout, temb, layers, groups, eps, add_down, heads=None, depth=0, cross_dim=None, linear_proj=False):
        super().__init__()
        self.resnets = nn.ModuleList([ResnetBlock2D(cin if i == 0 else cout, cout, temb, groups, eps) for i in range(layers)])
        self.attentions = nn.ModuleList([Transformer2DModel(cout, heads, depth, cross_dim, groups, linear_proj) for _ in range(layers)]) if heads else None
        self.downsamplers = nn.ModuleList([Downsample2D(cout)]) if add_down else None

    def forward(self, x, temb, context):
        outs = []
        for i, r in enumerate(self.resnets):
            x = r(x, temb)
            if self.attentions is not None:
                x = self.attentions[i](x, context)
            outs.append(x)
        if self.downsamplers is not None:
            x = self.downsamplers[0](x)
            outs.append(x)
        return x, outs


class MidBlockCrossAttn(nn.Module):
    def __init__(self, ch, temb, groups, eps, heads, depth, cross_dim, linear_proj):
        super().__init__()
        self.resnets = nn.ModuleList([ResnetBlock2D(ch, ch, temb, groups, eps), ResnetBlock2D(ch, ch, temb, groups, eps)])
        self.attentions = nn.ModuleList([Transformer2DModel(ch, heads, depth, cross_dim, groups, linear_proj)])

    def forward(self, x, temb, context):
        x = self.resnets[0](x, temb)
        x = self.attentions[0](x, context)
        return self.resnets[1](x, temb)


class UpBlock(nn.Module):
    def __init__(self, cin, cout, prev, temb, layers, groups, eps, add_up, heads=None, depth=0, cross_dim=None, linear_proj=False):
        super().__init__()
        rs = []
        for i in range(layers):
            skip = cin if i == layers - 1 else cout
            rin = prev if i == 0 else cout
            rs.append(ResnetBlock2D(rin + skip, cout, temb, groups, eps))
        self.resnets = nn.ModuleList(rs)
        self.attentions = nn.ModuleList([Transformer2DModel(cout, heads, depth, cross_dim, groups, linear_proj) for _ in range(layers)]) if heads else None
        self.upsamplers = nn.ModuleList([Upsample2D(cout)]) if add_up else None

    def forward(self, x, skips: List[torch.Tensor], temb, context):
        for i, r in enumerate(self.resnets):
            x = torch.cat([x, skips.pop()], dim=1)
            x = r(x, temb)
            if self.attentions is not None:
                x = self.attentions[i](x, context)
        if self.upsamplers is not None:
            x = self.upsamplers[0](x)
        return x


class _Encoder(nn.Module):
    """conv_in + time embedding + down blocks + mid block shared by the UNet and the ControlNet."""

    def build_encoder(self, cfg: UNetConfig):
        c0 = cfg.block_out_channels[0]
        temb = c0 * 4
        self.cfg = cfg
        self.conv_in = nn.Conv2d(cfg.in_channels, c0, 3, padding=1)
        self.time_embedding = TimestepEmbedding(c0, temb)
        if cfg.addition_embed_type == "text_time":
            self.add_embedding = TimestepEmbedding(cfg.projection_class_embeddings_input_dim, temb)
        downs = []
        cout = c0
        n = len(cfg.block_out_channels)
        for i, t in enumerate(cfg.down_block_types):
            cin, cout = cout, cfg.block_out_channels[i]
            cross = t.startswith("CrossAttn")
            downs.append(DownBlock(cin, cout, temb, cfg.layers_per_block, cfg.norm_num_groups, cfg.norm_eps, i != n - 1,
                                   cfg.num_attention_heads[i] if cross else None, cfg.transformer_layers_per_block[i], cfg.cross_attention_dim,
                                   cfg.use_linear_projection))
        self.down_blocks = nn.ModuleList(downs)
        self.mid_block = MidBlockCrossAttn(cfg.block_out_channels[-1], temb, cfg.norm_num_groups, cfg.norm_eps, cfg.num_attention_heads[-1],
                                           cfg.transformer_layers_per_block[-1], cfg.cross_attention_dim, cfg.use_linear_projection)

    def time_embed(self, sample, timestep, added_cond):
        cfg = self.cfg
        t = timestep
        if not torch.is_tensor(t):
            t = torch.tensor([t], dtype=torch.float32, device=sample.device)
        t = t.to(sample.device).reshape(-1).expand(sample.shape[0]).float()
        emb = self.time_embedding(get_timestep_embedding(t, cfg.block_out_channels[0], cfg.flip_sin_to_cos, cfg.freq_shift).to(sample.dtype))
        if cfg.addition_embed_type == "text_time":
            text_embeds, time_ids = added_cond["text_embeds"], added_cond["time_ids"]
            tid = get_timestep_embedding(time_ids.flatten(), cfg.addition_time_embed_dim, cfg.flip_sin_to_cos, cfg.freq_shift)
            tid = tid.reshape(text_embeds.shape[0], -1).to(sample.dtype)
            emb = emb + self.add_embedding(torch.cat([text_embeds, tid], dim=-1))
        return emb


class UNet2DConditionModel(_Encoder):
    def __init__(self, cfg: UNetConfig):
        super().__init__()
        self.build_encoder(cfg)
        temb = cfg.block_out_channels[0] * 4
        rev = list(reversed(cfg.block_out_channels))
        rev_heads = list(reversed(cfg.num_attention_heads))
        rev_depth = list(reversed(cfg.transformer_layers_per_block))
        ups = []
        cout = rev[0]
        n = len(rev)
        for i, t in enumerate(cfg.up_block_types):
            prev, cout = cout, rev[i]
            cin = rev[min(i + 1, n - 1)]
            cross = t.startswith("CrossAttn")
            ups.append(UpBlock(cin, cout, prev, temb, cfg.layers_per_block + 1, cfg.norm_num_groups, cfg.norm_eps, i != n - 1,
                               rev_heads[i] if cross else None, rev_depth[i], cfg.cross_attention_dim, cfg.use_linear_projection))
        self.up_blocks = nn.ModuleList(ups)
        self.conv_norm_out = nn.GroupNorm(cfg.norm_num_groups, cfg.block_out_channels[0], eps=cfg.norm_eps)
        self.conv_out = nn.Conv2d(cfg.block_out_channels[0], cfg.out_channels, 3, padding=1)

    def forward(self, sample, timestep, encoder_hidden_states, down_block_additional_residuals=None, mid_block_additional_residual=None,
                added_cond_kwargs=None):
        emb = self.time_embed(sample, timestep, added_cond_kwargs)
        x = self.conv_in(sample)
        skips = [x]
        for blk in self.down_blocks:
            x, outs = blk(x, emb, encoder_hidden_states)
            skips += outs
        if down_block_additional_residuals is not None:
            skips = [s + r for s, r in zip(skips, down_block_additional_residuals)]
        x = self.mid_block(x, emb, encoder_hidden_states)
        if mid_block_additional_residual is not None:
            x = x + mid_block_additional_residual
        for blk in self.up_blocks:
            x = blk(x, skips, emb, encoder_hidden_states)
        return self.conv_out(F.silu(self.conv_norm_out(x)))


class ControlNetConditioningEmbedding(nn.Module):
    def __init__(self, out_channels: int, block_out: Sequence[int], cond_channels: int = 3):
        super().__init__()
        self.conv_in = nn.Conv2d(cond_channels, block_out[0], 3, padding=1)
        blocks = []
        for i in range(len(block_out) - 1):
            blocks.append(nn.Conv2d(block_out[i], block_out[i], 3, padding=1))
            blocks.append(nn.Conv2d(block_out[i], block_out[i + 1], 3, padding=1, stride=2))
        self.blocks = nn.ModuleList(blocks)
        self.conv_out = nn.Conv2d(block_out[-1], out_channels, 3, padding=1)

    def forward(self, c):
        x = F.silu(self.conv_in(c))
        for b in self.blocks:
            x = F.silu(b(x))
        return self.conv_out(x)


class ControlNetModel(_Encoder):
    def __init__(self, cfg: UNetConfig):
        super().__init__()
        self.build_encoder(cfg)
        self.controlnet_cond_embedding = ControlNetConditioningEmbedding(cfg.block_out_channels[0], cfg.conditioning_embedding_out_channels)
        chans = [cfg.block_out_channels[0]]
        n = len(cfg.block_out_channels)
        for i, c in enumerate(cfg.block_out_channels):
            chans += [c] * cfg.layers_per_block
            if i != n - 1:
                chans.append(c)
        self.controlnet_down_blocks = nn.ModuleList([nn.Conv2d(c, c, 1) for c in chans])
        self.controlnet_mid_block = nn.Conv2d(cfg.block_out_channels[-1], cfg.block_out_channels[-1], 1)

    def forward(self, sample, timestep, encoder_hidden_states, controlnet_cond, conditioning_scale: float = 1.0, added_cond_kwargs=None):
        emb = self.time_embed(sample, timestep, added_cond_kwargs)
        x = self.conv_in(sample) + self.controlnet_cond_embedding(controlnet_cond)
        outs = [x]
        for blk in self.down_blocks:
            x, o = blk(x, emb, encoder_hidden_states)
            outs += o
        x = self.mid_block(x, emb, encoder_hidden_states)
        down = [conv(o) * conditioning_scale for conv, o in zip(self.controlnet_down_blocks, outs)]
        mid = self.controlnet_mid_block(x) * conditioning_scale
        return down, mid


# ------------------------------------------------------------------------------------------------
# VAE
# ------------------------------------------------------------------------------------------------
class VAEAttention(nn.Module):
    """The deprecated-AttentionBlock form diffusers keeps for the VAE mid block: GroupNorm, single head,
    biased q/k/v, residual connection."""

    def __init__(self, ch: int, groups: int):
        super().__init__()
        self.group_norm = nn.GroupNorm(groups, ch, eps=1e-6)
        self.to_q = nn.Linear(ch, ch)
        self.to_k = nn.Linear(ch, ch)
        self.to_v = nn.Linear(ch, ch)
        self.to_out = nn.ModuleList([nn.Linear(ch, ch), nn.Dropout(0.0)])

    def forward(self, x):
        b, c, h, w = x.shape
        res = x
        t = self.group_norm(x.view(b, c, h * w)).transpose(1, 2)
        q, k, v = self.to_q(t), self.to_k(t), self.to_v(t)
        o = F.scaled_dot_product_attention(q[:, None], k[:, None], v[:, None])[:, 0]
        o = self.to_out[0](o).transpose(1, 2).reshape(b, c, h, w)
        return o + res


class VAEMidBlock(nn.Module):
    def __init__(self, ch: int, groups: int):
        super().__init__()
        self.resnets = nn.ModuleList([ResnetBlock2D(ch, ch, None, groups, 1e-6), ResnetBlock2D(ch, ch, None, groups, 1e-6)])
        self.attentions = nn.ModuleList([VAEAttention(ch, groups)])

    def forward(self, x):
        return self.resnets[1](self.attentions[0](self.resnets[0](x)))


class EncoderBlock(nn.Module):
    def __init__(self, cin, cout, layers, groups, add_down):
        super().__init__()
        self.resnets = nn.ModuleList([ResnetBlock2D(cin if i == 0 else cout, cout, None, groups, 1e-6) for i in range(layers)])
        self.downsamplers = nn.ModuleList([Downsample2D(cout, padding=0)]) if add_down else None

    def forward(self, x):
        for r in self.resnets:
            x = r(x)
        if self.downsamplers is not None:
            x = self.downsamplers[0](x)
        return x


class DecoderBlock(nn.Module):
    def __init__(self, cin, cout, layers, groups, add_up):
        super().__init__()
        self.resnets = nn.ModuleList([ResnetBlock2D(cin if i == 0 else cout, cout, None, groups, 1e-6) for i in range(layers)])
        self.upsamplers = nn.ModuleList([Upsample2D(cout)]) if add_up else None

    def forward(self, x):
        for r in self.resnets:
            x = r(x)
        if self.upsamplers is not None:
            x = self.upsamplers[0](x)
        return x


class Encoder(nn.Module):
    def __init__(self, cfg: VAEConfig):
        super().__init__()
        ch = cfg.block_out_channels
        self.conv_in = nn.Conv2d(cfg.in_channels, ch[0], 3, padding=1)
        blocks, cout = [], ch[0]
        for i, c in enumerate(ch):
            cin, cout = cout, c
            blocks.append(EncoderBlock(cin, cout, cfg.layers_per_block, cfg.norm_num_groups, i != len(ch) - 1))
        self.down_blocks = nn.ModuleList(blocks)
        self.mid_block = VAEMidBlock(ch[-1], cfg.norm_num_groups)
        self.conv_norm_out = nn.GroupNorm(cfg.norm_num_groups, ch[-1], eps=1e-6)
        self.conv_out = nn.Conv2d(ch[-1], 2 * cfg.latent_channels, 3, padding=1)

    def forward(self, x):
        x = self.conv_in(x)
        for b in self.down_blocks:
            x = b(x)
        x = self.mid_block(x)
        return self.conv_out(F.silu(self.conv_norm_out(x)))


class Decoder(nn.Module):
    def __init__(self, cfg: VAEConfig):
        super().__init__()
        rev = list(reversed(cfg.block_out_channels))
        self.conv_in = nn.Conv2d(cfg.latent_channels, rev[0], 3, padding=1)
        self.mid_block = VAEMidBlock(rev[0], cfg.norm_num_groups)
        blocks, cout = [], rev[0]
        for i, c in enumerate(rev):
            cin, cout = cout, c
            blocks.append(DecoderBlock(cin, cout, cfg.layers_per_block + 1, cfg.norm_num_groups, i != len(rev) - 1))
        self.up_blocks = nn.ModuleList(blocks)
        self.conv_norm_out = nn.GroupNorm(cfg.norm_num_groups, rev[-1], eps=1e-6)
        self.conv_out = nn.Conv2d(rev[-1], cfg.out_channels, 3, padding=1)

    def forward(self, z):
        x = self.mid_block(self.conv_in(z))
        for b in self.up_blocks:
            x = b(x)
        return self.conv_out(F.silu(self.conv_norm_out(x)))


class AutoencoderKL(nn.Module):
    def __init__(self, cfg: VAEConfig):
        super().__init__()
        self.cfg = cfg
        self.encoder = Encoder(cfg)
        self.decoder = Decoder(cfg)
        self.quant_conv = nn.Conv2d(2 * cfg.latent_channels, 2 * cfg.latent_channels, 1)
        self.post_quant_conv = nn.Conv2d(cfg.latent_channels, cfg.latent_channels, 1)

    def encode_moments(self, x):
        mean, logvar = self.quant_conv(self.encoder(x)).chunk(2, dim=1)
        return mean, torch.clamp(logvar, -30.0, 20.0)

    def decode(self, z):
        return self.decoder(self.post_quant_conv(z))


# ------------------------------------------------------------------------------------------------
# random init (no checkpoints exist offline)
# ------------------------------------------------------------------------------------------------
def variance_preserving_init_(model: nn.Module, seed: int, zero_conv_std: float = 0.02) -> nn.Module:
    """Deterministic fan-in (Kaiming-normal, gain 1) init; residual-branch output layers scaled by 1/sqrt(2);
    norm affine ~ N(1, .05) / N(0, .05); biases ~ N(0, .02); ControlNet zero-convs NON-zero (std 0.02) so the
    residual-injection path is exercised."""
    g = torch.Generator().manual_seed(seed)
    for name, p in sorted(model.named_parameters()):
        with torch.no_grad():
            if p.dim() >= 2:
                fan_in = p[0].numel()
                std = 1.0 / math.sqrt(fan_in)
                if any(k in name for k in ("conv2.weight", "to_out.0.weight", "ff.net.2.weight", "proj_out.weight")):
                    std *= 1.0 / math.sqrt(2.0)
                if "controlnet_down_blocks" in name or "controlnet_mid_block" in name or "controlnet_cond_embedding.conv_out" in name:
                    std = zero_conv_std
                p.copy_(torch.randn(p.shape, generator=g) * std)
            elif "norm" in name and name.endswith("weight"):
                p.copy_(1.0 + 0.05 * torch.randn(p.shape, generator=g))
            else:
                p.copy_(0.02 * torch.randn(p.shape, generator=g))
    return model
