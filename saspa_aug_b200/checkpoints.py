"""Model configs + diffusers/transformers-keyed state dicts.

The reference builds its pipelines with ``from_pretrained`` of public HF repos (run_aug/run_aug.py:53-72,
:184-211).  No checkpoints, tokenizer vocabularies or network exist in this environment, so benchmarks and
parity tests use deterministic, variance-preserving RANDOM weights with the exact key names / shapes of
those checkpoints; a real ``*.safetensors`` state dict can be passed to the same
model constructors unchanged.  ``*_shapes`` enumerate (key, shape) from the config; tests load the generated
dict into the oracle modules with ``strict=True``, which cross-checks the two enumerations.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Dict, Iterator, Optional, Tuple

import torch


@dataclass
class UNetConfig:
    in_channels: int = 4
    out_channels: int = 4
    block_out_channels: Tuple[int, ...] = (320, 640, 1280, 1280)
    down_block_types: Tuple[str, ...] = ("CrossAttnDownBlock2D", "CrossAttnDownBlock2D", "CrossAttnDownBlock2D", "DownBlock2D")
    up_block_types: Tuple[str, ...] = ("UpBlock2D", "CrossAttnUpBlock2D", "CrossAttnUpBlock2D", "CrossAttnUpBlock2D")
    layers_per_block: int = 2
    transformer_layers_per_block: Tuple[int, ...] = (1, 1, 1, 1)
    num_attention_heads: Tuple[int, ...] = (8, 8, 8, 8)
    cross_attention_dim: int = 768
    norm_num_groups: int = 32
    norm_eps: float = 1e-5
    use_linear_projection: bool = False
    flip_sin_to_cos: bool = True
    freq_shift: float = 0.0
    addition_embed_type: Optional[str] = None
    addition_time_embed_dim: int = 256
    projection_class_embeddings_input_dim: int = 2816
    conditioning_embedding_out_channels: Tuple[int, ...] = (16, 32, 96, 256)

    @staticmethod
    def sd15() -> "UNetConfig":  # runwayml/stable-diffusion-v1-5, lllyasviel/control_v11p_sd15_canny
        return UNetConfig()

    @staticmethod
    def sdxl() -> "UNetConfig":  # stabilityai/sdxl-turbo, diffusers/controlnet-canny-sdxl-1.0
        return UNetConfig(block_out_channels=(320, 640, 1280), down_block_types=("DownBlock2D", "CrossAttnDownBlock2D", "CrossAttnDownBlock2D"),
                          up_block_types=("CrossAttnUpBlock2D", "CrossAttnUpBlock2D", "UpBlock2D"), transformer_layers_per_block=(1, 2, 10),
                          num_attention_heads=(5, 10, 20), cross_attention_dim=2048, use_linear_projection=True, addition_embed_type="text_time")

    @staticmethod
    def tiny_xl() -> "UNetConfig":  # SDXL topology at toy width: cross dim 64 + 128 (two text encoders), pooled 96 + 6 x 32 time ids
        return UNetConfig(block_out_channels=(64, 128, 128), down_block_types=("DownBlock2D", "CrossAttnDownBlock2D", "CrossAttnDownBlock2D"),
                          up_block_types=("CrossAttnUpBlock2D", "CrossAttnUpBlock2D", "UpBlock2D"), transformer_layers_per_block=(1, 2, 3),
                          num_attention_heads=(1, 2, 2), cross_attention_dim=192, use_linear_projection=True, addition_embed_type="text_time",
                          addition_time_embed_dim=32, projection_class_embeddings_input_dim=96 + 6 * 32,
                          conditioning_embedding_out_channels=(16, 32, 32, 64))

    @staticmethod
    def tiny(cross_attention_dim: int = 64) -> "UNetConfig":
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
    def sdxl() -> "VAEConfig":  # madebyollin/sdxl-vae-fp16-fix (run_aug/run_aug.py:189)
        return VAEConfig(scaling_factor=0.13025)

    @staticmethod
    def tiny() -> "VAEConfig":
        return VAEConfig(block_out_channels=(32, 64, 64, 64))


@dataclass
class CLIPTextConfig:
    vocab_size: int = 49408
    hidden_size: int = 768
    intermediate_size: int = 3072
    num_hidden_layers: int = 12
    num_attention_heads: int = 12
    max_position_embeddings: int = 77
    hidden_act: str = "quick_gelu"
    layer_norm_eps: float = 1e-5
    projection_dim: int = 0  # > 0: CLIPTextModelWithProjection (text_projection.weight [proj, hidden], no bias)

    @staticmethod
    def sd15() -> "CLIPTextConfig":  # openai/clip-vit-large-patch14 text tower
        return CLIPTextConfig()

    @staticmethod
    def sdxl_g() -> "CLIPTextConfig":  # SDXL text_encoder_2: OpenCLIP ViT-bigG/14 text tower (CLIPTextModelWithProjection)
        return CLIPTextConfig(hidden_size=1280, intermediate_size=5120, num_hidden_layers=32, num_attention_heads=20, hidden_act="gelu",
                              projection_dim=1280)

    @staticmethod
    def tiny() -> "CLIPTextConfig":
        return CLIPTextConfig(vocab_size=1000, hidden_size=64, intermediate_size=256, num_hidden_layers=2, num_attention_heads=4)

    @staticmethod
    def tiny_g() -> "CLIPTextConfig":
        return CLIPTextConfig(vocab_size=1000, hidden_size=128, intermediate_size=256, num_hidden_layers=3, num_attention_heads=2, hidden_act="gelu",
                              projection_dim=96)


@dataclass
class Blip2Config:
    """Salesforce/blipdiffusion(-controlnet) ``qformer/config.json`` (diffusers pipelines/blip_diffusion/modeling_blip2.py): BERT-style
    Q-Former with ``num_query_tokens`` learned queries, cross-attending every ``cross_attention_frequency``-th layer to its own ViT."""
    vocab_size: int = 30523
    hidden_size: int = 768
    num_hidden_layers: int = 12
    num_attention_heads: int = 12
    intermediate_size: int = 3072
    max_position_embeddings: int = 512
    layer_norm_eps: float = 1e-12
    cross_attention_frequency: int = 2
    num_query_tokens: int = 16
    vision_hidden_size: int = 1024
    vision_intermediate_size: int = 4096
    vision_num_hidden_layers: int = 23
    vision_num_attention_heads: int = 16
    image_size: int = 224
    patch_size: int = 14
    vision_layer_norm_eps: float = 1e-5

    @staticmethod
    def blipdiffusion() -> "Blip2Config":
        return Blip2Config()

    @staticmethod
    def tiny() -> "Blip2Config":
        return Blip2Config(vocab_size=500, hidden_size=64, num_hidden_layers=4, num_attention_heads=2, intermediate_size=128, max_position_embeddings=32,
                           num_query_tokens=16, vision_hidden_size=128, vision_intermediate_size=256, vision_num_hidden_layers=3,
                           vision_num_attention_heads=2, image_size=56, patch_size=14)


Shapes = Iterator[Tuple[str, Tuple[int, ...]]]


def _conv(p, cin, cout, k=3) -> Shapes:
    yield p + ".weight", (cout, cin, k, k)
    yield p + ".bias", (cout,)


def _lin(p, cin, cout, bias=True) -> Shapes:
    yield p + ".weight", (cout, cin)
    if bias:
        yield p + ".bias", (cout,)


def _norm(p, c) -> Shapes:
    yield p + ".weight", (c,)
    yield p + ".bias", (c,)


def _resnet(p, cin, cout, temb) -> Shapes:
    yield from _norm(p + ".norm1", cin)
    yield from _conv(p + ".conv1", cin, cout)
    if temb:
        yield from _lin(p + ".time_emb_proj", temb, cout)
    yield from _norm(p + ".norm2", cout)
    yield from _conv(p + ".conv2", cout, cout)
    if cin != cout:
        yield from _conv(p + ".conv_shortcut", cin, cout, 1)


def _transformer(p, c, depth, cross, linear) -> Shapes:
    yield from _norm(p + ".norm", c)
    if linear:
        yield from _lin(p + ".proj_in", c, c)
    else:
        yield from _conv(p + ".proj_in", c, c, 1)
    for i in range(depth):
        q = f"{p}.transformer_blocks.{i}"
        yield from _norm(q + ".norm1", c)
        for nm, kin in (("attn1", c), ("attn2", cross)):
            yield from _lin(f"{q}.{nm}.to_q", c, c, False)
            yield from _lin(f"{q}.{nm}.to_k", kin, c, False)
            yield from _lin(f"{q}.{nm}.to_v", kin, c, False)
            yield from _lin(f"{q}.{nm}.to_out.0", c, c)
            if nm == "attn1":
                yield from _norm(q + ".norm2", c)
        yield from _norm(q + ".norm3", c)
        yield from _lin(q + ".ff.net.0.proj", c, 8 * c)
        yield from _lin(q + ".ff.net.2", 4 * c, c)
    if linear:
        yield from _lin(p + ".proj_out", c, c)
    else:
        yield from _conv(p + ".proj_out", c, c, 1)


def _encoder_shapes(cfg: UNetConfig) -> Shapes:
    c0 = cfg.block_out_channels[0]
    temb = 4 * c0
    yield from _conv("conv_in", cfg.in_channels, c0)
    yield from _lin("time_embedding.linear_1", c0, temb)
    yield from _lin("time_embedding.linear_2", temb, temb)
    if cfg.addition_embed_type == "text_time":
        yield from _lin("add_embedding.linear_1", cfg.projection_class_embeddings_input_dim, temb)
        yield from _lin("add_embedding.linear_2", temb, temb)
    cout = c0
    n = len(cfg.block_out_channels)
    for i, t in enumerate(cfg.down_block_types):
        cin, cout = cout, cfg.block_out_channels[i]
        for j in range(cfg.layers_per_block):
            yield from _resnet(f"down_blocks.{i}.resnets.{j}", cin if j == 0 else cout, cout, temb)
            if t.startswith("CrossAttn"):
                yield from _transformer(f"down_blocks.{i}.attentions.{j}", cout, cfg.transformer_layers_per_block[i], cfg.cross_attention_dim,
                                        cfg.use_linear_projection)
        if i != n - 1:
            yield from _conv(f"down_blocks.{i}.downsamplers.0.conv", cout, cout)
    cm = cfg.block_out_channels[-1]
    yield from _resnet("mid_block.resnets.0", cm, cm, temb)
    yield from _transformer("mid_block.attentions.0", cm, cfg.transformer_layers_per_block[-1], cfg.cross_attention_dim, cfg.use_linear_projection)
    yield from _resnet("mid_block.resnets.1", cm, cm, temb)


def unet_shapes(cfg: UNetConfig) -> Shapes:
    yield from _encoder_shapes(cfg)
    temb = 4 * cfg.block_out_channels[0]
    rev = list(reversed(cfg.block_out_channels))
    rev_depth = list(reversed(cfg.transformer_layers_per_block))
    n = len(rev)
    cout = rev[0]
    for i, t in enumerate(cfg.up_block_types):
        prev, cout = cout, rev[i]
        cin = rev[min(i + 1, n - 1)]
        L = cfg.layers_per_block + 1
        for j in range(L):
            skip = cin if j == L - 1 else cout
            rin = prev if j == 0 else cout
            yield from _resnet(f"up_blocks.{i}.resnets.{j}", rin + skip, cout, temb)
            if t.startswith("CrossAttn"):
                yield from _transformer(f"up_blocks.{i}.attentions.{j}", cout, rev_depth[i], cfg.cross_attention_dim, cfg.use_linear_projection)
        if i != n - 1:
            yield from _conv(f"up_blocks.{i}.upsamplers.0.conv", cout, cout)
    yield from _norm("conv_norm_out", cfg.block_out_channels[0])
    yield from _conv("conv_out", cfg.block_out_channels[0], cfg.out_channels)


def controlnet_shapes(cfg: UNetConfig) -> Shapes:
    yield from _encoder_shapes(cfg)
    bo = cfg.conditioning_embedding_out_channels
    p = "controlnet_cond_embedding"
    yield from _conv(p + ".conv_in", 3, bo[0])
    for i in range(len(bo) - 1):
        yield from _conv(f"{p}.blocks.{2 * i}", bo[i], bo[i])
        yield from _conv(f"{p}.blocks.{2 * i + 1}", bo[i], bo[i + 1])
    yield from _conv(p + ".conv_out", bo[-1], cfg.block_out_channels[0])
    chans = [cfg.block_out_channels[0]]
    n = len(cfg.block_out_channels)
    for i, c in enumerate(cfg.block_out_channels):
        chans += [c] * cfg.layers_per_block
        if i != n - 1:
            chans.append(c)
    for i, c in enumerate(chans):
        yield from _conv(f"controlnet_down_blocks.{i}", c, c, 1)
    yield from _conv("controlnet_mid_block", cfg.block_out_channels[-1], cfg.block_out_channels[-1], 1)


def _vae_mid(p, c) -> Shapes:
    yield from _resnet(p + ".resnets.0", c, c, None)
    a = p + ".attentions.0"
    yield from _norm(a + ".group_norm", c)
    for nm in ("to_q", "to_k", "to_v", "to_out.0"):
        yield from _lin(f"{a}.{nm}", c, c)
    yield from _resnet(p + ".resnets.1", c, c, None)


def vae_shapes(cfg: VAEConfig) -> Shapes:
    ch = cfg.block_out_channels
    yield from _conv("encoder.conv_in", cfg.in_channels, ch[0])
    cout = ch[0]
    for i, c in enumerate(ch):
        cin, cout = cout, c
        for j in range(cfg.layers_per_block):
            yield from _resnet(f"encoder.down_blocks.{i}.resnets.{j}", cin if j == 0 else cout, cout, None)
        if i != len(ch) - 1:
            yield from _conv(f"encoder.down_blocks.{i}.downsamplers.0.conv", cout, cout)
    yield from _vae_mid("encoder.mid_block", ch[-1])
    yield from _norm("encoder.conv_norm_out", ch[-1])
    yield from _conv("encoder.conv_out", ch[-1], 2 * cfg.latent_channels)
    yield from _conv("quant_conv", 2 * cfg.latent_channels, 2 * cfg.latent_channels, 1)
    yield from _conv("post_quant_conv", cfg.latent_channels, cfg.latent_channels, 1)
    rev = list(reversed(ch))
    yield from _conv("decoder.conv_in", cfg.latent_channels, rev[0])
    yield from _vae_mid("decoder.mid_block", rev[0])
    cout = rev[0]
    for i, c in enumerate(rev):
        cin, cout = cout, c
        for j in range(cfg.layers_per_block + 1):
            yield from _resnet(f"decoder.up_blocks.{i}.resnets.{j}", cin if j == 0 else cout, cout, None)
        if i != len(rev) - 1:
            yield from _conv(f"decoder.up_blocks.{i}.upsamplers.0.conv", cout, cout)
    yield from _norm("decoder.conv_norm_out", rev[-1])
    yield from _conv("decoder.conv_out", rev[-1], cfg.out_channels)


def clip_text_shapes(cfg: CLIPTextConfig) -> Shapes:
    p = "text_model."
    yield p + "embeddings.token_embedding.weight", (cfg.vocab_size, cfg.hidden_size)
    yield p + "embeddings.position_embedding.weight", (cfg.max_position_embeddings, cfg.hidden_size)
    for i in range(cfg.num_hidden_layers):
        q = f"{p}encoder.layers.{i}."
        for nm in ("q_proj", "k_proj", "v_proj", "out_proj"):
            yield from _lin(q + "self_attn." + nm, cfg.hidden_size, cfg.hidden_size)
        yield from _norm(q + "layer_norm1", cfg.hidden_size)
        yield from _lin(q + "mlp.fc1", cfg.hidden_size, cfg.intermediate_size)
        yield from _lin(q + "mlp.fc2", cfg.intermediate_size, cfg.hidden_size)
        yield from _norm(q + "layer_norm2", cfg.hidden_size)
    yield from _norm(p + "final_layer_norm", cfg.hidden_size)
    if cfg.projection_dim > 0:
        yield "text_projection.weight", (cfg.projection_dim, cfg.hidden_size)


def qformer_shapes(cfg: Blip2Config) -> Shapes:
    """State-dict layout of diffusers' Blip2QFormerModel (qformer/ of Salesforce/blipdiffusion-controlnet)."""
    h, vh = cfg.hidden_size, cfg.vision_hidden_size
    yield "query_tokens", (1, cfg.num_query_tokens, h)
    yield "embeddings.word_embeddings.weight", (cfg.vocab_size, h)
    yield "embeddings.position_embeddings.weight", (cfg.max_position_embeddings, h)
    yield from _norm("embeddings.LayerNorm", h)
    v = "visual_encoder."
    n_pos = (cfg.image_size // cfg.patch_size) ** 2 + 1
    yield v + "embeddings.class_embedding", (1, 1, vh)
    yield v + "embeddings.position_embedding", (1, n_pos, vh)
    yield v + "embeddings.patch_embedding.weight", (vh, 3, cfg.patch_size, cfg.patch_size)
    yield from _norm(v + "pre_layernorm", vh)
    for i in range(cfg.vision_num_hidden_layers):
        q = f"{v}encoder.layers.{i}."
        yield from _lin(q + "self_attn.qkv", vh, 3 * vh)
        yield from _lin(q + "self_attn.projection", vh, vh)
        yield from _norm(q + "layer_norm1", vh)
        yield from _lin(q + "mlp.fc1", vh, cfg.vision_intermediate_size)
        yield from _lin(q + "mlp.fc2", cfg.vision_intermediate_size, vh)
        yield from _norm(q + "layer_norm2", vh)
    yield from _norm(v + "post_layernorm", vh)
    yield from _lin("proj_layer.dense1", h, 4 * h)
    yield from _lin("proj_layer.dense2", 4 * h, h)
    yield from _norm("proj_layer.LayerNorm", h)
    for i in range(cfg.num_hidden_layers):
        p = f"encoder.layer.{i}."
        blocks = [("attention", h)] + ([("crossattention", vh)] if i % cfg.cross_attention_frequency == 0 else [])
        for blk, kv in blocks:
            yield from _lin(p + blk + ".attention.query", h, h)
            yield from _lin(p + blk + ".attention.key", kv, h)
            yield from _lin(p + blk + ".attention.value", kv, h)
            yield from _lin(p + blk + ".output.dense", h, h)
            yield from _norm(p + blk + ".output.LayerNorm", h)
        for sfx in ("", "_query"):
            yield from _lin(p + "intermediate" + sfx + ".dense", h, cfg.intermediate_size)
            yield from _lin(p + "output" + sfx + ".dense", cfg.intermediate_size, h)
            yield from _norm(p + "output" + sfx + ".LayerNorm", h)


_RESIDUAL_OUT = ("conv2.weight", "to_out.0.weight", "ff.net.2.weight", "proj_out.weight", "out_proj.weight", "mlp.fc2.weight", "conv3.weight", "c_proj.weight")
_ZERO_CONV = ("controlnet_down_blocks", "controlnet_mid_block", "controlnet_cond_embedding.conv_out")


def random_state_dict(shapes: Shapes, seed: int, zero_conv_std: float = 0.02, dtype=torch.float32) -> Dict[str, torch.Tensor]:
    """Deterministic variance-preserving init: weights ~ N(0, 1/fan_in) (x 1/sqrt2 on residual-branch
    output layers), norm scales ~ N(1, .05), biases / norm shifts ~ N(0, .02), embeddings ~ N(0, .02);
    ControlNet zero-convs are NON-zero (std 0.02) so the residual-injection path is exercised."""
    g = torch.Generator().manual_seed(seed)
    sd = {}
    for name, shape in shapes:
        if len(shape) >= 2:
            if "embedding.weight" in name or "embeddings.weight" in name or name.endswith(("class_embedding", "position_embedding", "query_tokens")):
                std = 0.02
            else:
                fan_in = 1
                for s in shape[1:]:
                    fan_in *= s
                std = 1.0 / math.sqrt(fan_in)
                if name.endswith(_RESIDUAL_OUT):
                    std /= math.sqrt(2.0)
                if any(z in name for z in _ZERO_CONV):
                    std = zero_conv_std
            t = torch.randn(shape, generator=g) * std
        elif "norm" in name.lower() and name.endswith("weight"):
            t = 1.0 + 0.05 * torch.randn(shape, generator=g)
        else:
            t = 0.02 * torch.randn(shape, generator=g)
        sd[name] = t.to(dtype)
    return sd


def count_params(shapes: Shapes) -> int:
    tot = 0
    for _, s in shapes:
        n = 1
        for d in s:
            n *= d
        tot += n
    return tot


# ----------------------------------------------------------------------------------------------------
# filter networks: WSDAN_CAL (fgvc/models/cal.py) and openai CLIP RN50 (all_utils/utils.py:253)
# ----------------------------------------------------------------------------------------------------
def _bn(p, c) -> Shapes:
    yield p + ".weight", (c,)
    yield p + ".bias", (c,)
    yield p + ".running_mean", (c,)
    yield p + ".running_var", (c,)


RESNET_LAYERS = {"resnet50": (3, 4, 6, 3), "resnet101": (3, 4, 23, 3)}


def wsdan_shapes(num_classes: int, net: str = "resnet50", M: int = 32) -> Shapes:
    """State-dict keys of fgvc.models.cal.WSDAN_CAL(num_classes, net=net): `features` is the nn.Sequential of
    fgvc/models/resnet.py:168-178 (conv1, bn1, relu, maxpool, layer1..4 at indices 0,1,4,5,6,7; layer4 stride 1)."""
    yield "features.0.weight", (64, 3, 7, 7)
    yield from _bn("features.1", 64)
    inplanes = 64
    for li, (planes, blocks) in enumerate(zip((64, 128, 256, 512), RESNET_LAYERS[net])):
        stride = 1 if li in (0, 3) else 2
        for b in range(blocks):
            p = f"features.{4 + li}.{b}"
            yield p + ".conv1.weight", (planes, inplanes, 1, 1)
            yield from _bn(p + ".bn1", planes)
            yield p + ".conv2.weight", (planes, planes, 3, 3)
            yield from _bn(p + ".bn2", planes)
            yield p + ".conv3.weight", (planes * 4, planes, 1, 1)
            yield from _bn(p + ".bn3", planes * 4)
            if b == 0 and (stride != 1 or inplanes != planes * 4):
                yield p + ".downsample.0.weight", (planes * 4, inplanes, 1, 1)
                yield from _bn(p + ".downsample.1", planes * 4)
            inplanes = planes * 4
    yield "attentions.conv.weight", (M, 2048, 1, 1)
    yield from _bn("attentions.bn", M)
    yield "fc.weight", (num_classes, M * 2048)


def clip_rn50_shapes(embed_dim=1024, width=64, layers=(3, 4, 6, 3), t_width=512, t_layers=12, vocab=49408, ctx=77, res=224) -> Shapes:
    yield "positional_embedding", (ctx, t_width)
    yield "text_projection", (t_width, embed_dim)
    yield "logit_scale", ()
    v = "visual."
    yield v + "conv1.weight", (width // 2, 3, 3, 3)
    yield from _bn(v + "bn1", width // 2)
    yield v + "conv2.weight", (width // 2, width // 2, 3, 3)
    yield from _bn(v + "bn2", width // 2)
    yield v + "conv3.weight", (width, width // 2, 3, 3)
    yield from _bn(v + "bn3", width)
    inplanes = width
    for li, blocks in enumerate(layers):
        planes = width * (2 ** li)
        stride = 1 if li == 0 else 2
        for b in range(blocks):
            p = f"{v}layer{li + 1}.{b}"
            s = stride if b == 0 else 1
            yield p + ".conv1.weight", (planes, inplanes, 1, 1)
            yield from _bn(p + ".bn1", planes)
            yield p + ".conv2.weight", (planes, planes, 3, 3)
            yield from _bn(p + ".bn2", planes)
            yield p + ".conv3.weight", (planes * 4, planes, 1, 1)
            yield from _bn(p + ".bn3", planes * 4)
            if s > 1 or inplanes != planes * 4:
                yield p + ".downsample.0.weight", (planes * 4, inplanes, 1, 1)
                yield from _bn(p + ".downsample.1", planes * 4)
            inplanes = planes * 4
    ed = width * 32
    a = v + "attnpool."
    yield a + "positional_embedding", ((res // 32) ** 2 + 1, ed)
    for nm in ("k_proj", "q_proj", "v_proj"):
        yield from _lin(a + nm, ed, ed)
    yield from _lin(a + "c_proj", ed, embed_dim)
    for i in range(t_layers):
        q = f"transformer.resblocks.{i}."
        yield q + "attn.in_proj_weight", (3 * t_width, t_width)
        yield q + "attn.in_proj_bias", (3 * t_width,)
        yield from _lin(q + "attn.out_proj", t_width, t_width)
        yield from _norm(q + "ln_1", t_width)
        yield from _lin(q + "mlp.c_fc", t_width, 4 * t_width)
        yield from _lin(q + "mlp.c_proj", 4 * t_width, t_width)
        yield from _norm(q + "ln_2", t_width)
    yield "token_embedding.weight", (vocab, t_width)
    yield from _norm("ln_final", t_width)


def _clip_text_tower_shapes(t_width, t_layers, vocab, ctx, embed_dim) -> Shapes:
    yield "positional_embedding", (ctx, t_width)
    yield "text_projection", (t_width, embed_dim)
    yield "logit_scale", ()
    for i in range(t_layers):
        q = f"transformer.resblocks.{i}."
        yield q + "attn.in_proj_weight", (3 * t_width, t_width)
        yield q + "attn.in_proj_bias", (3 * t_width,)
        yield from _lin(q + "attn.out_proj", t_width, t_width)
        yield from _norm(q + "ln_1", t_width)
        yield from _lin(q + "mlp.c_fc", t_width, 4 * t_width)
        yield from _lin(q + "mlp.c_proj", 4 * t_width, t_width)
        yield from _norm(q + "ln_2", t_width)
    yield "token_embedding.weight", (vocab, t_width)
    yield from _norm("ln_final", t_width)


def clip_vit_shapes(embed_dim=768, v_width=1024, v_layers=24, patch=14, res=224, t_width=768, t_layers=12, vocab=49408, ctx=77) -> Shapes:
    """openai-clip state-dict layout of a VisionTransformer CLIP (defaults: ViT-L/14, BASELINE config 5): ``clip/model.py``
    VisionTransformer (conv1 patch embedding without bias, class_embedding, positional_embedding, ln_pre, transformer, ln_post, proj)."""
    v = "visual."
    yield v + "conv1.weight", (v_width, 3, patch, patch)
    yield v + "class_embedding", (v_width,)
    yield v + "positional_embedding", ((res // patch) ** 2 + 1, v_width)
    yield from _norm(v + "ln_pre", v_width)
    for i in range(v_layers):
        q = f"{v}transformer.resblocks.{i}."
        yield q + "attn.in_proj_weight", (3 * v_width, v_width)
        yield q + "attn.in_proj_bias", (3 * v_width,)
        yield from _lin(q + "attn.out_proj", v_width, v_width)
        yield from _norm(q + "ln_1", v_width)
        yield from _lin(q + "mlp.c_fc", v_width, 4 * v_width)
        yield from _lin(q + "mlp.c_proj", 4 * v_width, v_width)
        yield from _norm(q + "ln_2", v_width)
    yield from _norm(v + "ln_post", v_width)
    yield v + "proj", (v_width, embed_dim)
    yield from _clip_text_tower_shapes(t_width, t_layers, vocab, ctx, embed_dim)


def clip_vit_tiny_kwargs() -> dict:
    return dict(embed_dim=64, v_width=128, v_layers=2, patch=14, res=56, t_width=64, t_layers=2, vocab=1000, ctx=77)


def safety_checker_shapes(width=1024, layers=24, patch=14, res=224, proj=768, n_concepts=17, n_special=3) -> Shapes:
    """State-dict layout of diffusers' StableDiffusionSafetyChecker (pipelines/stable_diffusion/safety_checker.py; loaded by default with
    runwayml/stable-diffusion-v1-5): transformers CLIPVisionModel (ViT-L/14) + visual_projection + concept / special-care embeddings and
    their thresholds (the ``*_weights`` buffers)."""
    v = "vision_model.vision_model."
    yield v + "embeddings.class_embedding", (width,)
    yield v + "embeddings.patch_embedding.weight", (width, 3, patch, patch)
    yield v + "embeddings.position_embedding.weight", ((res // patch) ** 2 + 1, width)
    yield from _norm(v + "pre_layrnorm", width)  # (sic) transformers' spelling
    for i in range(layers):
        q = f"{v}encoder.layers.{i}."
        for nm in ("q_proj", "k_proj", "v_proj", "out_proj"):
            yield from _lin(q + "self_attn." + nm, width, width)
        yield from _norm(q + "layer_norm1", width)
        yield from _lin(q + "mlp.fc1", width, 4 * width)
        yield from _lin(q + "mlp.fc2", 4 * width, width)
        yield from _norm(q + "layer_norm2", width)
    yield from _norm(v + "post_layernorm", width)
    yield "visual_projection.weight", (proj, width)
    yield "concept_embeds", (n_concepts, proj)
    yield "special_care_embeds", (n_special, proj)
    yield "concept_embeds_weights", (n_concepts,)
    yield "special_care_embeds_weights", (n_special,)


def safety_checker_tiny_kwargs() -> dict:
    return dict(width=128, layers=2, patch=14, res=56, proj=64)


def random_safety_checker_state_dict(shapes: Shapes, seed: int) -> Dict[str, torch.Tensor]:
    """Seeded init for the safety checker: CLIP-style weights; thresholds drawn around the typical cosine of random embeddings so that
    both outcomes (flagged / clean) and the special-care adjustment occur on synthetic images."""
    sd = random_state_dict(shapes, seed)
    g = torch.Generator().manual_seed(seed + 1)
    for k in ("concept_embeds", "special_care_embeds"):
        sd[k] = torch.randn(sd[k].shape, generator=g)
    sigma = sd["concept_embeds"].shape[1] ** -0.5  # std of the cosine between random directions
    sd["concept_embeds_weights"] = sigma * (2.0 + 0.3 * torch.randn(sd["concept_embeds_weights"].shape, generator=g))
    sd["special_care_embeds_weights"] = sigma * (1.3 + 0.3 * torch.randn(sd["special_care_embeds_weights"].shape, generator=g))
    return sd


def random_filter_state_dict(shapes: Shapes, seed: int) -> Dict[str, torch.Tensor]:
    """He-normal conv/linear weights (ReLU nets), BatchNorm affine ~ N(1,.1)/N(0,.1) with non-trivial running stats,
    the last BN scale of each residual branch damped (x0.5) so depth does not blow the activations up."""
    g = torch.Generator().manual_seed(seed)
    sd = {}
    for name, shape in shapes:
        if name == "logit_scale":
            t = torch.tensor(math.log(1 / 0.07))
        elif name.endswith("running_var"):
            t = 0.5 + torch.rand(shape, generator=g)
        elif name.endswith("running_mean"):
            t = 0.1 * torch.randn(shape, generator=g)
        elif "positional_embedding" in name or name == "token_embedding.weight":
            t = 0.02 * torch.randn(shape, generator=g)
        elif name in ("text_projection", "visual.proj"):
            t = torch.randn(shape, generator=g) * shape[0] ** -0.5
        elif name == "visual.class_embedding":
            t = 0.02 * torch.randn(shape, generator=g)
        elif len(shape) >= 2:
            fan_in = 1
            for s in shape[1:]:
                fan_in *= s
            relu_net = name.startswith(("features.", "visual.", "attentions.")) and "attnpool" not in name and "transformer." not in name
            std = math.sqrt((2.0 if relu_net else 1.0) / fan_in)
            if name.endswith(_RESIDUAL_OUT) and not relu_net:
                std /= math.sqrt(2.0)
            t = torch.randn(shape, generator=g) * std
        elif (".bn" in name or "downsample.1" in name or "features.1" in name) and name.endswith("weight"):
            t = 1.0 + 0.1 * torch.randn(shape, generator=g)
            if ".bn3." in name:
                t = t * 0.5
        elif ("ln_" in name or "ln_final" in name) and name.endswith("weight"):
            t = 1.0 + 0.05 * torch.randn(shape, generator=g)
        else:
            t = 0.05 * torch.randn(shape, generator=g)
        sd[name] = t.float()
    return sd


def lpips_alex_shapes() -> Shapes:
    """lpips 0.1.4 ``LPIPS(net='alex')`` state dict, the entries the distance reads: torchvision AlexNet features sliced at the five
    ReLUs (``net.slice{k}.{idx}``) and the non-negative 1x1 heads (``lin{k}.model.1.weight``)."""
    for k, (idx, shape) in enumerate(zip((0, 3, 6, 8, 10), ((64, 3, 11, 11), (192, 64, 5, 5), (384, 192, 3, 3), (256, 384, 3, 3), (256, 256, 3, 3)))):
        yield f"net.slice{k + 1}.{idx}.weight", shape
        yield f"net.slice{k + 1}.{idx}.bias", (shape[0],)
    for k, c in enumerate((64, 192, 384, 256, 256)):
        yield f"lin{k}.model.1.weight", (1, c, 1, 1)


def random_lpips_state_dict(seed: int) -> Dict[str, torch.Tensor]:
    """He-normal convolutions (ReLU net), small biases, uniform non-negative ``lin`` weights (the learned heads are clamped >= 0)."""
    g = torch.Generator().manual_seed(seed)
    sd = {}
    for name, shape in lpips_alex_shapes():
        if name.startswith("lin"):
            sd[name] = torch.rand(shape, generator=g) / shape[1]
        elif name.endswith("bias"):
            sd[name] = 0.05 * torch.randn(shape, generator=g)
        else:
            sd[name] = torch.randn(shape, generator=g) * (2.0 / (shape[1] * shape[2] * shape[3])) ** 0.5
    return sd


HED_BLOCKS = ((3, 64, 2), (64, 128, 2), (128, 256, 3), (256, 512, 3), (512, 512, 3))  # (in, out, convolutions) of block1..block5


def hed_shapes() -> Shapes:
    """State-dict layout of the ControlNet annotators' HED network (``ControlNetHED.pth``, loaded by controlnet_aux's
    ``HEDdetector.from_pretrained``; run_aug/run_aug.py:311-312)."""
    yield "norm", (1, 3, 1, 1)
    for k, (cin, cout, layers) in enumerate(HED_BLOCKS, 1):
        for i in range(layers):
            yield from _conv(f"block{k}.convs.{i}", cin if i == 0 else cout, cout)
        yield from _conv(f"block{k}.projection", cout, 1, 1)


def random_hed_state_dict(seed: int) -> Dict[str, torch.Tensor]:
    """He-normal trunk (ReLU net, activations keep the 0..255 input scale), projections scaled so that the five side outputs are
    O(1) logits (a saturated sigmoid would hide every numerical difference), ``norm`` near the usual per-channel image mean."""
    g = torch.Generator().manual_seed(seed)
    sd = {}
    for name, shape in hed_shapes():
        if name == "norm":
            sd[name] = torch.tensor([123.7, 116.3, 103.5]).view(1, 3, 1, 1) + torch.randn(shape, generator=g)
        elif ".projection." in name:
            sd[name] = torch.randn(shape, generator=g) * (0.03 / shape[1] ** 0.5) if name.endswith("weight") else 0.2 * torch.randn(shape, generator=g)
        elif name.endswith("bias"):
            sd[name] = 2.0 * torch.randn(shape, generator=g)
        else:
            sd[name] = torch.randn(shape, generator=g) * (2.0 / (shape[1] * shape[2] * shape[3])) ** 0.5
    return sd
