"""CPU fp32 restatement of the BLIP-Diffusion front end of diffusers 0.32.2 (un-vendored dependency, environment.yml:17):
``pipelines/blip_diffusion/modeling_blip2.py`` (Blip2QFormerModel = BERT-style Q-Former with 16 learned query tokens + its own
ViT vision encoder + ProjLayer), ``modeling_ctx_clip.py`` (ContextCLIPTextModel: CLIP text tower with the 16 subject embeddings
spliced into the token sequence at ``ctx_begin_pos``), ``blip_image_processing.py`` (BlipImageProcessor) and the control flow of
``pipelines/controlnet/pipeline_controlnet_blip_diffusion.py`` (BlipDiffusionControlNetPipeline.__call__) as the reference
invokes it (run_aug/run_aug.py:243-250,268-271: reference_image, source/target_subject_category, condtioning_image, neg_prompt,
height/width from the control image; scheduler stays the checkpoint's PNDM, run_aug.py:217).

State-dict keys follow the checkpoint layout of Salesforce/blipdiffusion-controlnet (qformer/, text_encoder/).

TEST INFRASTRUCTURE.  **Parity unpinned** at the reference boundary (the reference holds no golden vectors); the modules are
cross-checked in tests/test_blip_cpu.py against the installed ``transformers`` BLIP-2 / CLIP building blocks
(Blip2VisionModel, Blip2QFormerEncoder, CLIPTextModel) on identical random weights.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import List, Optional, Sequence

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F


@dataclass
class Blip2Config:
    # qformer_config
    vocab_size: int = 30523
    hidden_size: int = 768
    num_hidden_layers: int = 12
    num_attention_heads: int = 12
    intermediate_size: int = 3072
    max_position_embeddings: int = 512
    layer_norm_eps: float = 1e-12
    cross_attention_frequency: int = 2
    num_query_tokens: int = 16
    # vision_config
    vision_hidden_size: int = 1024
    vision_intermediate_size: int = 4096
    vision_num_hidden_layers: int = 23
    vision_num_attention_heads: int = 16
    image_size: int = 224
    patch_size: int = 14
    vision_layer_norm_eps: float = 1e-5

    @staticmethod
    def blipdiffusion() -> "Blip2Config":  # Salesforce/blipdiffusion(-controlnet) qformer/config.json
        return Blip2Config()

    @staticmethod
    def tiny() -> "Blip2Config":
        return Blip2Config(vocab_size=500, hidden_size=64, num_hidden_layers=4, num_attention_heads=2, intermediate_size=128, max_position_embeddings=32,
                           num_query_tokens=16, vision_hidden_size=128, vision_intermediate_size=256, vision_num_hidden_layers=3,
                           vision_num_attention_heads=2, image_size=56, patch_size=14)


def quick_gelu(x):
    return x * torch.sigmoid(1.702 * x)


# ---- vision encoder (modeling_blip2.Blip2VisionModel; layers = transformers Blip2EncoderLayer) ----------------------------
class _VisionAttention(nn.Module):
    def __init__(self, dim, heads):
        super().__init__()
        self.heads = heads
        self.qkv = nn.Linear(dim, 3 * dim)  # checkpoint stores the fused bias (q_bias | 0 | v_bias)
        self.projection = nn.Linear(dim, dim)

    def forward(self, x):
        b, t, c = x.shape
        qkv = self.qkv(x).reshape(b, t, 3, self.heads, c // self.heads).permute(2, 0, 3, 1, 4)
        q, k, v = qkv[0], qkv[1], qkv[2]
        a = torch.softmax(q @ k.transpose(-1, -2) * (c // self.heads) ** -0.5, dim=-1)
        return self.projection((a @ v).permute(0, 2, 1, 3).reshape(b, t, c))


class _VisionMLP(nn.Module):
    def __init__(self, dim, inner):
        super().__init__()
        self.fc1, self.fc2 = nn.Linear(dim, inner), nn.Linear(inner, dim)

    def forward(self, x):
        return self.fc2(quick_gelu(self.fc1(x)))


class _VisionLayer(nn.Module):
    def __init__(self, dim, heads, inner, eps):
        super().__init__()
        self.self_attn = _VisionAttention(dim, heads)
        self.layer_norm1 = nn.LayerNorm(dim, eps=eps)
        self.mlp = _VisionMLP(dim, inner)
        self.layer_norm2 = nn.LayerNorm(dim, eps=eps)

    def forward(self, x):
        x = x + self.self_attn(self.layer_norm1(x))
        return x + self.mlp(self.layer_norm2(x))


class _VisionEmbeddings(nn.Module):
    def __init__(self, cfg: Blip2Config):
        super().__init__()
        d = cfg.vision_hidden_size
        self.class_embedding = nn.Parameter(torch.randn(1, 1, d))
        self.patch_embedding = nn.Conv2d(3, d, cfg.patch_size, cfg.patch_size, bias=False)
        self.position_embedding = nn.Parameter(torch.randn(1, (cfg.image_size // cfg.patch_size) ** 2 + 1, d))

    def forward(self, pixel_values):
        b = pixel_values.shape[0]
        p = self.patch_embedding(pixel_values.to(self.patch_embedding.weight.dtype)).flatten(2).transpose(1, 2)
        e = torch.cat([self.class_embedding.expand(b, 1, -1).to(p.dtype), p], dim=1)
        return e + self.position_embedding[:, : e.size(1), :].to(p.dtype)


class _VisionEncoder(nn.Module):
    def __init__(self, cfg):
        super().__init__()
        self.layers = nn.ModuleList([_VisionLayer(cfg.vision_hidden_size, cfg.vision_num_attention_heads, cfg.vision_intermediate_size, cfg.vision_layer_norm_eps)
                                     for _ in range(cfg.vision_num_hidden_layers)])


class Blip2VisionModel(nn.Module):
    """embeddings -> pre_layernorm -> encoder -> post_layernorm (last_hidden_state is what the Q-Former cross-attends to)."""

    def __init__(self, cfg: Blip2Config):
        super().__init__()
        self.embeddings = _VisionEmbeddings(cfg)
        self.pre_layernorm = nn.LayerNorm(cfg.vision_hidden_size, eps=cfg.vision_layer_norm_eps)
        self.encoder = _VisionEncoder(cfg)
        self.post_layernorm = nn.LayerNorm(cfg.vision_hidden_size, eps=cfg.vision_layer_norm_eps)

    def forward(self, pixel_values):
        h = self.pre_layernorm(self.embeddings(pixel_values))
        for l in self.encoder.layers:
            h = l(h)
        return self.post_layernorm(h)


# ---- Q-Former (BERT-style, post-LayerNorm) ---------------------------------------------------------------------------------
class _BertSelfAttention(nn.Module):
    def __init__(self, dim, heads, kv_dim):
        super().__init__()
        self.heads = heads
        self.query, self.key, self.value = nn.Linear(dim, dim), nn.Linear(kv_dim, dim), nn.Linear(kv_dim, dim)

    def forward(self, x, ctx=None):
        ctx = x if ctx is None else ctx
        b, t, c = x.shape
        d = c // self.heads

        def split(y):
            return y.view(b, -1, self.heads, d).permute(0, 2, 1, 3)

        q, k, v = split(self.query(x)), split(self.key(ctx)), split(self.value(ctx))
        a = torch.softmax(q @ k.transpose(-1, -2) / math.sqrt(d), dim=-1)  # all-ones masks (no padding) add 0
        return (a @ v).permute(0, 2, 1, 3).reshape(b, t, c)


class _BertSelfOutput(nn.Module):
    def __init__(self, dim, eps):
        super().__init__()
        self.dense = nn.Linear(dim, dim)
        self.LayerNorm = nn.LayerNorm(dim, eps=eps)

    def forward(self, h, inp):
        return self.LayerNorm(self.dense(h) + inp)


class _BertAttention(nn.Module):
    def __init__(self, dim, heads, kv_dim, eps):
        super().__init__()
        self.attention = _BertSelfAttention(dim, heads, kv_dim)
        self.output = _BertSelfOutput(dim, eps)

    def forward(self, x, ctx=None):
        return self.output(self.attention(x, ctx), x)


class _Intermediate(nn.Module):
    def __init__(self, dim, inner):
        super().__init__()
        self.dense = nn.Linear(dim, inner)

    def forward(self, x):
        return F.gelu(self.dense(x))


class _Output(nn.Module):
    def __init__(self, inner, dim, eps):
        super().__init__()
        self.dense = nn.Linear(inner, dim)
        self.LayerNorm = nn.LayerNorm(dim, eps=eps)

    def forward(self, h, inp):
        return self.LayerNorm(self.dense(h) + inp)


class _QFormerLayer(nn.Module):
    def __init__(self, cfg: Blip2Config, idx: int):
        super().__init__()
        d, e = cfg.hidden_size, cfg.layer_norm_eps
        self.attention = _BertAttention(d, cfg.num_attention_heads, d, e)
        self.has_cross_attention = idx % cfg.cross_attention_frequency == 0
        if self.has_cross_attention:
            self.crossattention = _BertAttention(d, cfg.num_attention_heads, cfg.vision_hidden_size, e)
        self.intermediate, self.output = _Intermediate(d, cfg.intermediate_size), _Output(cfg.intermediate_size, d, e)
        self.intermediate_query, self.output_query = _Intermediate(d, cfg.intermediate_size), _Output(cfg.intermediate_size, d, e)

    def forward(self, h, image_embeds, query_length):
        a = self.attention(h)
        q = a[:, :query_length]
        if self.has_cross_attention:
            q = self.crossattention(q, image_embeds)
        out = self.output_query(self.intermediate_query(q), q)
        if a.shape[1] > query_length:
            t = a[:, query_length:]
            out = torch.cat([out, self.output(self.intermediate(t), t)], dim=1)
        return out


class _QFormerEncoder(nn.Module):
    def __init__(self, cfg):
        super().__init__()
        self.layer = nn.ModuleList([_QFormerLayer(cfg, i) for i in range(cfg.num_hidden_layers)])


class _TextEmbeddings(nn.Module):
    def __init__(self, cfg: Blip2Config):
        super().__init__()
        self.word_embeddings = nn.Embedding(cfg.vocab_size, cfg.hidden_size)
        self.position_embeddings = nn.Embedding(cfg.max_position_embeddings, cfg.hidden_size)
        self.LayerNorm = nn.LayerNorm(cfg.hidden_size, eps=cfg.layer_norm_eps)

    def forward(self, input_ids, query_embeds):
        e = self.word_embeddings(input_ids) + self.position_embeddings(torch.arange(input_ids.shape[1], device=input_ids.device))[None]
        e = torch.cat([query_embeds.repeat(e.shape[0], 1, 1), e], dim=1).to(query_embeds.dtype)
        return self.LayerNorm(e)


class ProjLayer(nn.Module):
    def __init__(self, dim, hidden, eps=1e-12):
        super().__init__()
        self.dense1, self.dense2 = nn.Linear(dim, hidden), nn.Linear(hidden, dim)
        self.LayerNorm = nn.LayerNorm(dim, eps=eps)

    def forward(self, x):
        return self.dense2(quick_gelu(self.dense1(self.LayerNorm(x)))) + x


class Blip2QFormerModel(nn.Module):
    """forward(image_input [b,3,S,S] normalised, text ids [b,L] (BertTokenizer of the source subject, no padding))
    -> proj_layer(sequence_output[:, :num_query_tokens]) [b,16,hidden]."""

    def __init__(self, cfg: Blip2Config):
        super().__init__()
        self.cfg = cfg
        self.embeddings = _TextEmbeddings(cfg)
        self.visual_encoder = Blip2VisionModel(cfg)
        self.query_tokens = nn.Parameter(torch.zeros(1, cfg.num_query_tokens, cfg.hidden_size))
        self.proj_layer = ProjLayer(cfg.hidden_size, cfg.hidden_size * 4, eps=1e-12)
        self.encoder = _QFormerEncoder(cfg)

    def forward(self, image_input, input_ids):
        h = self.embeddings(input_ids, self.query_tokens)
        img = self.visual_encoder(image_input)
        nq = self.query_tokens.shape[1]
        for l in self.encoder.layer:
            h = l(h, img, nq)
        return self.proj_layer(h[:, :nq])


# ---- ContextCLIPTextModel -----------------------------------------------------------------------------------------------
class ContextCLIPTextModel(nn.Module):
    """Wraps a ``transformers`` CLIPTextModel (same parameters / keys as diffusers' ContextCLIPTextModel) and restates
    ContextCLIPTextEmbeddings: ctx embeddings are inserted after ``ctx_begin_pos`` token embeddings, positions 0..L+ctx-1, then the
    standard causal CLIP encoder + final LayerNorm."""

    def __init__(self, clip_text_model):
        super().__init__()
        self.m = clip_text_model

    def forward(self, input_ids, ctx_embeddings=None, ctx_begin_pos: Optional[Sequence[int]] = None):
        tm = self.m.text_model
        e = tm.embeddings.token_embedding(input_ids)
        if ctx_embeddings is not None:
            rows = []
            for i in range(e.shape[0]):
                c = ctx_begin_pos[i]
                rows.append(torch.cat([e[i, :c], ctx_embeddings[i].to(e.dtype), e[i, c:]], dim=0))
            e = torch.stack(rows, 0)
        t = e.shape[1]
        h = e + tm.embeddings.position_embedding(torch.arange(t, device=e.device))[None]
        mask = torch.full((t, t), float("-inf"), device=h.device, dtype=h.dtype).triu(1)[None, None]
        for layer in tm.encoder.layers:
            r = h
            x = layer.layer_norm1(h)
            a = layer.self_attn
            b, _, c = x.shape
            nh = a.num_heads if hasattr(a, "num_heads") else a.config.num_attention_heads
            d = c // nh

            def split(y):
                return y.view(b, t, nh, d).transpose(1, 2)

            q, k, v = split(a.q_proj(x)), split(a.k_proj(x)), split(a.v_proj(x))
            w = torch.softmax(q @ k.transpose(-1, -2) * d ** -0.5 + mask, dim=-1)
            h = r + a.out_proj((w @ v).transpose(1, 2).reshape(b, t, c))
            h = h + layer.mlp(layer.layer_norm2(h))
        return tm.final_layer_norm(h)


# ---- image processor -------------------------------------------------------------------------------------------------------
OPENAI_CLIP_MEAN = (0.48145466, 0.4578275, 0.40821073)
OPENAI_CLIP_STD = (0.26862954, 0.26130258, 0.27577711)


def blip_preprocess_reference(image_u8: np.ndarray, size: int = 224, mean=OPENAI_CLIP_MEAN, std=OPENAI_CLIP_STD) -> torch.Tensor:
    """BlipImageProcessor.preprocess(reference_image, image_mean, image_std): PIL bicubic resize to size x size (uint8 result),
    x 1/255, normalise; the final center crop to the same size is the identity.  [n,H,W,3] u8 -> [n,3,size,size] fp32."""
    from PIL import Image

    out = []
    for a in image_u8:
        r = np.asarray(Image.fromarray(a).convert("RGB").resize((size, size), resample=Image.BICUBIC))
        x = r.astype(np.float32) * np.float32(1 / 255.0)
        out.append((x - np.asarray(mean, np.float32)) / np.asarray(std, np.float32))
    return torch.from_numpy(np.stack(out)).permute(0, 3, 1, 2).contiguous()


def build_prompt(prompts: Sequence[str], tgt_subjects: Sequence[str], prompt_strength: float = 1.0, prompt_reps: int = 20) -> List[str]:
    """BlipDiffusionControlNetPipeline._build_prompt: "a {tgt} {prompt}" repeated int(strength * reps) times, comma-joined."""
    rv = []
    for prompt, tgt in zip(prompts, tgt_subjects):
        p = f"a {tgt} {prompt.strip()}"
        rv.append(", ".join([p] * int(prompt_strength * prompt_reps)))
    return rv
