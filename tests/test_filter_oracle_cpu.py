"""CPU: the filter-side oracles (WSDAN_CAL restatement, CLIP RN50 restatement driven through the reference's own
CLIP_selector arithmetic) against the committed golden outputs of the reference code
(tests/golden/make_filter_golden.py), and live against /root/reference when present."""
import json
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import clib, clip_rn50, ref_import, wsdan
from saspa_aug_b200 import checkpoints as ck
from saspa_aug_b200.synthetic import synthetic_source, synthetic_token_ids

G = os.path.join(os.path.dirname(__file__), "golden")
META = json.load(open(os.path.join(G, "filter_golden.json")))
GOLD = np.load(os.path.join(G, "filter_golden.npz"))


def preprocess_baseline(img_u8: np.ndarray) -> torch.Tensor:
    """all_utils/dataset_utils.py:78-85 via the pinned PIL-resize oracle: Resize(256,256) -> CenterCrop(224) -> /255 -> Normalize."""
    r = clib.pil_resize(img_u8, 256, 256, "bilinear")[16:240, 16:240]
    x = torch.from_numpy(r.astype(np.float32) / 255.0).permute(2, 0, 1)
    m, s = torch.tensor([0.485, 0.456, 0.406])[:, None, None], torch.tensor([0.229, 0.224, 0.225])[:, None, None]
    return (x - m) / s


def preprocess_clip(img_u8: np.ndarray) -> torch.Tensor:
    h, w = img_u8.shape[:2]
    oh, ow = (224, int(224 * w / h)) if h <= w else (int(224 * h / w), 224)
    r = clib.pil_resize(img_u8, oh, ow, "bicubic")
    cy, cx = int(round((oh - 224) / 2.0)), int(round((ow - 224) / 2.0))
    x = torch.from_numpy(r[cy:cy + 224, cx:cx + 224].astype(np.float32) / 255.0).permute(2, 0, 1)
    m = torch.tensor([0.48145466, 0.4578275, 0.40821073])[:, None, None]
    s = torch.tensor([0.26862954, 0.26130258, 0.27577711])[:, None, None]
    return (x - m) / s


def test_wsdan_oracle_matches_reference_golden_logits():
    sd = ck.random_filter_state_dict(ck.wsdan_shapes(META["classes"], "resnet50"), META["wsdan_seed"])
    m = wsdan.WSDANOracle(META["classes"], "resnet50").eval()
    missing, unexpected = m.load_state_dict(sd, strict=False)
    assert not unexpected and all("num_batches_tracked" in k for k in missing)
    x = torch.stack([preprocess_baseline(synthetic_source(s)) for s in META["image_seeds"][:3]])
    assert np.allclose(x[0, :, ::32, ::32].numpy(), GOLD["wsdan_input_sample"], atol=1e-6)  # the transform itself
    with torch.no_grad():
        logits = m(x)
    assert np.allclose(logits.numpy(), GOLD["wsdan_logits"][:3], atol=2e-4), np.abs(logits.numpy() - GOLD["wsdan_logits"][:3]).max()


def test_clip_oracle_matches_reference_selector_golden():
    sd = ck.random_filter_state_dict(ck.clip_rn50_shapes(), META["clip_seed"])
    c = clip_rn50.CLIP().eval()
    c.load_state_dict(sd)
    ids = torch.cat([synthetic_token_ids(s) for s in META["prompt_id_seeds"]])
    with torch.no_grad():
        x = torch.stack([preprocess_clip(synthetic_source(s)) for s in META["image_seeds"][:2]])
        fi, ft = c.encode_image(x), c.encode_text(ids)
        logits = c.logit_scale.exp() * F.normalize(fi, dim=-1) @ F.normalize(ft, dim=-1).t()
    assert np.allclose(logits.numpy(), GOLD["clip_logits"][:2], atol=2e-4)
    assert ((logits.argmax(-1) == 0).numpy().astype(np.uint8) == GOLD["semantic_keep"][:2]).all()


@pytest.mark.skipif(not ref_import.available(), reason="/root/reference not present (GPU box)")
def test_wsdan_oracle_against_reference_live_resnet101():
    cal = ref_import.import_reference_cal()
    sd = ck.random_filter_state_dict(ck.wsdan_shapes(17, "resnet101"), 5)
    ref = cal.WSDAN_CAL(17, net="resnet101", print_func=lambda *a: None).eval()
    ref.load_state_dict(sd)
    mine = wsdan.WSDANOracle(17, "resnet101").eval()
    mine.load_state_dict(sd, strict=False)
    x = torch.randn(2, 3, 224, 224, generator=torch.Generator().manual_seed(0))
    with torch.no_grad():
        assert torch.allclose(ref(x)[0], mine(x), atol=1e-4)


def test_topk_rule():
    # exact ties are implementation-defined in torch.topk (and measure-zero for fp32 logits); use distinct values
    logits = torch.tensor([[0.1, 0.9, 0.5, 0.4], [1.0, 0.7, 0.8, 0.9]])
    assert wsdan.in_topk(logits, [2, 1], 2).tolist() == [1, 0]
    assert wsdan.in_topk(logits, [3, 3], 2).tolist() == [0, 1]


def _openai_to_hf_clip(sd, v_layers, t_layers):
    """openai-clip state-dict keys -> transformers.CLIPModel keys (the published conversion: in_proj split into q/k/v, proj transposed)."""
    out = {"logit_scale": sd["logit_scale"], "text_projection.weight": sd["text_projection"].t(), "visual_projection.weight": sd["visual.proj"].t(),
           "text_model.embeddings.token_embedding.weight": sd["token_embedding.weight"], "text_model.embeddings.position_embedding.weight": sd["positional_embedding"],
           "text_model.final_layer_norm.weight": sd["ln_final.weight"], "text_model.final_layer_norm.bias": sd["ln_final.bias"],
           "vision_model.embeddings.class_embedding": sd["visual.class_embedding"], "vision_model.embeddings.patch_embedding.weight": sd["visual.conv1.weight"],
           "vision_model.embeddings.position_embedding.weight": sd["visual.positional_embedding"],
           "vision_model.pre_layrnorm.weight": sd["visual.ln_pre.weight"], "vision_model.pre_layrnorm.bias": sd["visual.ln_pre.bias"],
           "vision_model.post_layernorm.weight": sd["visual.ln_post.weight"], "vision_model.post_layernorm.bias": sd["visual.ln_post.bias"]}
    for src, dst, n in (("transformer.resblocks.", "text_model.encoder.layers.", t_layers), ("visual.transformer.resblocks.", "vision_model.encoder.layers.", v_layers)):
        for i in range(n):
            a, b = f"{src}{i}.", f"{dst}{i}."
            w, bias = sd[a + "attn.in_proj_weight"], sd[a + "attn.in_proj_bias"]
            c = w.shape[1]
            for j, nm in enumerate(("q_proj", "k_proj", "v_proj")):
                out[b + f"self_attn.{nm}.weight"], out[b + f"self_attn.{nm}.bias"] = w[j * c:(j + 1) * c], bias[j * c:(j + 1) * c]
            for s_, d_ in (("attn.out_proj", "self_attn.out_proj"), ("ln_1", "layer_norm1"), ("ln_2", "layer_norm2"), ("mlp.c_fc", "mlp.fc1"), ("mlp.c_proj", "mlp.fc2")):
                out[b + d_ + ".weight"], out[b + d_ + ".bias"] = sd[a + s_ + ".weight"], sd[a + s_ + ".bias"]
    return out


def test_clip_vit_oracle_matches_transformers_clipmodel():
    """The ViT CLIP restatement (openai layout; BASELINE config 5 uses ViT-L/14) against transformers.CLIPModel on the same weights."""
    from transformers import CLIPConfig, CLIPModel

    kw = ck.clip_vit_tiny_kwargs()
    sd = ck.random_filter_state_dict(ck.clip_vit_shapes(**kw), 9)
    c = clip_rn50.clip_vit(**kw).eval()
    c.load_state_dict(sd, strict=True)
    cfg = CLIPConfig(projection_dim=kw["embed_dim"],
                     text_config=dict(vocab_size=kw["vocab"], hidden_size=kw["t_width"], intermediate_size=4 * kw["t_width"], num_hidden_layers=kw["t_layers"],
                                      num_attention_heads=max(1, kw["t_width"] // 64), max_position_embeddings=kw["ctx"], hidden_act="quick_gelu",
                                      bos_token_id=kw["vocab"] - 2, eos_token_id=kw["vocab"] - 1),
                     vision_config=dict(hidden_size=kw["v_width"], intermediate_size=4 * kw["v_width"], num_hidden_layers=kw["v_layers"],
                                        num_attention_heads=kw["v_width"] // 64, image_size=kw["res"], patch_size=kw["patch"], hidden_act="quick_gelu"))
    hf = CLIPModel(cfg).eval()
    missing, unexpected = hf.load_state_dict(_openai_to_hf_clip(sd, kw["v_layers"], kw["t_layers"]), strict=False)
    assert not unexpected and all("position_ids" in k for k in missing), (missing, unexpected)
    g = torch.Generator().manual_seed(0)
    x = torch.randn((2, 3, kw["res"], kw["res"]), generator=g)
    ids = torch.cat([synthetic_token_ids(s, vocab=kw["vocab"]) for s in (1, 2, 3)])
    with torch.no_grad():
        fi, ft = c.encode_image(x), c.encode_text(ids)
        hi = hf.visual_projection(hf.vision_model(pixel_values=x).pooler_output)
        ht = hf.text_projection(hf.text_model(input_ids=ids).pooler_output)
    assert torch.allclose(fi, hi, atol=2e-5), (fi - hi).abs().max()
    assert torch.allclose(ft, ht, atol=2e-5), (ft - ht).abs().max()


def test_clip_vit_l14_shapes():
    n = ck.count_params(ck.clip_vit_shapes())
    assert 420e6 < n < 435e6, n  # openai ViT-L/14: 427.6 M parameters


def test_safety_checker_oracle_loads_product_layout_and_flags_both_ways():
    """The product's checkpoint layout == diffusers' StableDiffusionSafetyChecker keys (strict load into the oracle, whose tower is
    transformers' CLIPVisionModel); with the seeded thresholds both outcomes occur, and flagged images come back black."""
    from oracle import safety_checker as osc

    kw = ck.safety_checker_tiny_kwargs()
    sd = ck.random_safety_checker_state_dict(ck.safety_checker_shapes(**kw), 3)
    m = osc.SafetyCheckerOracle(**kw).eval()
    missing, unexpected = m.load_state_dict(sd, strict=False)
    assert not unexpected and all("position_ids" in k for k in missing), (missing, unexpected)
    imgs = np.stack([synthetic_source(100 + i, 96, 128, kind=("blobs", "noise", "smooth")[i % 3]) for i in range(12)])
    x = osc.clip_image_processor(imgs, kw["res"])
    assert x.shape == (12, 3, 56, 56)
    out, flags, res = m(x, imgs)
    assert 0 < sum(flags) < 12, flags
    for i, f in enumerate(flags):
        assert (out[i] == 0).all() if f else np.array_equal(out[i], imgs[i])
    n = ck.count_params(ck.safety_checker_shapes())
    assert 300e6 < n < 310e6, n  # CLIP ViT-L/14 vision tower + projection: 304 M
