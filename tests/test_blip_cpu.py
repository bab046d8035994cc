"""CPU checks of the BLIP-Diffusion front end (BASELINE config 3): the oracle restatement (oracle/diffusers_restated/blip.py) is
cross-checked against the installed ``transformers`` BLIP-2 / CLIP building blocks on identical random weights (diffusers'
modeling_blip2.py is built from exactly these blocks), the product's checkpoint layout loads strictly into the oracle, and the host
logic of the drop-in call (prompt building, tokenizer lengths, kwargs of run_aug.py:243-250,268-271) is exercised without a GPU."""
import numpy as np
import pytest
import torch

from oracle.diffusers_restated import blip as ob
from saspa_aug_b200 import checkpoints as ck


def _ocfg(c):
    return ob.Blip2Config(**{k: getattr(c, k) for k in ob.Blip2Config.__dataclass_fields__})


def _qformer(cfg_p, seed=5):
    sd = ck.random_state_dict(ck.qformer_shapes(cfg_p), seed)
    m = ob.Blip2QFormerModel(_ocfg(cfg_p)).eval()
    m.load_state_dict(sd, strict=True)  # the product's key layout == the oracle's == diffusers' checkpoint layout
    return m, sd


def test_qformer_shapes_load_strictly_and_configs_agree():
    for name in ("tiny", "blipdiffusion"):
        c, o = getattr(ck.Blip2Config, name)(), getattr(ob.Blip2Config, name)()
        assert {k: getattr(c, k) for k in ob.Blip2Config.__dataclass_fields__} == o.__dict__
    _qformer(ck.Blip2Config.tiny())
    n = ck.count_params(ck.qformer_shapes(ck.Blip2Config.blipdiffusion()))
    assert 380e6 < n < 520e6, n  # 23-layer ViT-L/14 (~290 M) + 12-layer Q-Former with text branch + embeddings


def test_oracle_vision_layers_match_transformers():
    from transformers import Blip2VisionConfig
    from transformers.models.blip_2 import modeling_blip_2 as mb

    cfg = ck.Blip2Config.tiny()
    m, sd = _qformer(cfg)
    vc = Blip2VisionConfig(hidden_size=cfg.vision_hidden_size, intermediate_size=cfg.vision_intermediate_size, num_hidden_layers=cfg.vision_num_hidden_layers,
                           num_attention_heads=cfg.vision_num_attention_heads, image_size=cfg.image_size, patch_size=cfg.patch_size, hidden_act="quick_gelu",
                           layer_norm_eps=cfg.vision_layer_norm_eps, qkv_bias=True, attention_dropout=0.0)
    vc._attn_implementation = "eager"
    enc = mb.Blip2Encoder(vc).eval()
    enc.load_state_dict({k[len("visual_encoder.encoder."):]: v for k, v in sd.items() if k.startswith("visual_encoder.encoder.")}, strict=True)
    emb = mb.Blip2VisionEmbeddings(vc).eval()
    esd = {k[len("visual_encoder.embeddings."):]: v for k, v in sd.items() if k.startswith("visual_encoder.embeddings.")}
    esd["patch_embedding.bias"] = torch.zeros(cfg.vision_hidden_size)  # diffusers' Blip2VisionEmbeddings has bias=False; transformers' has one
    emb.load_state_dict(esd, strict=True)
    x = torch.randn((2, 3, cfg.image_size, cfg.image_size), generator=torch.Generator().manual_seed(0))
    with torch.no_grad():
        e = emb(x)
        assert torch.allclose(e, m.visual_encoder.embeddings(x), atol=1e-5)
        h = m.visual_encoder.pre_layernorm(e)
        want = enc(inputs_embeds=h)[0]
        got = h
        for l in m.visual_encoder.encoder.layers:
            got = l(got)
    assert torch.allclose(got, want, atol=2e-5), (got - want).abs().max()


def test_oracle_qformer_layers_match_transformers():
    from transformers import Blip2QFormerConfig
    from transformers.models.blip_2 import modeling_blip_2 as mb

    cfg = ck.Blip2Config.tiny()
    m, sd = _qformer(cfg)
    qc = Blip2QFormerConfig(vocab_size=cfg.vocab_size, hidden_size=cfg.hidden_size, num_hidden_layers=cfg.num_hidden_layers,
                            num_attention_heads=cfg.num_attention_heads, intermediate_size=cfg.intermediate_size,
                            max_position_embeddings=cfg.max_position_embeddings, layer_norm_eps=cfg.layer_norm_eps,
                            cross_attention_frequency=cfg.cross_attention_frequency, encoder_hidden_size=cfg.vision_hidden_size,
                            hidden_dropout_prob=0.0, attention_probs_dropout_prob=0.0, use_qformer_text_input=True)
    qc._attn_implementation = "eager"
    qe = mb.Blip2QFormerEncoder(qc).eval()
    qe.load_state_dict({k[len("encoder."):]: v for k, v in sd.items() if k.startswith("encoder.")}, strict=True)
    g = torch.Generator().manual_seed(1)
    nq = cfg.num_query_tokens
    h = torch.randn((2, nq + 5, cfg.hidden_size), generator=g)
    img = torch.randn((2, 17, cfg.vision_hidden_size), generator=g)
    with torch.no_grad():
        want = qe(h, encoder_hidden_states=img, query_length=nq)[0]
        got = h
        for l in m.encoder.layer:
            got = l(got, img, nq)
    assert torch.allclose(got, want, atol=2e-5), (got - want).abs().max()


def test_oracle_ctx_clip_matches_plain_clip_without_ctx_and_splices_with_ctx():
    from tests.test_models_gpu import _text_model

    tcfg = ck.CLIPTextConfig.tiny()
    sd = ck.random_state_dict(ck.clip_text_shapes(tcfg), 31)
    clip = _text_model(tcfg, sd)
    ctx_model = ob.ContextCLIPTextModel(clip)
    g = torch.Generator().manual_seed(2)
    ids = torch.randint(0, tcfg.vocab_size, (2, 77), generator=g)
    with torch.no_grad():
        assert torch.allclose(ctx_model(ids), clip(ids)[0], atol=2e-5)
        # with ctx: equals the plain encoder run on the spliced embedding sequence (inputs_embeds is not exposed by CLIPTextModel,
        # so compare against a second, independent formulation: hooks replacing the token embedding output)
        q = torch.randn((2, 16, tcfg.hidden_size), generator=g)
        short = ids[:, :61]
        out = ctx_model(short, q, [2, 2])
        assert out.shape == (2, 77, tcfg.hidden_size)
        tok = clip.text_model.embeddings.token_embedding
        spliced = torch.cat([tok(short)[:, :2], q, tok(short)[:, 2:]], 1)
        hook = tok.register_forward_hook(lambda mod, inp, o: spliced)
        try:
            want = clip(ids)[0]  # ids only set the sequence length / positions here; embeddings come from the hook
        finally:
            hook.remove()
    assert torch.allclose(out, want, atol=2e-5), (out - want).abs().max()


def test_blip_preprocess_is_pil_bicubic_then_clip_normalise():
    from PIL import Image

    rng = np.random.default_rng(0)
    img = rng.integers(0, 256, (1, 96, 80, 3), dtype=np.uint8)
    x = ob.blip_preprocess_reference(img, 56)
    r = np.asarray(Image.fromarray(img[0]).resize((56, 56), resample=Image.BICUBIC)).astype(np.float32) / 255.0
    want = (r - np.array(ob.OPENAI_CLIP_MEAN, np.float32)) / np.array(ob.OPENAI_CLIP_STD, np.float32)
    assert x.shape == (1, 3, 56, 56) and np.allclose(x[0].permute(1, 2, 0).numpy(), want, atol=1e-6)


def test_build_prompt_and_tokenizers_host_logic():
    from saspa_aug_b200.pipelines import SyntheticBertTokenizer, SyntheticTokenizer, build_blip_prompt

    p = build_blip_prompt(["  flying over a city "], ["airplane"], 1.0, 20)
    assert p == ob.build_prompt(["  flying over a city "], ["airplane"], 1.0, 20)
    assert p[0].count("a airplane flying over a city") == 20 and p[0].count(", ") == 19
    assert build_blip_prompt(["x"], ["y"], 0.5, 20)[0].count("a y x") == 10
    tok = SyntheticTokenizer(1000, 77)
    ids = tok(p, max_length=77 - 16)
    assert ids.shape == (1, 61) and ids[0, 0] == 998 and ids[0, -1] == 999  # truncated: BOS ... EOS
    assert tok(["short prompt"]).shape == (1, 77)
    bt = SyntheticBertTokenizer(500, 32)
    s = bt(["airplane", "fighter jet"])
    assert [len(x) for x in s] == [3, 4] and all(int(x[0]) == 101 and int(x[-1]) == 102 and int(x.max()) < 500 for x in s)


def test_pass_thorugh_pipe_blip_kwargs():
    """run_aug.py:243-250,268-271: BLIP drops negative_prompt for neg_prompt, passes reference/condtioning images and the control size."""
    from PIL import Image

    from saspa_aug_b200 import run_aug as ra

    seen = {}

    class FakePipe:
        def __call__(self, **kw):
            seen.update(kw)

            class O:
                images = ["img"]

            return O()

    src, ctrl = Image.new("RGB", (64, 48)), Image.new("RGB", (128, 96))
    out = ra.pass_thorugh_pipe("blip_diffusion", FakePipe(), "a photo", src, False, 0.5, 20, None, 7.5, 0.75, control_image=ctrl,
                               blip_src_category="airplane", blip_target_category="airplane")
    assert out == "img"
    assert set(seen) == {"prompt", "num_inference_steps", "generator", "guidance_scale", "reference_image", "source_subject_category",
                         "target_subject_category", "height", "width", "neg_prompt", "condtioning_image"}
    assert seen["height"] == 96 and seen["width"] == 128 and seen["neg_prompt"] == ra.NEGATIVE_PROMPT and seen["reference_image"] is src


def test_oracle_blip_pipeline_tiny_runs():
    from oracle.diffusers_restated import models as om
    from oracle.diffusers_restated.pipelines import OracleBlipPipeline
    from saspa_aug_b200.pipelines import blip_configs, random_state_dicts
    from saspa_aug_b200.synthetic import synthetic_source, synthetic_token_ids
    from tests.test_models_gpu import _ocfg as _ucfg
    from tests.test_models_gpu import _text_model, _vcfg

    sds = random_state_dicts("tiny_blip", 300)
    ucfg, vcfg, tcfg, qcfg = blip_configs("tiny_blip")
    ou, oc, ov = om.UNet2DConditionModel(_ucfg(ucfg)), om.ControlNetModel(_ucfg(ucfg)), om.AutoencoderKL(_vcfg(vcfg))
    ou.load_state_dict(sds["unet"]); oc.load_state_dict(sds["controlnet"]); ov.load_state_dict(sds["vae"])
    qf = ob.Blip2QFormerModel(_ocfg(qcfg))
    qf.load_state_dict(sds["qformer"])
    pipe = OracleBlipPipeline(ou, oc, ov, ob.ContextCLIPTextModel(_text_model(tcfg, sds["text"])), qf)
    src = np.stack([synthetic_source(5, 64, 64)])
    ctrl = np.zeros((1, 64, 64, 3), np.uint8)
    ctrl[:, 20:40, 30] = 255
    ids = synthetic_token_ids(9, batch=1, vocab=tcfg.vocab_size)[:, :61]
    nids = synthetic_token_ids(10, batch=1, vocab=tcfg.vocab_size)
    img, lats, query, text = pipe(ids, nids, torch.tensor([[101, 150, 102]]), src, ctrl, generator=torch.Generator().manual_seed(1), num_inference_steps=3)
    assert img.shape == (1, 64, 64, 3) and img.dtype == np.uint8 and len(lats) == 4  # PLMS: the second timestep is visited twice (skip_prk_steps)
    assert query.shape == (1, 16, 64) and text.shape == (2, 77, 64) and torch.isfinite(lats[-1]).all()
