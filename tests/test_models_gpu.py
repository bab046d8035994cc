"""GPU parity of the model graphs (UNet, ControlNet, VAE, CLIP text) and the full pipelines against the CPU
fp32 oracle restatement on identical random-init weights, token ids and generator seeds.

Stated tolerance ("calibrate, don't guess"): the same oracle graph is also run in stock torch
bf16 on the GPU; our error vs the fp32 oracle must be <= max(2 x torch-bf16's error, 1e-2 x max|ref|)
for tensors, and image PSNR >= min(torch-bf16 PSNR - 1 dB, 40 dB)."""
import copy
import math

import numpy as np
import pytest
import torch

from oracle.diffusers_restated import models as om
from oracle.diffusers_restated.pipelines import OraclePipeline
from saspa_aug_b200 import checkpoints as ck
from saspa_aug_b200 import nn as snn
from saspa_aug_b200 import ops
from saspa_aug_b200.pipelines import SaspaControlNetPipeline, random_state_dicts
from saspa_aug_b200.synthetic import synthetic_source, synthetic_token_ids

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _ocfg(c):
    return om.UNetConfig(**{k: getattr(c, k) for k in om.UNetConfig.__dataclass_fields__})


def _vcfg(c):
    return om.VAEConfig(**{k: getattr(c, k) for k in om.VAEConfig.__dataclass_fields__})


def _text_model(tcfg, sd):
    from transformers import CLIPTextConfig, CLIPTextModel

    m = CLIPTextModel(CLIPTextConfig(vocab_size=tcfg.vocab_size, hidden_size=tcfg.hidden_size, intermediate_size=tcfg.intermediate_size,
                                     num_hidden_layers=tcfg.num_hidden_layers, num_attention_heads=tcfg.num_attention_heads,
                                     max_position_embeddings=tcfg.max_position_embeddings, hidden_act=tcfg.hidden_act,
                                     layer_norm_eps=tcfg.layer_norm_eps, bos_token_id=tcfg.vocab_size - 2, eos_token_id=tcfg.vocab_size - 1))
    missing, unexpected = m.load_state_dict(sd, strict=False)
    assert not unexpected and all("position_ids" in k for k in missing), (missing, unexpected)
    return m.eval()


def _bound(mine, ref32, refbf, floor=1e-2):
    e_mine = (mine - ref32).abs().max().item()
    e_bf = (refbf - ref32).abs().max().item()
    lim = max(2.0 * e_bf, floor * ref32.abs().max().item())
    assert math.isfinite(e_mine) and e_mine <= lim, f"ours {e_mine:.4g} vs torch-bf16 {e_bf:.4g} (limit {lim:.4g}, max|ref| {ref32.abs().max().item():.4g})"
    return e_mine, e_bf


def _unet_controlnet_case(ucfg_p, n, h, w, seed, text_dim, pooled_dim=0):
    g = torch.Generator().manual_seed(seed)
    usd = ck.random_state_dict(ck.unet_shapes(ucfg_p), seed)
    csd = ck.random_state_dict(ck.controlnet_shapes(ucfg_p), seed + 1)
    ou, oc = om.UNet2DConditionModel(_ocfg(ucfg_p)).eval(), om.ControlNetModel(_ocfg(ucfg_p)).eval()
    ou.load_state_dict(usd)
    oc.load_state_dict(csd)
    lat = torch.randn((n, 4, h, w), generator=g)
    text = torch.randn((n, 77, text_dim), generator=g)
    cond = (torch.rand((n, 3, 8 * h, 8 * w), generator=g) > 0.9).float()
    t = 481.0
    scale = 0.75
    added = addedb = addedm = None
    if pooled_dim:  # SDXL "text_time" conditioning
        pooled = torch.randn((n, pooled_dim), generator=g)
        tid = torch.tensor([[8.0 * h, 8.0 * w, 0, 0, 8.0 * h, 8.0 * w]]).repeat(n, 1)
        added = {"text_embeds": pooled, "time_ids": tid}
        addedb = {"text_embeds": pooled.to(DEV, torch.bfloat16), "time_ids": tid.to(DEV)}
        addedm = {"text_embeds": pooled.to(DEV, torch.bfloat16), "time_ids": tid.to(DEV)}
    with torch.no_grad():
        d32, m32 = oc(lat, t, text, cond, scale, added_cond_kwargs=added)
        e32 = ou(lat, t, text, d32, m32, added_cond_kwargs=added)
        oub, ocb = copy.deepcopy(ou).to(DEV, torch.bfloat16), copy.deepcopy(oc).to(DEV, torch.bfloat16)
        lb, tb, cb = lat.to(DEV, torch.bfloat16), text.to(DEV, torch.bfloat16), cond.to(DEV, torch.bfloat16)
        db, mb = ocb(lb, t, tb, cb, scale, added_cond_kwargs=addedb)
        ebf = oub(lb, t, tb, db, mb, added_cond_kwargs=addedb).float().cpu()
        del oub, ocb
    unet, cn = snn.UNet(usd, ucfg_p, torch.device(DEV)), snn.ControlNet(csd, ucfg_p, torch.device(DEV))
    textb = text.to(DEV, torch.bfloat16)
    kv_u, kv_c = unet.text_kv(textb), cn.text_kv(textb)
    x2 = ops.nchw_f32_to_nhwc_bf16(lat.to(DEV))
    tv = torch.full((n,), t, dtype=torch.float32, device=DEV)
    ce = cn.cond_embedding(cond.permute(0, 2, 3, 1).contiguous().to(DEV, torch.bfloat16))
    aug_u, aug_c = unet.added_embed(addedm), cn.added_embed(addedm)
    st = unet.encode(x2, unet.time_embed(tv, aug_u), kv_u, n, h, w)
    cn.inject(x2, cn.time_embed(tv, aug_c), kv_c, ce, scale, st)
    eps = ops.nhwc_to_nchw_f32(unet.decode(st, unet.time_embed(tv, aug_u), kv_u)).cpu()
    torch.cuda.synchronize()
    return eps, e32, ebf


def test_unet_controlnet_tiny(cuda_device):
    eps, e32, ebf = _unet_controlnet_case(ck.UNetConfig.tiny(), 2, 16, 16, 7, 64)
    _bound(eps, e32, ebf)


def test_unet_controlnet_tiny_nonsquare_batch3(cuda_device):
    eps, e32, ebf = _unet_controlnet_case(ck.UNetConfig.tiny(), 3, 8, 24, 8, 64)
    _bound(eps, e32, ebf)


def test_unet_controlnet_sd15_full_size(cuda_device):
    """The real SD v1.5 + ControlNet-canny architecture at 512x512 (latent 64x64), one CFG pair."""
    torch.set_num_threads(max(torch.get_num_threads(), 8))
    eps, e32, ebf = _unet_controlnet_case(ck.UNetConfig.sd15(), 2, 64, 64, 11, 768)
    e_mine, e_bf = _bound(eps, e32, ebf)
    print(f"sd15 UNet+ControlNet eps max-abs err: ours {e_mine:.4g}, torch-bf16 {e_bf:.4g}, max|eps| {e32.abs().max().item():.4g}")


def test_unet_controlnet_tiny_xl(cuda_device):
    """SDXL topology (DownBlock first, transformer depth 1/2/3, linear projections, text_time added conditioning), non-square."""
    eps, e32, ebf = _unet_controlnet_case(ck.UNetConfig.tiny_xl(), 3, 16, 24, 12, 192, pooled_dim=96)
    _bound(eps, e32, ebf)


def test_unet_controlnet_sdxl_full_size(cuda_device):
    """The real SDXL UNet (2.57 B) + ControlNet-canny-sdxl (1.25 B) architectures at 512x512 (latent 64x64; sd_xl-turbo's
    training resolution and the reference's RESOLUTION, run_aug.py:536), batch 1, no CFG."""
    torch.set_num_threads(max(torch.get_num_threads(), 8))
    eps, e32, ebf = _unet_controlnet_case(ck.UNetConfig.sdxl(), 1, 64, 64, 13, 2048, pooled_dim=1280)
    e_mine, e_bf = _bound(eps, e32, ebf)
    print(f"sdxl UNet+ControlNet eps max-abs err: ours {e_mine:.4g}, torch-bf16 {e_bf:.4g}, max|eps| {e32.abs().max().item():.4g}")


@pytest.mark.parametrize("vcfg_name", ["tiny", "sd15"])
def test_vae_decode_encode(cuda_device, vcfg_name):
    vcfg = ck.VAEConfig.tiny() if vcfg_name == "tiny" else ck.VAEConfig.sd15()
    sd = ck.random_state_dict(ck.vae_shapes(vcfg), 21)
    o = om.AutoencoderKL(_vcfg(vcfg)).eval()
    o.load_state_dict(sd)
    g = torch.Generator().manual_seed(3)
    hw = 16 if vcfg_name == "tiny" else 32
    z = torch.randn((2, 4, hw, hw), generator=g)
    img = torch.rand((2, 3, 8 * hw, 8 * hw), generator=g) * 2 - 1
    with torch.no_grad():
        d32 = o.decode(z)
        mean32, lv32 = o.encode_moments(img)
        ob = copy.deepcopy(o).to(DEV, torch.bfloat16)
        dbf = ob.decode(z.to(DEV, torch.bfloat16)).float().cpu()
        mb, lb = ob.encode_moments(img.to(DEV, torch.bfloat16))
    dec, enc = snn.VAEDecoder(sd, vcfg, torch.device(DEV)), snn.VAEEncoder(sd, vcfg, torch.device(DEV))
    mine = dec(ops.nchw_f32_to_nhwc_bf16(z.to(DEV)))
    _bound(mine.permute(0, 3, 1, 2).cpu(), d32, dbf)
    mom = enc(img.permute(0, 2, 3, 1).contiguous().to(DEV, torch.bfloat16)).permute(0, 3, 1, 2).cpu()
    _bound(mom[:, :4], mean32, mb.float().cpu())
    _bound(mom[:, 4:].clamp(-30, 20), lv32, lb.float().cpu())


@pytest.mark.parametrize("name", ["tiny", "sd15"])
def test_clip_text_encoder(cuda_device, name):
    tcfg = ck.CLIPTextConfig.tiny() if name == "tiny" else ck.CLIPTextConfig.sd15()
    sd = ck.random_state_dict(ck.clip_text_shapes(tcfg), 31)
    ref = _text_model(tcfg, sd)
    ids = synthetic_token_ids(5, batch=3, vocab=tcfg.vocab_size)
    with torch.no_grad():
        r32 = ref(ids)[0]
        rbf = copy.deepcopy(ref).to(DEV, torch.bfloat16)(ids.to(DEV))[0].float().cpu()
    te = snn.CLIPTextEncoder(sd, torch.device(DEV), tcfg.num_attention_heads, tcfg.hidden_act, tcfg.layer_norm_eps)
    mine = te(ids.to(DEV)).float().cpu()
    _bound(mine, r32, rbf)


@pytest.mark.parametrize("name", ["tiny_g", "sdxl_g"])
def test_clip_text_with_projection(cuda_device, name):
    """SDXL text_encoder_2 (CLIPTextModelWithProjection, gelu): hidden_states[-2] and pooled text_embeds vs transformers."""
    from tests.test_sdxl_cpu import text_models

    tcfg = ck.CLIPTextConfig.tiny_g() if name == "tiny_g" else ck.CLIPTextConfig.sdxl_g()
    sd = ck.random_state_dict(ck.clip_text_shapes(tcfg), 33)
    _, ref = text_models(ck.CLIPTextConfig.tiny(), tcfg, ck.random_state_dict(ck.clip_text_shapes(ck.CLIPTextConfig.tiny()), 1), sd)
    ids = synthetic_token_ids(6, batch=3, vocab=tcfg.vocab_size)
    with torch.no_grad():
        o32 = ref(ids, output_hidden_states=True)
        obf = copy.deepcopy(ref).to(DEV, torch.bfloat16)(ids.to(DEV), output_hidden_states=True)
    te = snn.CLIPTextEncoder(sd, torch.device(DEV), tcfg.num_attention_heads, tcfg.hidden_act, tcfg.layer_norm_eps)
    pen, pooled = te(ids.to(DEV), penultimate=True, pooled=True)
    _bound(pen.float().cpu(), o32.hidden_states[-2], obf.hidden_states[-2].float().cpu())
    _bound(pooled.float().cpu(), o32[0], obf[0].float().cpu())
    last = te(ids.to(DEV)).float().cpu()
    _bound(last, o32.last_hidden_state, obf.last_hidden_state.float().cpu())


def _psnr(a, b):
    mse = ((a.astype(np.float64) - b.astype(np.float64)) ** 2).mean()
    return 99.0 if mse == 0 else 10 * math.log10(255.0 ** 2 / mse)


def _check_pipeline(out, img_o, lat_o, img_b, lat_b, tag):
    """Calibrated tolerance (module docstring): per-step latent max-abs error <= max(2 x stock-torch-bf16's error on the same graph,
    1e-2 x max|latent|); image PSNR >= min(torch-bf16's PSNR - 3 dB, 40 dB).  (The recurrence amplifies bf16 rounding and tiny
    random-init nets are not contractive, so an absolute bound would be a guess.)"""
    assert len(out.latents_per_step) == len(lat_o)
    for k, (a, b, c) in enumerate(zip(out.latents_per_step, lat_o, lat_b)):
        e_mine, e_bf = (a.cpu() - b).abs().max().item(), (c - b).abs().max().item()
        lim = max(2.0 * e_bf, 1e-2 * b.abs().max().item())
        assert math.isfinite(e_mine) and e_mine <= lim, f"{tag} step {k}: ours {e_mine:.4g} vs torch-bf16 {e_bf:.4g} (max|x| {b.abs().max().item():.4g})"
    p, p_bf = _psnr(np.stack(out.images), img_o), _psnr(img_b, img_o)
    print(f"{tag}: final-latent err ours {e_mine:.4g} / torch-bf16 {e_bf:.4g}; image PSNR ours {p:.1f} dB / torch-bf16 {p_bf:.1f} dB")
    assert p >= min(p_bf - 3.0, 40.0), (p, p_bf)


@pytest.mark.parametrize("mode,sampler,steps,strength", [("t2i", "ddim", 6, 1.0), ("img2img", "unipc", 10, 0.5), ("t2i", "pndm", 5, 1.0)])
def test_pipeline_tiny_matches_oracle(cuda_device, mode, sampler, steps, strength):
    """Whole pipeline (tiny same-topology models, 128x128): per-step latents and final image vs the fp32 oracle."""
    sds = random_state_dicts("tiny", 100)
    ucfg, vcfg, tcfg = ck.UNetConfig.tiny(), ck.VAEConfig.tiny(), ck.CLIPTextConfig.tiny()
    ou, oc, ov = om.UNet2DConditionModel(_ocfg(ucfg)), om.ControlNetModel(_ocfg(ucfg)), om.AutoencoderKL(_vcfg(vcfg))
    ou.load_state_dict(sds["unet"]); oc.load_state_dict(sds["controlnet"]); ov.load_state_dict(sds["vae"])
    opipe = OraclePipeline(ou, oc, ov, _text_model(tcfg, sds["text"]), sampler)
    src = np.stack([synthetic_source(s, 128, 128) for s in (1, 2)])
    from oracle import clib
    ctrl = np.repeat(clib.canny(src, 120, 200)[..., None], 3, axis=3)
    ids = synthetic_token_ids(9, batch=2, vocab=tcfg.vocab_size)
    nids = synthetic_token_ids(10, batch=1, vocab=tcfg.vocab_size).expand(2, -1)
    kw = dict(num_inference_steps=steps, guidance_scale=7.5, strength=strength, controlnet_conditioning_scale=0.75)
    img_o, lat_o = opipe(ids, nids, ctrl, src if mode == "img2img" else None, generator=torch.Generator().manual_seed(1), **kw)
    pipe = SaspaControlNetPipeline.from_state_dicts(sds["unet"], sds["controlnet"], sds["vae"], sds["text"], unet_cfg=ucfg, vae_cfg=vcfg, text_cfg=tcfg,
                                                    sampler=sampler)
    call = dict(prompt_ids=ids, negative_prompt_ids=nids, generator=torch.Generator().manual_seed(1), return_latents_per_step=True, output_type="np", **kw)
    if mode == "img2img":
        out = pipe(image=src, control_image=ctrl, **call)
    else:
        out = pipe(image=ctrl, **call)
    img_b, lat_b = opipe.to(DEV, torch.bfloat16)(ids, nids, ctrl, src if mode == "img2img" else None, generator=torch.Generator().manual_seed(1), **kw)
    _check_pipeline(out, img_o, lat_o, img_b, lat_b, f"{mode}/{sampler}")


@pytest.mark.parametrize("mode,sampler,steps,strength,gs", [("t2i", "ddim_sdxl_turbo", 4, 1.0, 0.0), ("img2img", "ddim_sdxl_turbo", 4, 0.5, 0.0),
                                                            ("t2i", "unipc_sdxl_turbo", 3, 1.0, 5.0)])
def test_pipeline_tiny_xl_matches_oracle(cuda_device, mode, sampler, steps, strength, gs):
    """SD-XL(-turbo) pipeline (tiny same-topology models, two text encoders, added conditioning, 128x192) vs the fp32 oracle:
    turbo settings (guidance 0 => no CFG, DDIM trailing with the clip_sample clamp), SDEdit, and a CFG run (negative pooled path)."""
    from oracle.diffusers_restated.pipelines import OracleSDXLPipeline
    from saspa_aug_b200.pipelines import SaspaSDXLControlNetPipeline, sdxl_configs
    from tests.test_sdxl_cpu import text_models

    sds = random_state_dicts("tiny_xl", 200)
    ucfg, vcfg, tcfg, t2cfg = sdxl_configs("tiny_xl")
    ou, oc, ov = om.UNet2DConditionModel(_ocfg(ucfg)), om.ControlNetModel(_ocfg(ucfg)), om.AutoencoderKL(_vcfg(vcfg))
    ou.load_state_dict(sds["unet"]); oc.load_state_dict(sds["controlnet"]); ov.load_state_dict(sds["vae"])
    t1, t2 = text_models(tcfg, t2cfg, sds["text"], sds["text2"])
    opipe = OracleSDXLPipeline(ou, oc, ov, t1, t2, sampler)
    src = np.stack([synthetic_source(s, 128, 192) for s in (3, 4)])
    from oracle import clib
    ctrl = np.repeat(clib.canny(src, 120, 200)[..., None], 3, axis=3)
    ids = (synthetic_token_ids(9, batch=2, vocab=tcfg.vocab_size), synthetic_token_ids(11, batch=2, vocab=t2cfg.vocab_size))
    nids = (synthetic_token_ids(10, batch=1, vocab=tcfg.vocab_size).expand(2, -1), synthetic_token_ids(12, batch=1, vocab=t2cfg.vocab_size).expand(2, -1))
    kw = dict(num_inference_steps=steps, guidance_scale=gs, strength=strength, controlnet_conditioning_scale=0.75)
    img_o, lat_o = opipe(ids, nids, ctrl, src if mode == "img2img" else None, generator=torch.Generator().manual_seed(1), **kw)
    pipe = SaspaSDXLControlNetPipeline.from_state_dicts(sds["unet"], sds["controlnet"], sds["vae"], sds["text"], sds["text2"], unet_cfg=ucfg, vae_cfg=vcfg,
                                                        text_cfg=tcfg, text2_cfg=t2cfg, sampler=sampler)
    call = dict(prompt_ids=ids, negative_prompt_ids=nids, generator=torch.Generator().manual_seed(1), return_latents_per_step=True, output_type="np", **kw)
    out = pipe(image=src, control_image=ctrl, **call) if mode == "img2img" else pipe(image=ctrl, **call)
    img_b, lat_b = opipe.to(DEV, torch.bfloat16)(ids, nids, ctrl, src if mode == "img2img" else None, generator=torch.Generator().manual_seed(1), **kw)
    _check_pipeline(out, img_o, lat_o, img_b, lat_b, f"sdxl {mode}/{sampler}")
