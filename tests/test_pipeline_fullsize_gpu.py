"""Full-size pipeline parity on the BASELINE configs: the REAL architectures (SD v1.5 / BLIP-Diffusion / SD-XL UNet +
ControlNet-canny + VAE + text encoders, random-init) at the configs' resolutions and step counts, through the reference's call path
(run_aug/run_aug.py:233-279 -> pipe(**pipe_args)), against the CPU fp32 oracle restatement of diffusers 0.32.2 on identical weights,
token ids and generator seeds.

Stated tolerance, calibrated on the same oracle graph run in stock torch on the GPU in bf16 AND fp16 (the reference's
dtype, run_aug.py:323): per-step latent max-abs error <= max(2 x g_t, 1e-2 x max|latent_t|) with g_t the stock-torch-bf16 gap, and image
PSNR >= min(p - 1 dB, 40 dB) with p the stock-torch-bf16 PSNR.  The fp16 gap / PSNR are measured and printed beside it (a random-init
net is allowed to overflow fp16; then that column reads "overflow").  Every number lands in gpurun_out/r2_fullsize_parity.txt."""
import math
import os
import time

import numpy as np
import pytest
import torch

from oracle import clib
from oracle.diffusers_restated import models as om
from oracle.diffusers_restated.pipelines import OracleBlipPipeline, OraclePipeline, OracleSDXLPipeline
from saspa_aug_b200 import checkpoints as ck
from saspa_aug_b200.pipelines import (SaspaBlipControlNetPipeline, SaspaControlNetPipeline, SaspaSDXLControlNetPipeline, blip_configs, random_state_dicts,
                                      sdxl_configs)
from saspa_aug_b200.synthetic import synthetic_source, synthetic_token_ids
from tests.test_models_gpu import _ocfg, _psnr, _text_model, _vcfg

pytestmark = pytest.mark.gpu
DEV = "cuda"
LOG = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out", "r2_fullsize_parity.txt")


def _log(line):
    print(line)
    try:
        os.makedirs(os.path.dirname(LOG), exist_ok=True)
        with open(LOG, "a") as f:
            f.write(line + "\n")
    except OSError:
        pass


def _judge(tag, out_latents, out_images, ref, bf, fp16):
    """ref / bf / fp16 = (image u8, per-step latents) of the fp32-CPU, bf16-GPU and fp16-GPU oracle runs (fp16 may be None)."""
    img_o, lat_o = ref
    img_b, lat_b = bf
    assert len(out_latents) == len(lat_o)
    worst = 0.0
    for k, (a, b, c) in enumerate(zip(out_latents, lat_o, lat_b)):
        e_mine, g = (a.cpu() - b).abs().max().item(), (c - b).abs().max().item()
        lim = max(2.0 * g, 1e-2 * b.abs().max().item())
        g16 = "-"
        if fp16 is not None:
            d = (fp16[1][k] - b).abs().max().item()
            g16 = f"{d:.4g}" if math.isfinite(d) else "overflow"
        _log(f"{tag} step {k:2d}: latent max-abs err ours {e_mine:.4g} | torch-bf16 {g:.4g} | torch-fp16 {g16} | limit {lim:.4g} | max|x| {b.abs().max().item():.4g}")
        assert math.isfinite(e_mine) and e_mine <= lim, f"{tag} step {k}: ours {e_mine:.4g} > limit {lim:.4g} (torch-bf16 {g:.4g})"
        worst = max(worst, e_mine / lim)
    p, p_bf = _psnr(np.stack(out_images), img_o), _psnr(img_b, img_o)
    p16 = "-" if fp16 is None else (f"{_psnr(fp16[0], img_o):.2f}" if np.isfinite(fp16[0].astype(np.float64)).all() else "overflow")
    _log(f"{tag}: image PSNR vs fp32 oracle: ours {p:.2f} dB | torch-bf16 {p_bf:.2f} dB | torch-fp16 {p16} dB | gate {min(p_bf - 1.0, 40.0):.2f} dB | worst latent err / limit {worst:.2f}")
    assert p >= min(p_bf - 1.0, 40.0), (p, p_bf)


def _fp16_run(opipe, fn):
    try:
        r = fn(opipe.to(DEV, torch.float16))
        return r if all(torch.isfinite(x).all() for x in r[1]) else None
    except Exception as e:  # noqa: BLE001 -- calibration column only
        _log(f"torch-fp16 calibration run failed: {e!r}")
        return None


def test_config1_sd15_img2img_unipc_full_size(cuda_device):
    """BASELINE config 1: one synthetic 512x512 source, cv2.Canny-equivalent + SD v1.5 ControlNet-canny img2img, strength 0.5,
    20 UniPC steps (10 executed), CFG 7.5, conditioning scale 0.75, generator seed 1 -- per-step latents + image vs the fp32 CPU oracle."""
    torch.set_num_threads(os.cpu_count() or 8)
    sds = random_state_dicts("sd15", 1234)
    ucfg, vcfg, tcfg = ck.UNetConfig.sd15(), ck.VAEConfig.sd15(), ck.CLIPTextConfig.sd15()
    ou, oc, ov = om.UNet2DConditionModel(_ocfg(ucfg)), om.ControlNetModel(_ocfg(ucfg)), om.AutoencoderKL(_vcfg(vcfg))
    ou.load_state_dict(sds["unet"]); oc.load_state_dict(sds["controlnet"]); ov.load_state_dict(sds["vae"])
    opipe = OraclePipeline(ou, oc, ov, _text_model(tcfg, sds["text"]), "unipc")
    src = synthetic_source(0, 512, 512)[None]
    ctrl = np.repeat(clib.canny(src, 120, 200)[..., None], 3, axis=3)
    ids, nids = synthetic_token_ids(0, vocab=tcfg.vocab_size), synthetic_token_ids(1, vocab=tcfg.vocab_size)
    kw = dict(num_inference_steps=20, guidance_scale=7.5, strength=0.5, controlnet_conditioning_scale=0.75)
    t0 = time.perf_counter()
    ref = opipe(ids, nids, ctrl, src, generator=torch.Generator().manual_seed(1), **kw)
    _log(f"config1: fp32 CPU oracle {time.perf_counter() - t0:.1f} s on {torch.get_num_threads()} threads, {len(ref[1])} executed steps")
    assert len(ref[1]) == 10
    bf = opipe.to(DEV, torch.bfloat16)(ids, nids, ctrl, src, generator=torch.Generator().manual_seed(1), **kw)
    fp16 = _fp16_run(opipe, lambda p: p(ids, nids, ctrl, src, generator=torch.Generator().manual_seed(1), **kw))
    del opipe, ou, oc, ov
    pipe = SaspaControlNetPipeline.from_state_dicts(sds["unet"], sds["controlnet"], sds["vae"], sds["text"], sampler="unipc")
    out = pipe(image=src, control_image=ctrl, prompt_ids=ids, negative_prompt_ids=nids, generator=torch.Generator().manual_seed(1),
               return_latents_per_step=True, output_type="np", **kw)
    _judge("config1 sd15 img2img/unipc 512^2", out.latents_per_step, out.images, ref, bf, fp16)


def test_config3_blip_diffusion_plms_full_size(cuda_device):
    """BASELINE config 3 (one image of the batch): BLIP-Diffusion + ControlNet-canny, full-size Q-Former (12 L) + ViT-L vision tower +
    ctx-CLIP + SD v1.5-shaped UNet/ControlNet/VAE, 512x512, 20 PNDM/PLMS steps, CFG 7.5 (run_aug.py:243-250,268-271)."""
    from oracle.diffusers_restated import blip as ob
    from tests.test_blip_cpu import _ocfg as _qcfg

    torch.set_num_threads(os.cpu_count() or 8)
    sds = random_state_dicts("blip", 1234)
    ucfg, vcfg, tcfg, qcfg = blip_configs("blip")
    ou, oc, ov = om.UNet2DConditionModel(_ocfg(ucfg)), om.ControlNetModel(_ocfg(ucfg)), om.AutoencoderKL(_vcfg(vcfg))
    ou.load_state_dict(sds["unet"]); oc.load_state_dict(sds["controlnet"]); ov.load_state_dict(sds["vae"])
    qf = ob.Blip2QFormerModel(_qcfg(qcfg))
    qf.load_state_dict(sds["qformer"])
    opipe = OracleBlipPipeline(ou, oc, ov, ob.ContextCLIPTextModel(_text_model(tcfg, sds["text"])), qf)
    src = synthetic_source(2, 512, 512)[None]
    subj_img = synthetic_source(3, 512, 512)[None]  # the subject image is ANOTHER image (run_aug.py:445-451)
    ctrl = np.repeat(clib.canny(src, 120, 200)[..., None], 3, axis=3)
    ids = synthetic_token_ids(4, vocab=tcfg.vocab_size)[:, : 77 - qcfg.num_query_tokens]
    nids = synthetic_token_ids(5, vocab=tcfg.vocab_size)
    subj = torch.tensor([[101, 2000, 102]])
    t0 = time.perf_counter()
    ref = opipe(ids, nids, subj, subj_img, ctrl, generator=torch.Generator().manual_seed(1), num_inference_steps=20, guidance_scale=7.5)
    _log(f"config3: fp32 CPU oracle {time.perf_counter() - t0:.1f} s, {len(ref[1])} executed steps")
    bf = opipe.to(DEV, torch.bfloat16)(ids, nids, subj, subj_img, ctrl, generator=torch.Generator().manual_seed(1), num_inference_steps=20, guidance_scale=7.5)
    fp16 = _fp16_run(opipe, lambda p: p(ids, nids, subj, subj_img, ctrl, generator=torch.Generator().manual_seed(1), num_inference_steps=20, guidance_scale=7.5))
    del opipe, ou, oc, ov, qf
    pipe = SaspaBlipControlNetPipeline.from_state_dicts(sds["unet"], sds["controlnet"], sds["vae"], sds["text"], sds["qformer"])
    out = pipe(prompt_ids=ids, neg_ids=nids, subject_ids=subj, reference_image=subj_img, condtioning_image=ctrl, height=512, width=512,
               num_inference_steps=20, guidance_scale=7.5, generator=torch.Generator().manual_seed(1), return_latents_per_step=True, output_type="np")
    q_err, q_bf = (out.query_embeds.float().cpu() - ref[2]).abs().max().item(), (bf[2] - ref[2]).abs().max().item()
    _log(f"config3: subject embeddings max-abs err ours {q_err:.4g} | torch-bf16 {q_bf:.4g} | max|q| {ref[2].abs().max().item():.4g}")
    assert q_err <= max(2 * q_bf, 1e-2 * ref[2].abs().max().item())
    _judge("config3 blip-diffusion pndm 512^2", out.latents_per_step, out.images, ref[:2], bf[:2], fp16[:2] if fp16 else None)


def test_config4_sdxl_turbo_1024_full_size(cuda_device):
    """BASELINE config 4 (one image of the batch): SD-XL UNet (2.57 B) + ControlNet-canny-sdxl (1.25 B) + VAE, two text encoders,
    1024x1024 (latent 128x128), 4 trailing DDIM steps, guidance 0 (no CFG), conditioning scale 0.75."""
    from tests.test_sdxl_cpu import text_models

    torch.set_num_threads(os.cpu_count() or 8)
    sds = random_state_dicts("sdxl", 1234)
    ucfg, vcfg, tcfg, t2cfg = sdxl_configs("sdxl")
    ou, oc, ov = om.UNet2DConditionModel(_ocfg(ucfg)), om.ControlNetModel(_ocfg(ucfg)), om.AutoencoderKL(_vcfg(vcfg))
    ou.load_state_dict(sds["unet"]); oc.load_state_dict(sds["controlnet"]); ov.load_state_dict(sds["vae"])
    t1, t2 = text_models(tcfg, t2cfg, sds["text"], sds["text2"])
    opipe = OracleSDXLPipeline(ou, oc, ov, t1, t2, "ddim_sdxl_turbo")
    src = synthetic_source(6, 1024, 1024)[None]
    ctrl = np.repeat(clib.canny(src, 120, 200)[..., None], 3, axis=3)
    ids = (synthetic_token_ids(7, vocab=tcfg.vocab_size), synthetic_token_ids(8, vocab=t2cfg.vocab_size))
    kw = dict(num_inference_steps=4, guidance_scale=0.0, controlnet_conditioning_scale=0.75)
    t0 = time.perf_counter()
    ref = opipe(ids, None, ctrl, None, generator=torch.Generator().manual_seed(1), **kw)
    _log(f"config4: fp32 CPU oracle {time.perf_counter() - t0:.1f} s, {len(ref[1])} executed steps")
    bf = opipe.to(DEV, torch.bfloat16)(ids, None, ctrl, None, generator=torch.Generator().manual_seed(1), **kw)
    fp16 = _fp16_run(opipe, lambda p: p(ids, None, ctrl, None, generator=torch.Generator().manual_seed(1), **kw))
    del opipe, ou, oc, ov, t1, t2
    pipe = SaspaSDXLControlNetPipeline.from_state_dicts(sds["unet"], sds["controlnet"], sds["vae"], sds["text"], sds["text2"], sampler="ddim_sdxl_turbo",
                                                        img2img=False)
    out = pipe(image=ctrl, prompt_ids=ids, generator=torch.Generator().manual_seed(1), return_latents_per_step=True, output_type="np", **kw)
    _judge("config4 sdxl-turbo ddim-trailing 1024^2", out.latents_per_step, out.images, ref, bf, fp16)
