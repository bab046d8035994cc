"""GPU parity of the BLIP-Diffusion + ControlNet path (BASELINE config 3; reference call run_aug.py:243-250,268-271) against the
CPU fp32 oracle (oracle/diffusers_restated/blip.py + OracleBlipPipeline) on identical random-init weights, ids and seeds.
Tolerance: calibrated against stock torch bf16 on the same oracle graph (tests/test_models_gpu.py docstring)."""
import copy

import numpy as np
import pytest
import torch

from oracle.diffusers_restated import blip as ob
from oracle.diffusers_restated import models as om
from oracle.diffusers_restated.pipelines import OracleBlipPipeline
from saspa_aug_b200 import checkpoints as ck
from saspa_aug_b200 import nn as snn
from saspa_aug_b200 import ops
from saspa_aug_b200.pipelines import SaspaBlipControlNetPipeline, blip_configs, random_state_dicts
from saspa_aug_b200.synthetic import synthetic_source, synthetic_token_ids
from tests.test_blip_cpu import _ocfg as _qcfg
from tests.test_models_gpu import _bound, _check_pipeline, _ocfg, _text_model, _vcfg

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _reference_images(n, h, w, seed0=20):
    return np.stack([synthetic_source(seed0 + s, h, w) for s in range(n)])


def test_blip_preprocess_bit_exact_resize_then_normalise(cuda_device):
    """BlipImageProcessor: the PIL bicubic resize is reproduced bit-exactly (u8), then /255 and CLIP mean/std in bf16."""
    cfg = ck.Blip2Config.tiny()
    sd = ck.random_state_dict(ck.qformer_shapes(cfg), 5)
    pipe = SaspaBlipControlNetPipeline.__new__(SaspaBlipControlNetPipeline)
    pipe.qformer = snn.QFormer(sd, cfg, torch.device(DEV))
    pipe.device = torch.device(DEV)
    img = _reference_images(2, 96, 80)
    want = ob.blip_preprocess_reference(img, cfg.image_size).permute(0, 2, 3, 1)
    got = pipe.preprocess_reference(torch.from_numpy(img).to(DEV)).float().cpu()
    assert got.shape == want.shape
    assert (got - want).abs().max().item() <= 2.0 ** -7 * want.abs().max().item() + 1e-3  # one bf16 rounding of the fp32 value


@pytest.mark.parametrize("name,n,L", [("tiny", 3, 3), ("blipdiffusion", 2, 4)])
def test_qformer_matches_oracle(cuda_device, name, n, L):
    """Blip2QFormerModel (ViT vision tower + 12-layer Q-Former + ProjLayer) -> 16 subject embeddings."""
    torch.set_num_threads(max(torch.get_num_threads(), 8))
    cfg = getattr(ck.Blip2Config, name)()
    sd = ck.random_state_dict(ck.qformer_shapes(cfg), 41)
    o = ob.Blip2QFormerModel(_qcfg(cfg)).eval()
    o.load_state_dict(sd)
    x = ob.blip_preprocess_reference(_reference_images(n, 160, 160), cfg.image_size)
    ids = torch.randint(103, cfg.vocab_size, (n, L), generator=torch.Generator().manual_seed(3))
    ids[:, 0], ids[:, -1] = 101, 102
    with torch.no_grad():
        r32 = o(x, ids)
        rbf = copy.deepcopy(o).to(DEV, torch.bfloat16)(x.to(DEV, torch.bfloat16), ids.to(DEV)).float().cpu()
    q = snn.QFormer(sd, cfg, torch.device(DEV))
    mine = q(x.permute(0, 2, 3, 1).contiguous().to(DEV, torch.bfloat16), ids.to(DEV)).float().cpu()
    e_mine, e_bf = _bound(mine, r32, rbf)
    print(f"qformer {name}: max-abs err ours {e_mine:.4g}, torch-bf16 {e_bf:.4g}, max|ref| {r32.abs().max().item():.4g}")


def test_ctx_clip_text_encoder_matches_oracle(cuda_device):
    """ContextCLIPTextModel at full CLIP ViT-L text size: 16 subject embeddings spliced after 2 token embeddings."""
    tcfg = ck.CLIPTextConfig.sd15()
    sd = ck.random_state_dict(ck.clip_text_shapes(tcfg), 31)
    ref = ob.ContextCLIPTextModel(_text_model(tcfg, sd)).eval()
    g = torch.Generator().manual_seed(4)
    ids = synthetic_token_ids(5, batch=3, vocab=tcfg.vocab_size)[:, :61]
    ctx = torch.randn((3, 16, tcfg.hidden_size), generator=g) * 0.5
    with torch.no_grad():
        r32 = ref(ids, ctx, [2, 2, 2])
        rbf = copy.deepcopy(ref).to(DEV, torch.bfloat16)(ids.to(DEV), ctx.to(DEV, torch.bfloat16), [2, 2, 2]).float().cpu()
    te = snn.CLIPTextEncoder(sd, torch.device(DEV), tcfg.num_attention_heads, tcfg.hidden_act, tcfg.layer_norm_eps)
    mine = te(ids.to(DEV), ctx_embeddings=ctx.to(DEV, torch.bfloat16), ctx_begin_pos=2).float().cpu()
    assert mine.shape == (3, 77, tcfg.hidden_size)
    _bound(mine, r32, rbf)


def _tiny_oracle(sds):
    ucfg, vcfg, tcfg, qcfg = blip_configs("tiny_blip")
    ou, oc, ov = om.UNet2DConditionModel(_ocfg(ucfg)), om.ControlNetModel(_ocfg(ucfg)), om.AutoencoderKL(_vcfg(vcfg))
    ou.load_state_dict(sds["unet"]); oc.load_state_dict(sds["controlnet"]); ov.load_state_dict(sds["vae"])
    qf = ob.Blip2QFormerModel(_qcfg(qcfg))
    qf.load_state_dict(sds["qformer"])
    return OracleBlipPipeline(ou, oc, ov, ob.ContextCLIPTextModel(_text_model(tcfg, sds["text"])), qf), (ucfg, vcfg, tcfg, qcfg)


@pytest.mark.parametrize("steps,gs", [(6, 7.5), (4, 1.0)])
def test_blip_pipeline_tiny_matches_oracle(cuda_device, steps, gs):
    """Whole BLIP-Diffusion ControlNet call (tiny same-topology models, 128x128, batch 2 with two different reference images):
    subject embeddings, spliced text embeddings, per-step PLMS latents and the final image vs the fp32 oracle."""
    from oracle import clib

    sds = random_state_dicts("tiny_blip", 300)
    opipe, (ucfg, vcfg, tcfg, qcfg) = _tiny_oracle(sds)
    src = _reference_images(2, 128, 128, 30)
    ctrl = np.repeat(clib.canny(src, 120, 200)[..., None], 3, axis=3)
    ids = synthetic_token_ids(9, batch=2, vocab=tcfg.vocab_size)[:, : 77 - qcfg.num_query_tokens]
    nids = synthetic_token_ids(10, batch=1, vocab=tcfg.vocab_size).expand(2, -1)
    subj = torch.tensor([[101, 150, 102], [101, 333, 102]])
    img_o, lat_o, q_o, t_o = opipe(ids, nids, subj, src, ctrl, generator=torch.Generator().manual_seed(1), num_inference_steps=steps, guidance_scale=gs)
    pipe = SaspaBlipControlNetPipeline.from_state_dicts(sds["unet"], sds["controlnet"], sds["vae"], sds["text"], sds["qformer"], unet_cfg=ucfg,
                                                        vae_cfg=vcfg, text_cfg=tcfg, qformer_cfg=qcfg)
    out = pipe(prompt_ids=ids, neg_ids=nids, subject_ids=subj, reference_image=src, condtioning_image=ctrl, height=128, width=128,
               num_inference_steps=steps, guidance_scale=gs, generator=torch.Generator().manual_seed(1), return_latents_per_step=True, output_type="np")
    img_b, lat_b, q_b, t_b = opipe.to(DEV, torch.bfloat16)(ids, nids, subj, src, ctrl, generator=torch.Generator().manual_seed(1),
                                                           num_inference_steps=steps, guidance_scale=gs)
    _bound(out.query_embeds.float().cpu(), q_o, q_b)
    _bound(out.text_embeds.float().cpu(), t_o, t_b)
    _check_pipeline(out, img_o, lat_o, img_b, lat_b, f"blip pndm gs={gs}")


def test_blip_pipeline_reference_call_signature(cuda_device):
    """The exact kwargs run_aug.py sends (strings + PIL images), through pass_thorugh_pipe; the subject embedding of a repeated
    (reference image, subject) is the same tensor and images are deterministic for a fixed generator seed."""
    from PIL import Image

    from saspa_aug_b200 import run_aug as ra

    pipe = ra.init_pipeline("tiny_blip", "canny", False)
    assert isinstance(pipe, SaspaBlipControlNetPipeline) and type(pipe.scheduler).__name__ == "PNDMScheduler"
    src = Image.fromarray(synthetic_source(40, 128, 128))
    canny = ra.generate_canny(src, 120, 200, 128)
    outs = [ra.pass_thorugh_pipe("blip_diffusion", pipe, "flying at sunset", src, False, 0.5, 4, torch.Generator().manual_seed(7), 7.5, 0.75,
                                 control_image=canny, blip_src_category="airplane", blip_target_category="airplane") for _ in range(2)]
    a, b = (np.asarray(o) for o in outs)
    assert a.shape == (128, 128, 3) and a.dtype == np.uint8 and a.std() > 0 and np.array_equal(a, b)
