"""CPU fp32 restatement of the control flow of diffusers 0.32.2
``StableDiffusionControlNetPipeline.__call__`` / ``StableDiffusionControlNetImg2ImgPipeline.__call__``
(pipelines/controlnet/pipeline_controlnet{,_img2img}.py) as the reference invokes them
(run_aug/run_aug.py:233-279): encode prompt (cond + negative, CFG concat [neg, pos]), control image /255,
timesteps (+ img2img get_timesteps), latents (randn_tensor from the CPU generator; img2img draws the VAE
posterior noise first), per step ControlNet -> UNet -> CFG -> scheduler.step, then VAE decode and
postprocess ((x/2+.5).clamp(0,1) -> round(x*255) u8).  safety_checker=None.

TEST INFRASTRUCTURE; **parity unpinned** (see models.py header).
"""
from __future__ import annotations

from typing import List, Optional

import numpy as np
import torch

from . import models as om
from . import schedulers as osched


def make_scheduler(name: str):
    name = name.lower()
    if name == "ddim":
        return osched.DDIMScheduler()
    if name == "ddim_sdxl_turbo":
        return osched.DDIMScheduler(timestep_spacing="trailing", set_alpha_to_one=True, clip_sample=True)
    if name == "unipc_sdxl_turbo":
        return osched.UniPCMultistepScheduler(timestep_spacing="trailing")
    if name in ("unipc", "unipcmultistep"):
        return osched.UniPCMultistepScheduler()
    if name in ("pndm", "plms"):
        return osched.PNDMScheduler()
    raise ValueError(name)


class OraclePipeline:
    def __init__(self, unet: om.UNet2DConditionModel, controlnet: Optional[om.ControlNetModel], vae: om.AutoencoderKL, text_encoder, sampler: str = "ddim"):
        self.unet, self.controlnet, self.vae, self.text_encoder = unet.eval(), controlnet.eval() if controlnet else None, vae.eval(), text_encoder.eval()
        self.sampler = sampler
        self.device, self.dtype = torch.device("cpu"), torch.float32

    def to(self, device, dtype):
        """Stock-torch reduced-precision run of the SAME graph (calibrates the tolerance): modules cast to
        ``dtype`` on ``device``; latents / scheduler state stay fp32 on the CPU, exactly as in the fp32 run."""
        import copy

        o = copy.copy(self)
        for name in ("unet", "controlnet", "vae", "text_encoder", "text_encoder_2"):
            m = getattr(self, name, None)
            if m is not None:
                setattr(o, name, copy.deepcopy(m).to(device, dtype))
        o.device, o.dtype = torch.device(device), dtype
        return o

    def _m(self, x):  # model-side tensor
        return x.to(self.device, self.dtype) if x.is_floating_point() else x.to(self.device)

    def _encode(self, ids):
        """-> (encoder_hidden_states, pooled or None).  SD v1.5: CLIPTextModel last_hidden_state."""
        return self.text_encoder(self._m(ids))[0], None

    def _added(self, pooled, H, W):
        return None

    @torch.no_grad()
    def __call__(self, prompt_ids: torch.Tensor, negative_prompt_ids: Optional[torch.Tensor], control_u8: Optional[np.ndarray],
                 source_u8: Optional[np.ndarray] = None, *, generator: torch.Generator, num_inference_steps: int, guidance_scale: float,
                 strength: float = 1.0, controlnet_conditioning_scale: float = 1.0):
        sched = make_scheduler(self.sampler)
        do_cfg = guidance_scale > 1.0
        pos, pooled = self._encode(prompt_ids)
        B = pos.shape[0]
        if do_cfg:
            neg, npooled = self._encode(negative_prompt_ids)
            text = torch.cat([neg, pos])
            pooled = torch.cat([npooled, pooled]) if pooled is not None else None
        else:
            text = pos
        cond = None
        if control_u8 is not None:
            cond = self._m(torch.from_numpy(control_u8.astype(np.float32) / 255.0).permute(0, 3, 1, 2))
            if do_cfg:
                cond = torch.cat([cond] * 2)
            H, W = control_u8.shape[1:3]
        else:
            H, W = source_u8.shape[1:3]
        added = self._added(pooled, H, W)
        sched.set_timesteps(num_inference_steps)
        timesteps = sched.timesteps
        shape = (B, self.vae.cfg.latent_channels, H // 8, W // 8)
        sf = self.vae.cfg.scaling_factor
        if source_u8 is not None:
            init = min(int(num_inference_steps * strength), num_inference_steps)
            t_start = max(num_inference_steps - init, 0)
            timesteps = timesteps[t_start * sched.order :]
            if hasattr(sched, "set_begin_index"):
                sched.set_begin_index(t_start * sched.order)
            img = torch.from_numpy(source_u8.astype(np.float32) / 255.0).permute(0, 3, 1, 2) * 2.0 - 1.0
            mean, logvar = (m.float().cpu() for m in self.vae.encode_moments(self._m(img)))
            z0 = (mean + torch.exp(0.5 * logvar) * torch.randn(mean.shape, generator=generator, dtype=torch.float32)) * sf
            noise = torch.randn(shape, generator=generator, dtype=torch.float32)
            latents = sched.add_noise(z0, noise, timesteps[0])
        else:
            latents = torch.randn(shape, generator=generator, dtype=torch.float32) * sched.init_noise_sigma
        per_step: List[torch.Tensor] = []
        for t in timesteps:
            x2 = torch.cat([latents] * 2) if do_cfg else latents
            x2 = self._m(sched.scale_model_input(x2, t))
            down, mid = (None, None)
            if self.controlnet is not None:
                down, mid = self.controlnet(x2, t, text, cond, controlnet_conditioning_scale, added_cond_kwargs=added)
            eps = self.unet(x2, t, text, down, mid, added_cond_kwargs=added).float().cpu()
            if do_cfg:
                eu, ec = eps.chunk(2)
                eps = eu + guidance_scale * (ec - eu)
            latents = sched.step(eps, t, latents)
            per_step.append(latents.clone())
        image = self.vae.decode(self._m(latents / sf)).float().cpu()
        image = (image / 2 + 0.5).clamp(0, 1).permute(0, 2, 3, 1).numpy()
        return (image * 255).round().astype("uint8"), per_step


class OracleSDXLPipeline(OraclePipeline):
    """diffusers 0.32.2 ``StableDiffusionXLControlNet{,Img2Img}Pipeline`` (pipelines/controlnet/pipeline_controlnet_sd_xl{,_img2img}.py) as
    the reference builds them for "sd_xl-turbo" (run_aug/run_aug.py:188-199): two text encoders (CLIPTextModel + CLIPTextModelWithProjection),
    ``encode_prompt`` = concat of both ``hidden_states[-2]`` on the channel axis, pooled = ``text_embeds`` of the second encoder;
    ``add_time_ids`` = (original_size, crops_coords_top_left=(0,0), target_size) = [H, W, 0, 0, H, W] (requires_aesthetics_score=False);
    the same added_cond_kwargs go to the ControlNet and the UNet.  guidance_scale 0 (turbo) disables CFG."""

    def __init__(self, unet, controlnet, vae, text_encoder, text_encoder_2, sampler: str = "ddim_sdxl_turbo"):
        super().__init__(unet, controlnet, vae, text_encoder, sampler)
        self.text_encoder_2 = text_encoder_2.eval()

    def _encode(self, ids):
        ids1, ids2 = ids if isinstance(ids, (tuple, list)) else (ids, ids)  # tokenizer / tokenizer_2 outputs
        o1 = self.text_encoder(self._m(ids1), output_hidden_states=True)
        o2 = self.text_encoder_2(self._m(ids2), output_hidden_states=True)
        return torch.cat([o1.hidden_states[-2], o2.hidden_states[-2]], dim=-1), o2[0]

    def _added(self, pooled, H, W):
        tid = torch.tensor([[H, W, 0, 0, H, W]], dtype=torch.float32).repeat(pooled.shape[0], 1)
        return {"text_embeds": pooled, "time_ids": self._m(tid)}


class OracleBlipPipeline(OraclePipeline):
    """diffusers 0.32.2 ``BlipDiffusionControlNetPipeline.__call__`` (pipelines/controlnet/pipeline_controlnet_blip_diffusion.py) as the
    reference calls it (run_aug/run_aug.py:243-250,268-271).  Token ids replace strings (no vocabularies offline):
      prompt_ids   [B, 77-16]  CLIP ids of _build_prompt(prompt, target_subject) truncated/padded to max_len - num_query_tokens
      neg_ids      [B, 77]     CLIP ids of neg_prompt
      subject_ids  [B, L]      BERT ids of source_subject_category (Q-Former text input)
    query_embeds = qformer(preprocess(reference_image), subject) ; text = ctx_clip(prompt_ids, ctx=query_embeds, ctx_begin_pos=2);
    CFG concat [uncond, text]; PNDM (skip_prk_steps); ControlNet with conditioning_scale 1.0 (the pipeline passes none)."""

    def __init__(self, unet, controlnet, vae, ctx_text_encoder, qformer, ctx_begin_pos: int = 2):
        super().__init__(unet, controlnet, vae, ctx_text_encoder, "pndm")
        self.qformer = qformer.eval()
        self.ctx_begin_pos = ctx_begin_pos

    def to(self, device, dtype):
        import copy

        o = super().to(device, dtype)
        o.qformer = copy.deepcopy(self.qformer).to(device, dtype)
        return o

    @torch.no_grad()
    def __call__(self, prompt_ids, neg_ids, subject_ids, reference_u8: np.ndarray, control_u8: np.ndarray, *, generator, num_inference_steps: int,
                 guidance_scale: float = 7.5):
        from .blip import blip_preprocess_reference

        ref = self._m(blip_preprocess_reference(reference_u8, self.qformer.cfg.image_size))
        query = self.qformer(ref, self._m(subject_ids))
        B = prompt_ids.shape[0]
        text = self.text_encoder(self._m(prompt_ids), query, [self.ctx_begin_pos] * B)
        do_cfg = guidance_scale > 1.0
        if do_cfg:
            text = torch.cat([self.text_encoder(self._m(neg_ids)), text])
        H, W = control_u8.shape[1:3]
        sched = make_scheduler("pndm")
        latents = torch.randn((B, self.vae.cfg.latent_channels, H // 8, W // 8), generator=generator, dtype=torch.float32) * sched.init_noise_sigma
        sched.set_timesteps(num_inference_steps)
        cond = self._m(torch.from_numpy(control_u8.astype(np.float32) / 255.0).permute(0, 3, 1, 2))
        if do_cfg:
            cond = torch.cat([cond] * 2)
        per_step: List[torch.Tensor] = []
        for t in sched.timesteps:
            x2 = self._m(torch.cat([latents] * 2) if do_cfg else latents)
            down, mid = self.controlnet(x2, t, text, cond, 1.0)
            eps = self.unet(x2, t, text, down, mid).float().cpu()
            if do_cfg:
                eu, ec = eps.chunk(2)
                eps = eu + guidance_scale * (ec - eu)
            latents = sched.step(eps, t, latents)
            per_step.append(latents.clone())
        image = self.vae.decode(self._m(latents / self.vae.cfg.scaling_factor)).float().cpu()
        image = (image / 2 + 0.5).clamp(0, 1).permute(0, 2, 3, 1).numpy()
        return (image * 255).round().astype("uint8"), per_step, query.float().cpu(), text.float().cpu()
