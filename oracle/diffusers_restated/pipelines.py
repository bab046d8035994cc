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
    if name in ("unipc", "unipcmultistep"):
        return osched.UniPCMultistepScheduler()
    if name in ("pndm", "plms"):
        return osched.PNDMScheduler()
    raise ValueError(name)


class OraclePipeline:
    def __init__(self, unet: om.UNet2DConditionModel, controlnet: Optional[om.ControlNetModel], vae: om.AutoencoderKL, text_encoder, sampler: str = "ddim"):
        self.unet, self.controlnet, self.vae, self.text_encoder = unet.eval(), controlnet.eval() if controlnet else None, vae.eval(), text_encoder.eval()
        self.sampler = sampler

    @torch.no_grad()
    def __call__(self, prompt_ids: torch.Tensor, negative_prompt_ids: Optional[torch.Tensor], control_u8: Optional[np.ndarray],
                 source_u8: Optional[np.ndarray] = None, *, generator: torch.Generator, num_inference_steps: int, guidance_scale: float,
                 strength: float = 1.0, controlnet_conditioning_scale: float = 1.0):
        sched = make_scheduler(self.sampler)
        do_cfg = guidance_scale > 1.0
        pos = self.text_encoder(prompt_ids)[0]
        B = pos.shape[0]
        if do_cfg:
            neg = self.text_encoder(negative_prompt_ids)[0]
            text = torch.cat([neg, pos])
        else:
            text = pos
        cond = None
        if control_u8 is not None:
            cond = torch.from_numpy(control_u8.astype(np.float32) / 255.0).permute(0, 3, 1, 2)
            if do_cfg:
                cond = torch.cat([cond] * 2)
            H, W = control_u8.shape[1:3]
        else:
            H, W = source_u8.shape[1:3]
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
            mean, logvar = self.vae.encode_moments(img)
            z0 = (mean + torch.exp(0.5 * logvar) * torch.randn(mean.shape, generator=generator, dtype=torch.float32)) * sf
            noise = torch.randn(shape, generator=generator, dtype=torch.float32)
            latents = sched.add_noise(z0, noise, timesteps[0])
        else:
            latents = torch.randn(shape, generator=generator, dtype=torch.float32) * sched.init_noise_sigma
        per_step: List[torch.Tensor] = []
        for t in timesteps:
            x2 = torch.cat([latents] * 2) if do_cfg else latents
            x2 = sched.scale_model_input(x2, t)
            down, mid = (None, None)
            if self.controlnet is not None:
                down, mid = self.controlnet(x2, t, text, cond, controlnet_conditioning_scale)
            eps = self.unet(x2, t, text, down, mid)
            if do_cfg:
                eu, ec = eps.chunk(2)
                eps = eu + guidance_scale * (ec - eu)
            latents = sched.step(eps, t, latents)
            per_step.append(latents.clone())
        image = self.vae.decode(latents / sf)
        image = (image / 2 + 0.5).clamp(0, 1).permute(0, 2, 3, 1).numpy()
        return (image * 255).round().astype("uint8"), per_step
