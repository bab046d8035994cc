"""Images/s of the other BASELINE configs (not bench.py lines: bench.py measures config 2): config 3 = BLIP-Diffusion + ControlNet, batch
16, 512x512 (PLMS, CFG 7.5; subject embedding computed once per batch); config 4 = SD-XL-turbo + ControlNet, 1024x1024, 4 steps, batch 8
(no CFG).  Random-init weights of the real architectures, synthetic inputs resident on the device, CUDA events, 1 warm-up + 2 timed runs.
Usage (GPU box): python tools/config_timing.py [blip] [sdxl]"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))  # repo root (python tools/<name>.py)
import numpy as np
import torch

from saspa_aug_b200 import ops
from saspa_aug_b200.synthetic import synthetic_source, synthetic_token_ids


def timed(fn, n_img, tag, runs=2):
    fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(runs):
        fn()
    e.record()
    torch.cuda.synchronize()
    ms = s.elapsed_time(e) / runs
    print(f"{tag}: {ms:.1f} ms per batch of {n_img} -> {n_img / ms * 1e3:.2f} images/s", flush=True)


def blip(steps=20, batch=16):
    from saspa_aug_b200.pipelines import SaspaBlipControlNetPipeline

    t0 = time.time()
    pipe = SaspaBlipControlNetPipeline.random_init("blip", seed=1234)
    print(f"BLIP-Diffusion pipeline built in {time.time() - t0:.0f} s", flush=True)
    dev = pipe.device
    src = torch.from_numpy(np.stack([synthetic_source(s) for s in range(batch)])).to(dev)
    _, ctrl = ops.canny(src, 120, 200, want_ctrl=True)
    ids = torch.cat([synthetic_token_ids(i) for i in range(batch)])[:, :61].to(dev)
    neg = pipe.encode_prompt_ids(synthetic_token_ids(999_999).to(dev)).expand(batch, -1, -1).contiguous()
    subj = [torch.tensor([101, 2000, 102])] * batch
    noise = torch.randn((batch, 4, 64, 64), generator=torch.Generator().manual_seed(1)).to(dev)

    def run():
        q = pipe.get_query_embeddings(src, subj)
        text = pipe.encode_subject_prompt(ids, q)
        return pipe.generate_batch(text, neg, None, None, noise=noise, num_inference_steps=steps, guidance_scale=7.5, controlnet_conditioning_scale=1.0,
                                   control_bf16=ctrl)

    timed(run, batch, f"config 3  BLIP-Diffusion + ControlNet-canny, batch {batch}, 512x512, {steps} PLMS steps (+1 repeated), CFG 7.5, Q-Former per batch")


def sdxl(steps=4, batch=8, res=1024):
    from saspa_aug_b200.pipelines import SaspaSDXLControlNetPipeline

    t0 = time.time()
    pipe = SaspaSDXLControlNetPipeline.random_init("sdxl", seed=1234, img2img=False)
    print(f"SD-XL pipeline built in {time.time() - t0:.0f} s", flush=True)
    dev = pipe.device
    pipe.vae_micro_batch = 2
    src = torch.from_numpy(np.stack([synthetic_source(s, res, res) for s in range(batch)])).to(dev)
    _, ctrl = ops.canny(src, 120, 200, want_ctrl=True)
    ids = torch.cat([synthetic_token_ids(i) for i in range(batch)]).to(dev)
    noise = torch.randn((batch, 4, res // 8, res // 8), generator=torch.Generator().manual_seed(1)).to(dev)

    def run():
        text, pooled = pipe.encode_prompt_ids(ids)
        added = {"text_embeds": pooled.contiguous(), "time_ids": pipe.time_ids(batch, res, res)}
        return pipe.generate_batch(text, None, None, None, noise=noise, num_inference_steps=steps, guidance_scale=0.0, controlnet_conditioning_scale=0.75,
                                   control_bf16=ctrl, added=added)

    timed(run, batch, f"config 4  SD-XL-turbo + ControlNet-canny, batch {batch}, {res}x{res}, {steps} DDIM-trailing steps, no CFG")


if __name__ == "__main__":
    which = sys.argv[1:] or ["blip", "sdxl"]
    if "blip" in which:
        blip()
        torch.cuda.empty_cache()
    if "sdxl" in which:
        sdxl()
