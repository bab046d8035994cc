"""UNet+ControlNet CFG step time at a given micro-batch (CUDA events over N steps of generate_batch, no decode), and the tensor-pipe
fraction it implies (2135 GFLOP per image-step).  Env SASPA_FOLD_LN=0 times the unfolded LayerNorm path.
Usage: python tools/step_time.py [--mb 32] [--steps 6]"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from saspa_aug_b200 import ops
from saspa_aug_b200.pipelines import SaspaControlNetPipeline
from saspa_aug_b200.synthetic import synthetic_source, synthetic_token_ids


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--mb", type=int, default=32)
    ap.add_argument("--steps", type=int, default=6)
    ap.add_argument("--graph", action="store_true", help="replay the whole generation from a captured CUDA graph instead of launching it")
    ap.add_argument("--attn", type=int, default=0, help="saspa_attention_impl value (3 = older cross-attention kernel)")
    a = ap.parse_args()
    from saspa_aug_b200 import _lib

    _lib.load().saspa_attention_impl(a.attn)
    dev = torch.device("cuda:0")
    pipe = SaspaControlNetPipeline.random_init("sd15", seed=1234, sampler="unipc", device=dev, img2img=False)
    src = torch.from_numpy(np.stack([synthetic_source(s) for s in range(max(1, a.mb // 2))])).to(dev)
    ids = torch.cat([synthetic_token_ids(i) for i in range(a.mb)]).to(dev)
    neg = pipe.encode_prompt_ids(synthetic_token_ids(999_999).to(dev)).expand(a.mb, -1, -1).contiguous()
    noise = torch.randn((a.mb, 4, 64, 64), generator=torch.Generator().manual_seed(1)).to(dev)
    _, ctrl = ops.canny(src, 120, 200, want_ctrl=True)
    c = ctrl.index_select(0, torch.arange(a.mb, device=dev) // 2)
    text = pipe.encode_prompt_ids(ids)

    def run(steps):
        if a.graph:
            return pipe._generate_graphed(dict(text_embeds=text, neg_embeds=neg, control_bf16=c, noise=noise), num_inference_steps=steps, guidance_scale=7.5,
                                          controlnet_conditioning_scale=0.75, decode=False)
        return pipe.generate_batch(text, neg, None, None, noise=noise, num_inference_steps=steps, guidance_scale=7.5, controlnet_conditioning_scale=0.75,
                                   control_bf16=c, decode=False)

    run(2)
    torch.cuda.synchronize()
    peak = 1398.9
    try:
        peak = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["bf16_tflops_sustained"]
    except Exception:
        pass
    best = None
    for rep in range(3):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        run(a.steps)
        e.record()
        torch.cuda.synchronize()
        ms = s.elapsed_time(e) / a.steps  # includes the once-per-image work (text K/V, cond embedding) amortised over the steps
        best = ms if best is None else min(best, ms)
    frac = 2135e9 * a.mb / (best * 1e-3) / (peak * 1e12)
    print(f"fold_ln={os.environ.get('SASPA_FOLD_LN', '1')} attn_impl={a.attn} graph={int(a.graph)} mb={a.mb}: {best:.2f} ms per step ({best / a.mb:.3f} ms per image-step), "
          f"step_tensor_frac {frac:.3f} of {peak:.0f} TF/s")


if __name__ == "__main__":
    main()
