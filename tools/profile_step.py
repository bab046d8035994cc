"""One micro-batch generation bracketed by cudaProfilerStart/Stop, for `ncu --profile-from-start off`.
Usage (GPU box):
  ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
      --log-file gpurun_out/launches.csv python tools/profile_step.py [--mb 16] [--steps 2] [--decode]
Not the bench.py contract: numbers printed under ncu are never bench values; this only produces launch lists."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))  # repo root (python tools/<name>.py)
import argparse

import numpy as np
import torch

from saspa_aug_b200 import ops
from saspa_aug_b200.pipelines import SaspaControlNetPipeline
from saspa_aug_b200.synthetic import synthetic_source, synthetic_token_ids


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--mb", type=int, default=16)
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--decode", action="store_true")
    ap.add_argument("--shapes", action="store_true", help="CUDA-event time every tcgen05 launch and print a per-shape table (no ncu)")
    a = ap.parse_args()
    dev = torch.device("cuda:0")
    pipe = SaspaControlNetPipeline.random_init("sd15", seed=1234, sampler="unipc", device=dev, img2img=False)
    src = torch.from_numpy(np.stack([synthetic_source(s) for s in range(max(1, a.mb // 2))])).to(dev)
    ids = torch.cat([synthetic_token_ids(i) for i in range(a.mb)]).to(dev)
    neg = pipe.encode_prompt_ids(synthetic_token_ids(999_999).to(dev)).expand(a.mb, -1, -1).contiguous()
    noise = torch.randn((a.mb, 4, 64, 64), generator=torch.Generator().manual_seed(1)).to(dev)
    _, ctrl = ops.canny(src, 120, 200, want_ctrl=True)
    c = ctrl.index_select(0, torch.arange(a.mb, device=dev) // 2)
    text = pipe.encode_prompt_ids(ids)

    def run(steps, decode):
        return pipe.generate_batch(text, neg, None, None, noise=noise, num_inference_steps=steps, guidance_scale=7.5,
                                   controlnet_conditioning_scale=0.75, control_bf16=c, decode=decode)

    run(1, a.decode)  # warm-up (module load, smem attribute calls)
    torch.cuda.synchronize()
    if a.shapes:
        ops.PROFILE = []
        run(a.steps, a.decode)
        torch.cuda.synchronize()
        prof, ops.PROFILE = ops.PROFILE, None
        agg = {}
        for kind, flops, s, e, shape in prof:
            d = agg.setdefault((kind, shape), [0, 0.0, 0.0])
            d[0] += 1
            d[1] += s.elapsed_time(e)
            d[2] += flops
        tot = sum(v[1] for v in agg.values())
        print(f"event-timed tcgen05/attention launches: {sum(v[0] for v in agg.values())}, {tot:.2f} ms over {a.steps} steps")
        for (kind, shape), v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            print(f"{v[1]:8.3f} ms {100 * v[1] / tot:5.1f}% {v[0]:4d}x {1e3 * v[1] / v[0]:8.1f} us {v[2] / v[1] / 1e9:7.0f} TF/s  {kind} {shape}")
        return
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.profiler.start()
    s.record()
    run(a.steps, a.decode)
    e.record()
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
    print(f"profiled region: {a.steps} steps mb={a.mb} decode={a.decode}: {s.elapsed_time(e):.2f} ms")


if __name__ == "__main__":
    main()
