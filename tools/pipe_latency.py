"""Latency of the reference-shaped single call -- pass_thorugh_pipe(...) -> pipe(prompt, image=<PIL canny>, ...) -> PIL (run_aug.py:233-279)
-- at batch 1 on the full-size SD v1.5 pipeline, with and without the CUDA-graph replay of the call shape (SASPA_CUDA_GRAPH=0|1).
Usage: python tools/pipe_latency.py [steps]"""
import os
import statistics
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from PIL import Image

from saspa_aug_b200 import ops, run_aug
from saspa_aug_b200.pipelines import SaspaControlNetPipeline
from saspa_aug_b200.synthetic import synthetic_source


def main():
    steps = int(sys.argv[1]) if len(sys.argv) > 1 else 20
    pipe = SaspaControlNetPipeline.random_init("sd15", seed=1234, sampler="unipc", device="cuda:0", img2img=False)
    src = Image.fromarray(synthetic_source(7))
    canny = run_aug.generate_canny(src, 120, 200, 512)
    outs = {}
    for graph in (0, 4):
        pipe.cuda_graph_max_images = graph
        times = []
        for rep in range(5):
            torch.cuda.synchronize()
            l0, t0 = ops.LAUNCHES, time.perf_counter()
            img = run_aug.pass_thorugh_pipe("sd_v1.5", pipe, "an airplane flying over a snowy mountain range at sunset", src, 0, 0.85, steps,
                                            torch.Generator().manual_seed(3), 7.5, 0.75, control_image=canny)
            torch.cuda.synchronize()
            if rep >= 2:  # call 0 builds the negative-prompt cache, call 1 captures the graph
                times.append(time.perf_counter() - t0)
                launches = ops.LAUNCHES - l0
        outs[graph] = img
        print(f"cuda_graph={'on' if graph else 'off'}: pass_thorugh_pipe batch 1, {steps} steps: median {statistics.median(times) * 1e3:.1f} ms "
              f"({launches} host-side kernel launches per call)")
    import numpy as np

    print("images identical with and without the graph:", bool(np.array_equal(np.asarray(outs[0]), np.asarray(outs[4]))))


if __name__ == "__main__":
    main()
