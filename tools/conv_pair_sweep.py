"""3x3 convolutions that run with one TMA box per tap (8x8 maps, stride 2) at micro-batch 32: single-CTA tiles against CTA pairs and N-tile
widths forced through the tuning hooks.  CUDA events, L2 flushed."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from kernel_bench import rnd, timeit
from saspa_aug_b200 import _lib, ops


def main():
    lib = _lib.load()
    shapes = [(64, 8, 8, 1280, 1280, 1), (64, 8, 8, 2560, 1280, 1), (64, 16, 16, 1280, 1280, 2), (64, 32, 32, 640, 640, 2), (64, 64, 64, 320, 320, 2),
              (64, 16, 16, 1280, 1280, 1), (64, 16, 16, 2560, 1280, 1), (64, 32, 32, 640, 640, 1), (64, 32, 32, 1280, 640, 1), (64, 64, 64, 320, 320, 1),
              (64, 64, 64, 640, 320, 1)]
    for n, h, w, cin, cout, stride in shapes:
        x, wk = rnd(n, h, w, cin), rnd(cout, 9 * cin) * (1.0 / (9 * cin) ** 0.5)
        bias = torch.zeros(cout, device="cuda")
        oh, ow = (h - 1) // stride + 1, (w - 1) // stride + 1
        out = torch.empty(n, oh, ow, cout, dtype=torch.bfloat16, device="cuda")
        res = []
        for ctas, bn in [(0, 0), (1, 160), (1, 256), (2, 160), (2, 256)]:
            lib.saspa_gemm_force_ctas(ctas)
            lib.saspa_gemm_force_bn(bn)
            try:
                ms = timeit(lambda: ops.conv2d_igemm(x, wk, 3, out=out, bias=bias, stride=stride, pad=1, out_hw=(oh, ow)), iters=6, warmup=2)
                res.append(f"{'auto' if not ctas else f'{ctas}x{bn}'} {ms * 1e3:7.1f} us {2.0 * n * oh * ow * cout * 9 * cin / ms / 1e9:5.0f} TF/s")
            except Exception as e:  # noqa: BLE001
                res.append(f"{ctas}x{bn} n/a ({str(e)[:40]})")
            finally:
                lib.saspa_gemm_force_ctas(0)
                lib.saspa_gemm_force_bn(0)
        print(f"conv3x3 {n}x{h}x{w} {cin}->{cout} s{stride}: " + " | ".join(res), flush=True)


if __name__ == "__main__":
    main()
