"""Tile-shape sweep over the transformer-block GEMM shapes of one UNet+ControlNet step (micro-batch 32 => 64 rows batches): each shape
with the automatic tile choice and with BN / CTA-pair choices forced through the tuning hooks, in the epilogue flavour the step uses
(LayerNorm folded for the q / qkv projections, bias + residual for the output projections).  CUDA events, L2 flushed.
Usage: python tools/gemm_shape_sweep.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from kernel_bench import rnd, timeit
from saspa_aug_b200 import _lib, ops


def main():
    lib = _lib.load()
    shapes = [  # (rows, N, K, flavour)
        (262144, 960, 320, "ln"), (262144, 320, 320, "ln"), (65536, 1920, 640, "ln"), (65536, 640, 640, "ln"), (16384, 3840, 1280, "ln"),
        (16384, 1280, 1280, "ln"), (65536, 640, 640, "res"), (65536, 640, 2560, "res"), (16384, 1280, 1280, "res"), (16384, 1280, 5120, "res"),
        (4096, 1280, 1280, "res"), (4096, 3840, 1280, "ln"),
    ]
    for M, N, K, flavour in shapes:
        x, w = rnd(M, K), rnd(N, K) * (1.0 / K ** 0.5)
        bias = torch.zeros(N, device="cuda")
        out = torch.empty(M, N, dtype=torch.bfloat16, device="cuda")
        if flavour == "ln":
            _, st = ops.gemm(rnd(M, K), rnd(K, K) * 0.05, row_stats=True)  # statistics in the producer's slot layout for a K-wide row
            cs = w.float().sum(1).contiguous()
            fn = lambda: ops.gemm(x, w, out=out, bias=bias, ln_stats=st, ln_colsum=cs, ln_eps=1e-5)
        else:
            h = rnd(M, N)
            fn = lambda: ops.gemm(x, w, out=out, bias=bias, residual=h, beta=1.0)
        res = []
        for ctas, bn in [(0, 0), (1, 128), (1, 160), (1, 256), (2, 160), (2, 256)]:
            lib.saspa_gemm_force_ctas(ctas)
            lib.saspa_gemm_force_bn(bn)
            try:
                ms = timeit(fn, iters=6, warmup=2)
                res.append(f"{'auto' if not ctas else f'{ctas}x{bn}'} {ms * 1e3:7.1f} us {2.0 * M * N * K / ms / 1e9:5.0f} TF/s")
            except Exception as e:  # noqa: BLE001
                res.append(f"{ctas}x{bn} n/a")
            finally:
                lib.saspa_gemm_force_ctas(0)
                lib.saspa_gemm_force_bn(0)
        print(f"({M}, {N}, {K}) {flavour:3s}: " + " | ".join(res), flush=True)


if __name__ == "__main__":
    main()
