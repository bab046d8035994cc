"""The small-channel layers at benchmark size (32 images, 512^2 control maps; 64 latent rows for conv_in): direct kernel
(saspa_conv3x3_small_bf16) against the path it replaced (im2col + tcgen05 GEMM for Cin 3 / 4, 64-channel-granular implicit GEMM for
Cin 16 / 32).  CUDA events, L2 flushed.  Usage: python tools/conv_small_bench.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from kernel_bench import rnd, timeit
from saspa_aug_b200 import ops
from saspa_aug_b200.layout import conv_weight_kmajor


def main():
    layers = [(32, 512, 512, 3, 16, 1), (32, 512, 512, 16, 16, 1), (32, 512, 512, 16, 32, 2), (32, 256, 256, 32, 32, 1), (32, 256, 256, 32, 96, 2),
              (64, 64, 64, 4, 320, 1), (16, 512, 512, 3, 64, 1), (16, 512, 512, 3, 128, 1)]
    tot_new = tot_old = 0.0
    for n, h, w, cin, cout, stride in layers:
        x = rnd(n, h, w, cin)
        wt = torch.randn(cout, cin, 3, 3) / (9 * cin) ** 0.5
        kpad = (9 * cin + 7) // 8 * 8
        wk = conv_weight_kmajor(wt, kpad).to(torch.bfloat16).cuda()
        bias = torch.zeros(cout, device="cuda")
        oh, ow = (h - 1) // stride + 1, (w - 1) // stride + 1
        out = torch.empty(n, oh, ow, cout, dtype=torch.bfloat16, device="cuda")
        t_new = timeit(lambda: ops.conv3x3_small(x, wk, bias, ops.ACT_SILU, stride, out=out), iters=6, warmup=2)
        if cin % 8 == 0:
            old = lambda: ops.conv2d_igemm(x, wk, 3, out=out, bias=bias, act=ops.ACT_SILU, stride=stride, pad=1, out_hw=(oh, ow))
        else:
            old = lambda: ops.gemm(ops.im2col(x, 3, 3, stride, 1, 1, oh, ow, kpad), wk, out=out.view(-1, cout), bias=bias, act=ops.ACT_SILU)
        t_old = timeit(old, iters=6, warmup=2)
        mb = (n * h * w * cin + n * oh * ow * cout) * 2 / 1e6
        print(f"conv {cin:3d}->{cout:3d} s{stride} n={n} {h}x{w}: direct {t_new * 1e3:8.1f} us ({mb / t_new:6.0f} GB/s algorithmic) | previous path {t_old * 1e3:8.1f} us")
        if n == 32 or cin == 4:
            tot_new += t_new
            tot_old += t_old
    print(f"conditioning embedding (first five layers) + one conv_in: {tot_new:.2f} ms direct vs {tot_old:.2f} ms before")


if __name__ == "__main__":
    main()
