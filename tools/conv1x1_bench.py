"""1x1 convolutions of one UNet+ControlNet step at micro-batch 32: the convolution entry point (4-D TMA pixel boxes) against the GEMM entry
point on the same NHWC tensor viewed as [pixels, Cin] (2-D boxes; CTA pairs from K = 1024).  CUDA events, L2 flushed."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from kernel_bench import rnd, timeit
from saspa_aug_b200 import ops


def main():
    shapes = [(64, 64, 64, 320, 320), (64, 64, 64, 640, 320), (64, 64, 64, 960, 320), (64, 32, 32, 640, 640), (64, 32, 32, 320, 640), (64, 32, 32, 1280, 640),
              (64, 32, 32, 1920, 640), (64, 16, 16, 1280, 1280), (64, 16, 16, 640, 1280), (64, 16, 16, 2560, 1280), (64, 8, 8, 1280, 1280), (64, 8, 8, 2560, 1280)]
    ta = tb = 0.0
    for n, h, w, cin, cout in shapes:
        x, wk = rnd(n, h, w, cin), rnd(cout, cin) * (1.0 / cin ** 0.5)
        bias = torch.zeros(cout, device="cuda")
        res = rnd(n, h, w, cout)
        out = torch.empty(n, h, w, cout, dtype=torch.bfloat16, device="cuda")
        t_conv = timeit(lambda: ops.conv2d_igemm(x, wk, 1, out=out, bias=bias, residual=res, beta=1.0), iters=6, warmup=2)
        t_gemm = timeit(lambda: ops.gemm(x.view(-1, cin), wk, out=out.view(-1, cout), bias=bias, residual=res.view(-1, cout), beta=1.0), iters=6, warmup=2)
        fl = 2.0 * n * h * w * cin * cout
        print(f"1x1 conv {n}x{h}x{w} {cin:4d}->{cout:4d} (+ bias + residual): conv entry {t_conv * 1e3:7.1f} us {fl / t_conv / 1e9:5.0f} TF/s | gemm entry {t_gemm * 1e3:7.1f} us {fl / t_gemm / 1e9:5.0f} TF/s")
        ta += t_conv
        tb += t_gemm
    print(f"sum over the shapes: {ta:.3f} ms -> {tb:.3f} ms")


if __name__ == "__main__":
    main()
