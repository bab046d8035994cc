"""Launches the tcgen05 GEMM / implicit-conv kernel on the dominant SD v1.5 shapes (for ncu captures).
Order per round: conv3x3 32x64x64 320->320, GEGLU GEMM 131072x2560x320, in-place residual GEMM 131072x320x320."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))  # repo root (python tools/<name>.py)
import torch

from saspa_aug_b200 import ops
from kernel_bench import rnd

B = 32
x, wk = rnd(B, 64, 64, 320), rnd(320, 9 * 320) * 0.02
co = torch.empty(B, 64, 64, 320, dtype=torch.bfloat16, device="cuda")
a, wg = rnd(B * 4096, 320), rnd(2560, 320) * 0.05
go = torch.empty(B * 4096, 1280, dtype=torch.bfloat16, device="cuda")
wr = rnd(320, 320) * 0.05
h = rnd(B * 4096, 320)
bias_g, bias_r = torch.zeros(2560, device="cuda"), torch.zeros(320, device="cuda")
for _ in range(3):
    ops.conv2d_igemm(x, wk, 3, out=co)
    ops.gemm(a, wg, out=go, bias=bias_g, act=ops.ACT_GEGLU)
    ops.gemm(a, wr, out=h, bias=bias_r, residual=h, beta=1.0)
torch.cuda.synchronize()
