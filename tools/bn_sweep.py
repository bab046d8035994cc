"""MMA efficiency by N-tile width: a compute-bound GEMM (148 M-tiles, K = 4096, N = 2560) with BN forced to 64/128/160/256."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))  # repo root (python tools/<name>.py)
import ctypes
import torch
from saspa_aug_b200 import _lib, ops
from kernel_bench import rnd, timeit

lib = ctypes.CDLL(_lib.SO_PATH)
M, N, K = 148 * 128, 2560, 4096
a, b = rnd(M, K), rnd(N, K)
out = torch.empty(M, N, dtype=torch.bfloat16, device="cuda")
for ctas in (1, 2):
  lib.saspa_gemm_force_ctas(ctas)
  for bn in (64, 128, 160, 256):
    lib.saspa_gemm_force_bn(bn)
    ms = timeit(lambda: ops.gemm(a, b, out=out))
    print(f"CTAS={ctas} BN={bn}: {ms:.3f} ms {2.0*M*N*K/ms/1e9:.0f} TF/s")
lib.saspa_gemm_force_bn(0)
x, wk = rnd(32, 64, 64, 320), rnd(320, 9 * 320)
co = torch.empty(32, 64, 64, 320, dtype=torch.bfloat16, device="cuda")
for ctas in (1, 2):
  lib.saspa_gemm_force_ctas(ctas)
  for bn in (128, 160, 256):
    lib.saspa_gemm_force_bn(bn)
    ms = timeit(lambda: ops.conv2d_igemm(x, wk, 3, out=co))
    print(f"conv 32x64x64 320->320 CTAS={ctas} BN={bn}: {ms:.3f} ms {2.0*32*4096*320*2880/ms/1e9:.0f} TF/s")
lib.saspa_gemm_force_ctas(0)
lib.saspa_gemm_force_bn(0)
