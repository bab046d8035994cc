"""A/B timing of the two attention kernels (mma.sync flash vs tcgen05/TMEM) on the SD v1.5 shapes of one UNet step at
micro-batch 16 (32 CFG rows).  CUDA events, L2 flushed between timed launches.  A tuning aid, not the bench.py contract."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))  # repo root (python tools/<name>.py)
import torch

from saspa_aug_b200 import _lib, ops
from kernel_bench import rnd, timeit


def main():
    lib = _lib.load()
    B = 32
    for b, heads, tq, tkv, d in [(B, 8, 4096, 4096, 40), (B, 8, 1024, 1024, 80), (B, 8, 256, 256, 160), (B, 8, 64, 64, 160), (B, 8, 4096, 77, 40),
                                 (B, 8, 1024, 77, 80), (B, 8, 256, 77, 160), (B, 12, 77, 77, 64), (16, 16, 4096, 4096, 64), (8, 16, 4096, 4096, 128)]:
        qkv = rnd(b, tq, 3 * heads * d)
        c = heads * d
        q = qkv[..., :c]
        if tkv == tq:
            k, v = qkv[..., c:2 * c], qkv[..., 2 * c:]
        else:
            kv = rnd(b, tkv, 2 * c)
            k, v = kv[..., :c], kv[..., c:]
        out = torch.empty(b, tq, c, dtype=torch.bfloat16, device="cuda")
        res = []
        for impl in (1, 2, 0):
            lib.saspa_attention_impl(impl)
            ms = timeit(lambda: ops.attention(q, k, v, heads, out=out))
            res.append((ms, 4.0 * b * heads * tq * tkv * d / ms / 1e9))
        lib.saspa_attention_impl(0)
        print(f"attn b{b} h{heads} {tq}x{tkv} d{d}: mma.sync {res[0][0]:.3f} ms | tcgen05 {res[1][0]:.3f} ms {res[1][1]:.0f} TF/s | auto {res[2][0]:.3f} ms {res[2][1]:.0f} TF/s")


if __name__ == "__main__":
    main()
