"""A/B timing of the cross-attention kernels on the SD v1.5 / SD-XL shapes of one UNet+ControlNet step at micro-batch 32 (64 CFG rows):
the persistent tcgen05 kernel (xattention_tc.cu) vs the older mma.sync K/V-resident kernel.  CUDA events, L2 flushed between timed
launches; GB/s = algorithmic Q + O bytes (the op is HBM-bound on them).  A tuning aid, not the bench.py contract."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import torch

from saspa_aug_b200 import _lib, ops
from kernel_bench import rnd, timeit


def main():
    lib = _lib.load()
    for b, heads, tq, tkv, d in [(64, 8, 4096, 77, 40), (64, 8, 1024, 77, 80), (32, 8, 4096, 77, 40), (8, 10, 16384, 77, 64), (8, 20, 4096, 77, 64), (64, 8, 256, 77, 160)]:
        c = heads * d
        q = rnd(b, tq, c)
        kv = rnd(b, tkv, 2 * c)
        k, v = kv[..., :c], kv[..., c:]
        out = torch.empty(b, tq, c, dtype=torch.bfloat16, device="cuda")
        res = []
        for impl in (3, 0):
            lib.saspa_attention_impl(impl)
            res.append(timeit(lambda: ops.attention(q, k, v, heads, out=out)))
        lib.saspa_attention_impl(0)
        gb = 2.0 * b * tq * c * 2 / 1e9
        print(f"xattn b{b} h{heads} {tq}x{tkv} d{d}: mma.sync resident {res[0]:.3f} ms {gb / res[0] * 1e3:.0f} GB/s | tcgen05 persistent {res[1]:.3f} ms {gb / res[1] * 1e3:.0f} GB/s")


if __name__ == "__main__":
    main()
