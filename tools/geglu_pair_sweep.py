"""GEGLU feed-forward GEMMs (ff.net.0.proj, LayerNorm folded) of the three transformer levels at micro-batch 32: single-CTA tiles against CTA
pairs (tuning hook).  CUDA events, L2 flushed."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from kernel_bench import rnd, timeit
from saspa_aug_b200 import _lib, ops


def main():
    lib = _lib.load()
    for M, C in [(262144, 320), (65536, 640), (16384, 1280), (4096, 1280)]:
        N = 8 * C
        x, w = rnd(M, C), rnd(N, C) * (1.0 / C ** 0.5)
        bias = torch.zeros(N, device="cuda")
        out = torch.empty(M, N // 2, dtype=torch.bfloat16, device="cuda")
        _, st = ops.gemm(rnd(M, C), rnd(C, C) * 0.05, row_stats=True)
        cs = w.float().sum(1).contiguous()
        res = []
        for ctas in (1, 2):
            lib.saspa_gemm_force_ctas(ctas)
            try:
                ms = timeit(lambda: ops.gemm(x, w, out=out, bias=bias, act=ops.ACT_GEGLU, ln_stats=st, ln_colsum=cs, ln_eps=1e-5), iters=6, warmup=2)
                ms2 = timeit(lambda: ops.gemm(x, w, out=out, bias=bias, act=ops.ACT_GEGLU), iters=6, warmup=2)
                res.append(f"ctas={ctas}: LN-folded {ms * 1e3:7.1f} us {2.0 * M * N * C / ms / 1e9:5.0f} TF/s, plain {ms2 * 1e3:7.1f} us {2.0 * M * N * C / ms2 / 1e9:5.0f} TF/s")
            finally:
                lib.saspa_gemm_force_ctas(0)
        print(f"GEGLU ({M}, {N}, {C}): " + " | ".join(res), flush=True)


if __name__ == "__main__":
    main()
