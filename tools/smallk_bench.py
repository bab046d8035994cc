"""Small-K GEMMs of the 64x64 transformer blocks at the benchmark micro-batch (M = 64 * 4096 rows), each epilogue flavour the UNet uses:
plain / bias, + residual, + row statistics out, LayerNorm folded in, GEGLU.  CUDA events, L2 flushed between launches; prints TF/s and
the HBM traffic rate (these shapes sit between the two roofs).  --only N launches just variant N three times (for ncu -k gemm_tc_kernel).
Usage: python tools/smallk_bench.py [--only N] [--rows 262144]"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from kernel_bench import rnd, timeit
from saspa_aug_b200 import ops


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--only", type=int, default=-1)
    ap.add_argument("--rows", type=int, default=262144)
    ap.add_argument("--sweep", action="store_true", help="also time forced tile shapes (CTA pairs, BN)")
    ap.add_argument("--chain", action="store_true", help="time consumers right after a producer (no L2 flush), forward vs reversed row-tile order")
    a = ap.parse_args()
    M = a.rows
    x = rnd(M, 320)
    x1280 = rnd(M, 1280)
    w320, w960, wg, w1280 = rnd(320, 320) * 0.05, rnd(960, 320) * 0.05, rnd(2560, 320) * 0.05, rnd(320, 1280) * 0.03
    b320, b960, bg = torch.zeros(320, device="cuda"), torch.zeros(960, device="cuda"), torch.zeros(2560, device="cuda")
    h = rnd(M, 320)
    o320, o960, o1280 = torch.empty_like(h), torch.empty(M, 960, dtype=torch.bfloat16, device="cuda"), torch.empty(M, 1280, dtype=torch.bfloat16, device="cuda")
    _, st = ops.gemm(x, w320, bias=b320, residual=h, beta=1.0, row_stats=True)
    cs320, cs960, csg = w320.float().sum(1).contiguous(), w960.float().sum(1).contiguous(), wg.float().sum(1).contiguous()
    variants = [
        ("320->320 bias", 320, 320, lambda: ops.gemm(x, w320, out=o320, bias=b320), 2),
        ("320->320 bias + residual", 320, 320, lambda: ops.gemm(x, w320, out=o320, bias=b320, residual=h, beta=1.0), 3),
        ("320->320 bias + residual + stats out", 320, 320, lambda: ops.gemm(x, w320, out=o320, bias=b320, residual=h, beta=1.0, row_stats=True), 3),
        ("320->320 LN folded", 320, 320, lambda: ops.gemm(x, w320, out=o320, bias=b320, ln_stats=st, ln_colsum=cs320, ln_eps=1e-5), 2),
        ("320->960 (QKV) plain", 960, 320, lambda: ops.gemm(x, w960, out=o960), 4),
        ("320->960 (QKV) LN folded", 960, 320, lambda: ops.gemm(x, w960, out=o960, bias=b960, ln_stats=st, ln_colsum=cs960, ln_eps=1e-5), 4),
        ("320->2560 GEGLU", 2560, 320, lambda: ops.gemm(x, wg, out=o1280, bias=bg, act=ops.ACT_GEGLU), 5),
        ("320->2560 GEGLU LN folded", 2560, 320, lambda: ops.gemm(x, wg, out=o1280, bias=bg, act=ops.ACT_GEGLU, ln_stats=st, ln_colsum=csg, ln_eps=1e-5), 5),
        ("1280->320 bias + residual + stats out", 320, 1280, lambda: ops.gemm(x1280, w1280, out=o320, bias=b320, residual=h, beta=1.0, row_stats=True), 6),
    ]
    if a.chain:
        # consumer timed right after its producer, L2 NOT flushed in between: forward-after-forward against reverse-after-forward
        from saspa_aug_b200 import _lib
        from kernel_bench import FLUSH  # noqa: F401

        lib = _lib.load()
        x0 = rnd(M, 320)
        flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")
        cons = [("320->320 bias + residual", lambda: ops.gemm(x, w320, out=o320, bias=b320, residual=h, beta=1.0)),
                ("320->960 (QKV) LN folded", lambda: ops.gemm(x, w960, out=o960, bias=b960, ln_stats=st, ln_colsum=cs960, ln_eps=1e-5)),
                ("320->2560 GEGLU LN folded", lambda: ops.gemm(x, wg, out=o1280, bias=bg, act=ops.ACT_GEGLU, ln_stats=st, ln_colsum=csg, ln_eps=1e-5))]
        for name, fn in cons:
            for rev in (0, 1):
                tot = 0.0
                for rep in range(8):
                    flush.zero_()
                    lib.saspa_gemm_reverse_m(0)
                    ops.gemm(x0, w320, out=x, bias=b320)  # producer writes x front to back
                    lib.saspa_gemm_reverse_m(rev)
                    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    s.record()
                    fn()
                    e.record()
                    torch.cuda.synchronize()
                    if rep >= 2:
                        tot += s.elapsed_time(e)
                lib.saspa_gemm_reverse_m(0)
                print(f"chain: {name:32s} consumer {'reversed' if rev else 'forward '}: {tot / 6 * 1e3:8.1f} us")
        return
    if a.only >= 0:
        for _ in range(3):
            variants[a.only][3]()
        torch.cuda.synchronize()
        return
    from saspa_aug_b200 import _lib

    lib = _lib.load()
    combos = [(0, 0)] if not a.sweep else [(0, 0), (2, 0), (1, 256), (2, 256), (2, 128)]
    for name, N, K, fn, units in variants:
      for ctas, bn in combos:
        lib.saspa_gemm_force_ctas(ctas)
        lib.saspa_gemm_force_bn(bn)
        try:
            ms = timeit(fn, iters=6, warmup=2)
        except Exception as e:  # noqa: BLE001 -- a forced tile the shape cannot take
            print(f"{name:42s} ctas={ctas} bn={bn}: {str(e)[:80]}")
            continue
        finally:
            lib.saspa_gemm_force_ctas(0)
            lib.saspa_gemm_force_bn(0)
        name = f"{name.split(' [')[0]} [ctas={ctas or 'auto'} bn={bn or 'auto'}]" if a.sweep else name
        # units: 320-column bf16 row blocks moved through HBM (in + out + residual), M * 640 B each
        gb = units * M * 640 / 1e9 if K == 320 else (M * 1280 * 2 + 2 * M * 640) / 1e9
        print(f"{name:62s} {ms * 1e3:8.1f} us  {2.0 * M * N * K / ms / 1e9:7.0f} TF/s  {gb / ms * 1e3:6.0f} GB/s algorithmic HBM")


if __name__ == "__main__":
    main()
