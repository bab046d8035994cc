"""Per-kernel microbenchmarks on a real B200 (CUDA events, L2-flushed between timed launches).
Prints achieved TFLOP/s or GB/s against MEASURED_PEAKS.json.  Not the bench.py contract; a tuning aid."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))  # repo root (python tools/<name>.py)
import json
import os
import sys

import torch

from saspa_aug_b200 import ops

PEAKS = json.load(open("MEASURED_PEAKS.json")) if os.path.exists("MEASURED_PEAKS.json") else {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0}
FLUSH = None


def timeit(fn, iters=10, warmup=3, flush=True):
    global FLUSH
    if FLUSH is None:
        FLUSH = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    tot = 0.0
    for _ in range(iters):
        if flush:
            FLUSH.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        fn()
        e.record()
        torch.cuda.synchronize()
        tot += s.elapsed_time(e)
    return tot / iters


def rnd(*shape):
    return torch.randn(*shape, device="cuda").to(torch.bfloat16)


def main():
    rows = []
    B = 32  # CFG rows (16 images)
    only = sys.argv[1] if len(sys.argv) > 1 else "all"
    if only == "hbm":
        return hbm(B)
    if only != "conv":
        gemm_section(B, only)
    if only == "gemm":
        return
    conv_section(B, only)


def gemm_section(B, only):
    print("== GEMM (tcgen05) ==")
    for M, N, K in [(8192, 8192, 8192), (B * 4096, 320, 320), (B * 4096, 2560, 320), (B * 4096, 320, 1280), (B * 1024, 640, 640), (B * 1024, 5120, 640),
                    (B * 256, 1280, 1280), (B * 256, 10240, 1280), (B * 256, 1280, 5120), (B * 64, 1280, 1280), (B * 4096, 320, 2880)]:
        a, b = rnd(M, K), rnd(N, K)
        out = torch.empty(M, N, dtype=torch.bfloat16, device="cuda")
        ms = timeit(lambda: ops.gemm(a, b, out=out))
        tf = 2.0 * M * N * K / ms / 1e9
        ref_ms = timeit(lambda: torch.matmul(a, b.t()))
        print(f"gemm {M}x{N}x{K}: {ms:.3f} ms {tf:.0f} TF/s ({tf / PEAKS['bf16_tflops']:.2f} of measured peak) | cuBLAS {2.0*M*N*K/ref_ms/1e9:.0f} TF/s")
    print("== GEMM epilogue variants (GEGLU / in-place residual / wide QKV) ==")
    for M, N, K, kind in [(B * 4096, 2560, 320, "geglu"), (B * 1024, 5120, 640, "geglu"), (B * 256, 10240, 1280, "geglu"), (B * 4096, 320, 320, "res"),
                          (B * 1024, 640, 640, "res"), (B * 256, 1280, 1280, "res"), (B * 4096, 320, 1280, "res"), (B * 4096, 960, 320, "plain"),
                          (B * 1024, 1920, 640, "plain")]:
        a, b = rnd(M, K), rnd(N, K) * 0.05
        bias = torch.zeros(N, device="cuda")
        if kind == "geglu":
            out = torch.empty(M, N // 2, dtype=torch.bfloat16, device="cuda")
            fn = lambda: ops.gemm(a, b, out=out, bias=bias, act=ops.ACT_GEGLU)
            byt = 2.0 * (M * K + M * N // 2)
        elif kind == "res":
            out = rnd(M, N)
            fn = lambda: ops.gemm(a, b, out=out, bias=bias, residual=out, beta=1.0)
            byt = 2.0 * (M * K + 2 * M * N)
        else:
            out = torch.empty(M, N, dtype=torch.bfloat16, device="cuda")
            fn = lambda: ops.gemm(a, b, out=out)
            byt = 2.0 * (M * K + M * N)
        ms = timeit(fn)
        tf = 2.0 * M * N * K / ms / 1e9
        print(f"gemm[{kind}] {M}x{N}x{K}: {ms:.3f} ms {tf:.0f} TF/s ({tf / PEAKS['bf16_tflops']:.2f} of peak), {byt / ms / 1e6:.0f} GB/s algorithmic")


def conv_section(B, only):
    print("== conv3x3 implicit GEMM ==")
    for n, h, w, cin, cout in [(B, 64, 64, 320, 320), (B, 32, 32, 640, 640), (B, 16, 16, 1280, 1280), (B, 8, 8, 1280, 1280), (B, 64, 64, 640, 320),
                               (B, 16, 16, 2560, 1280), (4, 512, 512, 128, 128), (4, 256, 256, 256, 256), (8, 128, 128, 512, 512), (16, 64, 64, 512, 512)]:
        x, wk = rnd(n, h, w, cin), rnd(cout, 9 * cin)
        out = torch.empty(n, h, w, cout, dtype=torch.bfloat16, device="cuda")
        from saspa_aug_b200 import _lib
        res = []
        for impl in (1, 0):
            _lib.load().saspa_conv_impl(impl)
            ms = timeit(lambda: ops.conv2d_igemm(x, wk, 3, out=out))
            res.append((ms, 2.0 * n * h * w * cout * 9 * cin / ms / 1e9))
        _lib.load().saspa_conv_impl(0)
        print(f"conv {n}x{h}x{w} {cin}->{cout}: per-tap {res[0][0]:.3f} ms {res[0][1]:.0f} TF/s | auto(halo) {res[1][0]:.3f} ms {res[1][1]:.0f} TF/s "
              f"({res[1][1] / PEAKS['bf16_tflops']:.2f} of peak)")
    if only == "conv":
        return
    print("== attention (mma.sync flash) ==")
    for b, heads, tq, tkv, d in [(B, 8, 4096, 4096, 40), (B, 8, 1024, 1024, 80), (B, 8, 256, 256, 160), (B, 8, 4096, 77, 40), (B, 8, 1024, 77, 80)]:
        q, k, v = rnd(b, tq, heads * d), rnd(b, tkv, heads * d), rnd(b, tkv, heads * d)
        out = torch.empty_like(q)
        ms = timeit(lambda: ops.attention(q, k, v, heads, out=out))
        tf = 4.0 * b * heads * tq * tkv * d / ms / 1e9
        print(f"attn b{b} h{heads} {tq}x{tkv} d{d}: {ms:.3f} ms {tf:.0f} TF/s")
    hbm(B)


def hbm(B):
    print("== HBM-bound ==")
    for n, hw, c in [(B, 4096, 320), (B, 1024, 640), (B, 256, 1280), (B, 4096, 640), (B, 4096, 960), (B, 1024, 1920), (B, 64, 2560), (2 * B, 1024, 640), (2 * B, 256, 1280), (2 * B, 4096, 320), (8, 262144, 128), (8, 65536, 256)]:
        x = rnd(n, hw, c)
        out = torch.empty_like(x)
        g, bt = torch.ones(c, device="cuda"), torch.zeros(c, device="cuda")
        from saspa_aug_b200 import _lib
        res = []
        for impl in (1, 0):
            _lib.load().saspa_groupnorm_impl(impl)
            ms = timeit(lambda: ops.groupnorm(x, 32, 1e-5, g, bt, ops.ACT_SILU, out=out))
            res.append((ms, 2.0 * x.numel() * 2 / ms / 1e6))  # algorithmic: read once + write once
        _lib.load().saspa_groupnorm_impl(0)
        print(f"groupnorm+silu {n}x{hw}x{c}: two-pass {res[0][0]:.3f} ms {res[0][1]:.0f} GB/s | auto {res[1][0]:.3f} ms {res[1][1]:.0f} GB/s algorithmic "
              f"({res[1][1] / PEAKS['hbm_gbs']:.2f} of HBM peak)")
    for rows_, c in [(B * 4096, 320), (2 * B * 4096, 320), (2 * B * 1024, 640), (B * 1024, 640), (B * 256, 1280), (B * 77, 768)]:
        x = rnd(rows_, c)
        out = torch.empty_like(x)
        g, bt = torch.ones(c, device="cuda"), torch.zeros(c, device="cuda")
        ms = timeit(lambda: ops.layernorm(x, 1e-5, g, bt, out=out))
        gb = 2.0 * x.numel() * 2 / ms / 1e6
        print(f"layernorm {rows_}x{c}: {ms:.3f} ms {gb:.0f} GB/s ({gb / PEAKS['hbm_gbs']:.2f})")
    imgs = torch.randint(0, 256, (256, 512, 512, 3), dtype=torch.uint8, device="cuda")
    import numpy as np
    from saspa_aug_b200.synthetic import synthetic_source
    base = torch.from_numpy(np.stack([synthetic_source(s) for s in range(16)])).cuda()
    blobs = base.repeat(16, 1, 1, 1).contiguous()
    for name, t in [("noise", imgs), ("blobs", blobs)]:
        ms = timeit(lambda: ops.canny(t, 120, 200))
        print(f"canny 256x512x512 {name}: {ms:.3f} ms {4.0 * 256 * 512 * 512 / ms / 1e6:.0f} GB/s algorithmic ({4.0*256*512*512/ms/1e6/PEAKS['hbm_gbs']:.2f})")


if __name__ == "__main__":
    main()
