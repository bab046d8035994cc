"""Where does the d = 40 self-attention time go?  Times the tcgen05 kernel with parts disabled (results are wrong on purpose)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))  # repo root (python tools/<name>.py)
import ctypes
import torch
from saspa_aug_b200 import _lib, ops
from kernel_bench import rnd, timeit

lib = ctypes.CDLL(_lib.SO_PATH)
_lib.load().saspa_attention_impl(2)
for (b, heads, t, d) in [(32, 8, 4096, 40), (32, 8, 1024, 80), (16, 16, 4096, 64), (8, 16, 4096, 128)]:
    qkv = rnd(b, t, 3 * heads * d)
    c = heads * d
    out = torch.empty(b, t, c, dtype=torch.bfloat16, device="cuda")
    for flags, name in [(0, "full"), (1, "no exp"), (2, "no PV mma"), (4, "no QK mma"), (6, "no mma"), (7, "no exp, no mma"), (3, "no exp no PV")]:
        lib.saspa_attention_debug(flags)
        ms = timeit(lambda: ops.attention(qkv[..., :c], qkv[..., c:2 * c], qkv[..., 2 * c:], heads, out=out), iters=5)
        print(f"d{d} t{t} {name:16s}: {ms:.3f} ms")
    lib.saspa_attention_debug(0)

# ---- phase timers of one CTA (clock64 sums over its 32 key tiles) ----
lib.saspa_attention_trace.argtypes = [ctypes.c_void_p]
lib.saspa_attention_trace.restype = None
for (b, heads, t, d) in [(32, 8, 4096, 40), (8, 16, 4096, 128)]:
    qkv = rnd(b, t, 3 * heads * d)
    c = heads * d
    out = torch.empty(b, t, c, dtype=torch.bfloat16, device="cuda")
    ref = None
    for flags in (0,):
        lib.saspa_attention_debug(flags)
        buf = torch.zeros(24, dtype=torch.int64, device="cuda")
        for _ in range(2):
            ops.attention(qkv[..., :c], qkv[..., c:2 * c], qkv[..., 2 * c:], heads, out=out)
        lib.saspa_attention_trace(buf.data_ptr())
        ops.attention(qkv[..., :c], qkv[..., c:2 * c], qkv[..., 2 * c:], heads, out=out)
        torch.cuda.synchronize()
        lib.saspa_attention_trace(None)
        if ref is None:
            ref = out.float().clone()
        v = buf.cpu().tolist()
        steps = (t + 127) // 128 if d <= 64 else (t + 63) // 64
        print(f"d{d} t{t} flags {flags}: clocks per key tile ({steps} tiles), CTA (1,3); max |out - out(flags 0)| = {(out.float() - ref).abs().max().item():.3g}")
        print("  issuer tile 0: " + ", ".join(f"{n} {x / steps:.0f}" for n, x in zip(["prologue", "wait K", "wait S-free", "-", "QK issue", "wait V", "wait P", "PV issue"], v[:8])))
        for i in (0, 1):
            print(f"  softmax tile {i}: " + ", ".join(f"{n} {x / steps:.0f}" for n, x in zip(["wait S", "TMEM ld", "max/rescale", "exp+pack", "wait PV", "TMEM st"], v[8 + 8 * i: 14 + 8 * i])))
    lib.saspa_attention_debug(0)

# ---- FMA-pipe share of the exponentials (needs a build with -DSASPA_ATTN_POLY_SWEEP; otherwise every row shows the default) ----
lib.saspa_attention_poly.argtypes = [ctypes.c_int]
lib.saspa_attention_poly.restype = None
for (b, heads, t, d) in [(32, 8, 4096, 40), (16, 16, 4096, 64)]:
    qkv = rnd(b, t, 3 * heads * d)
    c = heads * d
    out = torch.empty(b, t, c, dtype=torch.bfloat16, device="cuda")
    for pe in (0, 8, 6, 4, 3, 2):
        lib.saspa_attention_poly(pe)
        ms = timeit(lambda: ops.attention(qkv[..., :c], qkv[..., c:2 * c], qkv[..., 2 * c:], heads, out=out), iters=5)
        print(f"d{d} t{t} poly every {pe}: {ms:.3f} ms")
    lib.saspa_attention_poly(-1)
