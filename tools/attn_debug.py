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
for (b, heads, t, d) in [(32, 8, 4096, 40), (32, 8, 1024, 80)]:
    qkv = rnd(b, t, 3 * heads * d)
    c = heads * d
    out = torch.empty(b, t, c, dtype=torch.bfloat16, device="cuda")
    for flags, name in [(0, "full"), (1, "no exp"), (2, "no PV mma"), (4, "no QK mma"), (6, "no mma"), (7, "no exp, no mma"), (3, "no exp no PV")]:
        lib.saspa_attention_debug(flags)
        ms = timeit(lambda: ops.attention(qkv[..., :c], qkv[..., c:2 * c], qkv[..., 2 * c:], heads, out=out), iters=5)
        print(f"d{d} t{t} {name:16s}: {ms:.3f} ms")
    lib.saspa_attention_debug(0)
