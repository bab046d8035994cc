"""Launches the tcgen05 attention kernel a few times on one SD v1.5 shape (for ncu captures)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))  # repo root (python tools/<name>.py)
import sys

import torch

from saspa_aug_b200 import _lib, ops
from kernel_bench import rnd

b, heads, tq, tkv, d = [int(x) for x in (sys.argv[1:6] if len(sys.argv) > 5 else (32, 8, 4096, 4096, 40))]
_lib.load().saspa_attention_impl(2)
qkv = rnd(b, tq, 3 * heads * d)
c = heads * d
out = torch.empty(b, tq, c, dtype=torch.bfloat16, device="cuda")
for _ in range(3):
    ops.attention(qkv[..., :c], qkv[..., c:2 * c], qkv[..., 2 * c:], heads, out=out)
torch.cuda.synchronize()
