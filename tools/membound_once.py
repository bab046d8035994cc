"""One launch of each memory-bound kernel of the path at its BASELINE size, for `ncu --set full -k regex:<name>` captures
(Canny 256 x 512^2, PIL-exact resize 512 -> 256 / 224, GroupNorm+SiLU at the 64x64 and 32x32 levels, LayerNorm 262144 x 320, the
filter heads).  Not a benchmark: numbers printed under ncu are never bench values."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from saspa_aug_b200 import ops
from saspa_aug_b200.synthetic import synthetic_source


def main():
    dev = "cuda"
    g = torch.Generator().manual_seed(0)
    src = torch.from_numpy(np.stack([synthetic_source(s % 8, kind="blobs") for s in range(64)])).to(dev)
    big = src.repeat(4, 1, 1, 1).contiguous()  # 256 x 512 x 512 x 3
    x64 = torch.randn((64, 4096, 320), generator=g).to(torch.bfloat16).to(dev)
    x32 = torch.randn((64, 1024, 640), generator=g).to(torch.bfloat16).to(dev)
    gam, bet = torch.ones(640, device=dev), torch.zeros(640, device=dev)
    rows = torch.randn((262144, 320), generator=g).to(torch.bfloat16).to(dev)
    for rep in range(2):  # second round = warm module / attribute state
        ops.canny(big, 120, 200, out_channels=1, want_ctrl=True)
        r = ops.resize_pil(src, 256, 256, "bilinear")
        ops.crop_normalize(r, 16, 16, 224, 224, (0.485, 0.456, 0.406), (0.229, 0.224, 0.225), out_c=8)
        ops.resize_pil(src, 224, 224, "bicubic")
        ops.groupnorm(x64, 32, 1e-5, gam[:320], bet[:320], ops.ACT_SILU)
        ops.groupnorm(x32, 32, 1e-5, gam, bet, ops.ACT_SILU)
        ops.layernorm(rows, 1e-5, gam[:320], bet[:320])
        torch.cuda.synchronize()
    print("membound_once done")


if __name__ == "__main__":
    main()
