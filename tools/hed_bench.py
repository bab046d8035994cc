"""HED conditioning throughput: HEDdetector.detect_batch on n synthetic 512x512 sources (CUDA events, L2-sized inputs rotate), split into
the network (13 VGG convolutions + 5 projections, 2 x 87.6 GFLOP per image) and the fused tail kernel.  Usage: python tools/hed_bench.py [--n 16]"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from saspa_aug_b200 import checkpoints as ck
from saspa_aug_b200 import ops
from saspa_aug_b200.hed import HEDdetector
from saspa_aug_b200.synthetic import synthetic_source


def timed(fn, reps=5):
    fn()
    torch.cuda.synchronize()
    best = None
    for _ in range(reps):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        fn()
        e.record()
        torch.cuda.synchronize()
        ms = s.elapsed_time(e)
        best = ms if best is None else min(best, ms)
    return best


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=16)
    a = ap.parse_args()
    det = HEDdetector.from_state_dict(ck.random_hed_state_dict(1), "cuda")
    imgs = torch.from_numpy(np.stack([synthetic_source(s) for s in range(a.n)])).cuda()
    flops = 0.0
    hw = 512 * 512
    for k, (cin, cout, layers) in enumerate(ck.HED_BLOCKS):
        px = hw >> (2 * k)
        flops += 2.0 * px * (9 * cin * cout + (layers - 1) * 9 * cout * cout + cout)
    t_all = timed(lambda: det.detect_batch(imgs))
    sides = det.netNetwork(imgs)
    t_tail = timed(lambda: ops.hed_fuse(sides, 512, 512))
    t_net = timed(lambda: det.netNetwork(imgs))
    tail_bytes = a.n * (hw * 3 + sum((hw >> (2 * k)) * 4 for k in range(5)))
    print(f"HED detect_batch n={a.n} 512x512: {t_all:.2f} ms = {t_all / a.n:.3f} ms/image ({a.n / t_all * 1e3:.0f} images/s); network {t_net:.2f} ms = "
          f"{flops * a.n / t_net / 1e9:.0f} TFLOP/s ({flops / 1e9:.1f} GFLOP/image); tail kernel {t_tail * 1e3:.1f} us = {tail_bytes / t_tail / 1e6:.0f} GB/s "
          f"({tail_bytes / a.n / 1e6:.2f} MB/image algorithmic)")


if __name__ == "__main__":
    main()
