"""Filter-side stages of SURVEY.md 8(d) on B200: PIL-exact resize + normalise (HBM-bound, integer), WSDAN_CAL-R50 classifier, CLIP RN50 and
CLIP ViT-L/14 image towers (tensor-bound), and the whole AugmentationFilter on device-resident u8 images.  CUDA events, 3 warm-ups.
FLOP figures: SURVEY.md C.8 (12.5 / ~12 / 162 GFLOP per image)."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))  # repo root (python tools/<name>.py)
import numpy as np
import torch

from saspa_aug_b200 import checkpoints as ck
from saspa_aug_b200 import ops
from saspa_aug_b200.filter_nets import CLIP_MEAN, CLIP_STD, IMAGENET_MEAN, IMAGENET_STD, AugmentationFilter, CLIPRN50, CLIPViT, WSDANClassifier
from saspa_aug_b200.synthetic import synthetic_source, synthetic_token_ids

PEAKS = {"hbm_gbs": 6456.0, "bf16_tflops_sustained": 1400.0}
try:
    PEAKS.update(json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json"))))
except Exception:
    pass


def timeit(fn, iters=5, warmup=3):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(iters):
        fn()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / iters


def main():
    n = 256
    base = torch.from_numpy(np.stack([synthetic_source(s) for s in range(16)])).cuda()
    imgs = base.repeat(n // 16, 1, 1, 1).contiguous()
    labels = (torch.arange(n, dtype=torch.int32) % 100).cuda()

    def prep_cls():
        r = ops.resize_pil(imgs, 256, 256, "bilinear")
        return ops.crop_normalize(r, 16, 16, 224, 224, IMAGENET_MEAN, IMAGENET_STD, out_c=8)

    def prep_clip(c):
        r = ops.resize_pil(imgs, 224, 224, "bicubic")
        return ops.crop_normalize(r, 0, 0, 224, 224, CLIP_MEAN, CLIP_STD, out_c=c)

    ms = timeit(prep_cls)
    byt = n * (3 * 512 * 512 + 3 * 224 * 224 * 2)
    print(f"resize(256, bilinear, PIL-exact) + crop 224 + normalise, {n} x 512^2: {ms:.3f} ms = {n / ms * 1e3:.0f} img/s, {byt / ms / 1e6:.0f} GB/s algorithmic "
          f"({byt / ms / 1e6 / PEAKS['hbm_gbs']:.3f} of HBM peak)")
    wsd = WSDANClassifier(ck.random_filter_state_dict(ck.wsdan_shapes(100, "resnet50"), 4242), 100, "resnet50")
    x = prep_cls()
    for mb in (64, 256):
        ms = timeit(lambda: [wsd(x[i:i + mb]) for i in range(0, n, mb)])
        print(f"WSDAN_CAL-R50 logits, batch {mb}: {ms / n * 1e3:.1f} us/img = {n / ms * 1e3:.0f} img/s, {12.5 * n / ms:.0f} TFLOP/s ({12.5 * n / ms / PEAKS['bf16_tflops_sustained']:.2f} of sustained bf16 peak)")
    rn = CLIPRN50(ck.random_filter_state_dict(ck.clip_rn50_shapes(), 777))
    x8 = prep_clip(8)
    ms = timeit(lambda: [rn.encode_image(x8[i:i + 64]) for i in range(0, n, 64)])
    print(f"CLIP RN50 image tower, batch 64: {ms / n * 1e3:.1f} us/img = {n / ms * 1e3:.0f} img/s, {12.0 * n / ms:.0f} TFLOP/s ({12.0 * n / ms / PEAKS['bf16_tflops_sustained']:.2f})")
    vit = CLIPViT(ck.random_filter_state_dict(ck.clip_vit_shapes(), 777))
    x3 = prep_clip(3)
    ms = timeit(lambda: [vit.encode_image(x3[i:i + 64]) for i in range(0, n, 64)])
    print(f"CLIP ViT-L/14 image tower, batch 64: {ms / n * 1e3:.1f} us/img = {n / ms * 1e3:.0f} img/s, {162.0 * n / ms:.0f} TFLOP/s ({162.0 * n / ms / PEAKS['bf16_tflops_sustained']:.2f})")
    ids = torch.cat([synthetic_token_ids(s) for s in range(7)])
    for name, clip in (("RN50", rn), ("ViT-L/14", vit)):
        flt = AugmentationFilter(wsd, clip, ids, conf_top_k=10, micro_batch=64)
        ms = timeit(lambda: flt(imgs, labels))
        print(f"whole filter (resize x2, WSDAN-R50 top-10, CLIP {name} semantic), {n} device-resident 512^2 images: {ms:.1f} ms = {n / ms * 1e3:.0f} img/s")


if __name__ == "__main__":
    main()
