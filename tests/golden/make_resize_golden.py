"""Golden vectors for the PIL-exact resize used by the filter transforms, produced by the same calls the
reference makes: torchvision.transforms.Resize on PIL images (all_utils/dataset_utils.py:78-85) and the
openai-clip _transform (Resize(224, BICUBIC)).  Run in the build container:
    python tests/golden/make_resize_golden.py
"""
import hashlib
import json
import os
import sys

import numpy as np
import torchvision.transforms as T
from PIL import Image

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from saspa_aug_b200.synthetic import synthetic_source  # noqa: E402


def main():
    cases = []
    baseline_resize = T.Resize((256, 256))  # dataset_utils.py:80 (default bilinear, antialias on PIL)
    clip_resize = T.Resize(224, interpolation=T.InterpolationMode.BICUBIC)
    for seed in range(3):
        for kind in ("blobs", "noise"):
            for h, w in ((512, 512), (512, 704)):
                img = Image.fromarray(synthetic_source(seed, h, w, kind))
                a = np.array(baseline_resize(img))
                b = np.array(clip_resize(img))
                cases.append({"seed": seed, "kind": kind, "h": h, "w": w,
                              "bilinear_256": {"shape": list(a.shape), "sha256": hashlib.sha256(a.tobytes()).hexdigest()},
                              "bicubic_224": {"shape": list(b.shape), "sha256": hashlib.sha256(b.tobytes()).hexdigest()}})
    import PIL
    json.dump({"pillow": PIL.__version__, "cases": cases}, open(os.path.join(ROOT, "tests/golden/resize_golden.json"), "w"), indent=1)
    print("wrote", len(cases))


if __name__ == "__main__":
    main()
