"""Generates tests/golden/canny_golden.json + canny_small.npz by running the REFERENCE's own
all_utils.utils.generate_canny (-> cv2.Canny) from /root/reference on the synthetic generators.
Run in the build container only (the reference does not exist on the GPU box):
    python tests/golden/make_canny_golden.py
"""
import hashlib
import json
import os
import sys

import numpy as np
from PIL import Image

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref_import  # noqa: E402
from saspa_aug_b200.synthetic import synthetic_source  # noqa: E402

LOW, HIGH, RES = 120, 200, 512  # run_aug/run_aug.py:543-544, :536


def main():
    ru = ref_import.import_reference_utils()
    import cv2

    cases = []
    for kind in ("blobs", "noise", "smooth"):
        for seed in range(4):
            for h, w in ((512, 512), (512, 704)):
                img = synthetic_source(seed, h, w, kind)
                out = np.array(ru.generate_canny(Image.fromarray(img), LOW, HIGH, RES))
                assert out.shape == (h, w, 3)
                cases.append({"kind": kind, "seed": seed, "h": h, "w": w, "sha256_hwc3": hashlib.sha256(out.tobytes()).hexdigest(),
                              "edge_pixels": int((out[..., 0] > 0).sum())})
    json.dump({"low": LOW, "high": HIGH, "resolution": RES, "opencv": cv2.__version__, "generator": "all_utils.utils.generate_canny",
               "cases": cases}, open(os.path.join(ROOT, "tests/golden/canny_golden.json"), "w"), indent=1)
    # small explicit vectors (full maps) through the reference's CannyDetector (all_utils/utils.py:81-85)
    small = {}
    for i, (h, w) in enumerate(((64, 64), (48, 80), (33, 70))):
        img = synthetic_source(100 + i, h, w, "blobs" if i < 2 else "noise")
        small[f"img{i}"] = img
        small[f"edge{i}"] = ru.apply_canny(img, LOW, HIGH)
    np.savez_compressed(os.path.join(ROOT, "tests/golden/canny_small.npz"), **small)
    print("wrote", len(cases), "hash cases and", len(small) // 2, "small vectors")


if __name__ == "__main__":
    main()
