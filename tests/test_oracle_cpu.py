"""CPU tests: the oracle restatements against the committed golden vectors (made by the reference itself,
tests/golden/make_*.py) and, when /root/reference is present, against the reference directly."""
import hashlib
import json
import os

import numpy as np
import pytest

from oracle import canny_np, clib, ref_import
from saspa_aug_b200.synthetic import synthetic_source

G = os.path.join(os.path.dirname(__file__), "golden")


def test_canny_c_oracle_matches_golden_hashes():
    gold = json.load(open(os.path.join(G, "canny_golden.json")))
    assert len(gold["cases"]) >= 20
    for rec in gold["cases"]:
        img = synthetic_source(rec["seed"], rec["h"], rec["w"], rec["kind"])
        e = clib.canny(img, gold["low"], gold["high"])
        hwc3 = np.repeat(e[..., None], 3, axis=2)
        assert hashlib.sha256(hwc3.tobytes()).hexdigest() == rec["sha256_hwc3"], rec
        assert int((e > 0).sum()) == rec["edge_pixels"]


def test_canny_numpy_and_c_oracles_match_small_vectors():
    z = np.load(os.path.join(G, "canny_small.npz"))
    for i in range(3):
        img, edge = z[f"img{i}"], z[f"edge{i}"]
        assert (clib.canny(img, 120, 200) == edge).all()
        assert (canny_np.canny(img, 120, 200) == edge).all()


def test_canny_numpy_oracle_one_golden_full_size():
    gold = json.load(open(os.path.join(G, "canny_golden.json")))
    rec = gold["cases"][0]
    img = synthetic_source(rec["seed"], rec["h"], rec["w"], rec["kind"])
    out = canny_np.generate_canny_np(img, gold["low"], gold["high"], gold["resolution"])
    assert hashlib.sha256(out.tobytes()).hexdigest() == rec["sha256_hwc3"]


def test_canny_edge_cases():
    import cv2

    rng = np.random.default_rng(0)
    for shape in [(1, 1, 3), (2, 2, 3), (1, 17, 3), (17, 1, 3), (31, 33, 1)]:
        img = rng.integers(0, 256, shape, dtype=np.uint8)
        assert (clib.canny(img, 120, 200) == cv2.Canny(img if shape[2] == 3 else img[..., 0], 120, 200)).all()
    img = synthetic_source(1, 64, 64)
    assert (clib.canny(img, 200, 120) == clib.canny(img, 120, 200)).all()  # swapped thresholds
    assert clib.canny(np.zeros((0, 8, 8, 3), np.uint8), 1, 2).shape == (0, 8, 8)


@pytest.mark.skipif(not ref_import.available(), reason="/root/reference not present (GPU box)")
def test_canny_oracles_against_reference_live():
    from PIL import Image

    ru = ref_import.import_reference_utils()
    for seed, kind in [(11, "blobs"), (12, "noise")]:
        img = synthetic_source(seed, 512, 512, kind)
        ref = np.array(ru.generate_canny(Image.fromarray(img), 120, 200, 512))
        assert (np.repeat(clib.canny(img, 120, 200)[..., None], 3, 2) == ref).all()


def test_pil_resize_oracle_matches_golden():
    gold = json.load(open(os.path.join(G, "resize_golden.json")))
    for rec in gold["cases"]:
        img = synthetic_source(rec["seed"], rec["h"], rec["w"], rec["kind"])
        a = clib.pil_resize(img, 256, 256, "bilinear")
        assert hashlib.sha256(a.tobytes()).hexdigest() == rec["bilinear_256"]["sha256"]
        sh = rec["bicubic_224"]["shape"]
        b = clib.pil_resize(img, sh[0], sh[1], "bicubic")
        assert hashlib.sha256(b.tobytes()).hexdigest() == rec["bicubic_224"]["sha256"]


def test_pil_resize_oracle_live_pil():
    from PIL import Image

    img = synthetic_source(9, 300, 200, "noise")
    for oh, ow, f, pf in [(256, 256, "bilinear", Image.BILINEAR), (224, 150, "bicubic", Image.BICUBIC), (600, 640, "bicubic", Image.BICUBIC)]:
        assert (clib.pil_resize(img, oh, ow, f) == np.array(Image.fromarray(img).resize((ow, oh), pf))).all()
