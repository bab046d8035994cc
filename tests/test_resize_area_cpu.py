"""The INTER_AREA restatement (oracle/cv2_area.py) against the installed cv2 itself -- the library call the reference's resize_image makes
(all_utils/utils.py:58-79) -- bit for bit, on all three OpenCV code paths, and on the sizes resize_image produces for typical sources."""
import numpy as np
import pytest

cv2 = pytest.importorskip("cv2")

from oracle import cv2_area
from saspa_aug_b200 import run_aug


@pytest.mark.parametrize("sh,sw,dh,dw", [
    (37, 53, 16, 24), (100, 150, 64, 96), (333, 500, 192, 256), (61, 61, 32, 32), (700, 1000, 512, 704),      # float area (non-integer scales)
    (64, 96, 32, 48), (96, 96, 32, 32), (128, 192, 64, 64), (90, 120, 30, 60), (256, 256, 64, 64), (100, 160, 50, 40),  # integer scales
    (75, 50, 64, 64), (525, 700, 512, 704), (700, 525, 704, 512), (130, 60, 128, 64), (40, 100, 64, 64),      # one axis up-scaled
    (64, 64, 64, 64)])
def test_area_restatement_equals_cv2(sh, sw, dh, dw):
    src = np.random.default_rng(sh * 1000 + sw).integers(0, 256, (sh, sw, 3), dtype=np.uint8)
    want = cv2.resize(src, (dw, dh), interpolation=cv2.INTER_AREA)
    got = cv2_area.resize_area(src, dw, dh)
    assert got.shape == want.shape and np.array_equal(got, want)


@pytest.mark.parametrize("hw", [(695, 1024), (768, 1024), (525, 700), (1200, 1600), (512, 683), (600, 600), (1365, 2048)])
def test_area_restatement_on_resize_image_shapes(hw):
    """The shapes the reference's rule produces (short side 512, both sides rounded to x64, 1.2 MP cap) from typical FGVC source sizes."""
    src = np.random.default_rng(hw[0]).integers(0, 256, (*hw, 3), dtype=np.uint8)
    want = run_aug.resize_image(src, 512)  # cv2 on the host: the product's (and the reference's) path
    H, W, k = run_aug.resized_hw(hw[0], hw[1], 512)
    assert k <= 1
    assert np.array_equal(cv2_area.resize_area(src, W, H), want)


@pytest.mark.parametrize("sh,sw,dh,dw", [(30, 40, 64, 64), (375, 500, 512, 704), (300, 400, 512, 704), (100, 100, 128, 192), (64, 64, 128, 128), (333, 250, 704, 512),
                                         (200, 200, 512, 512), (97, 131, 256, 320)])
def test_lanczos4_restatement_equals_cv2(sh, sw, dh, dw):
    """INTER_LANCZOS4 (resize_image's k > 1 branch: sources under the resolution) restated and compared with the installed cv2 bit for bit."""
    src = np.random.default_rng(sh * 1000 + sw).integers(0, 256, (sh, sw, 3), dtype=np.uint8)
    want = cv2.resize(src, (dw, dh), interpolation=cv2.INTER_LANCZOS4)
    assert np.array_equal(cv2_area.resize_lanczos4(src, dw, dh), want)
    if min(sh, sw) < 512 and (dh, dw) == run_aug.resized_hw(sh, sw, 512)[:2]:
        assert np.array_equal(run_aug.resize_image(src, 512), want)
