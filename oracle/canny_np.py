"""NumPy restatement of the Canny edge conditioning used by SaSPA.

TEST INFRASTRUCTURE (see oracle/__init__.py).  Parity is PINNED: the functions
below are checked bit-for-bit against the reference's own
``all_utils.utils.generate_canny`` (-> ``cv2.Canny``) by
``tests/golden/make_canny_golden.py`` and ``tests/test_canny_oracle.py``.

Reference call chain (``/root/reference``):
  run_aug/run_aug.py:436-437      generate_canny(orig_img, 120, 200, 512)
  all_utils/utils.py:102-109      generate_canny  -> preprocess_canny
  all_utils/utils.py:87-99        preprocess_canny: HWC3 -> resize_image -> cv2.Canny -> HWC3
  all_utils/utils.py:81-85        CannyDetector.__call__ = cv2.Canny(img, low, high)
  all_utils/utils.py:39-55        HWC3
  all_utils/utils.py:58-79        resize_image (identity copy when input is already 512x512)

``cv2.Canny`` itself is third-party (opencv-python 4.8.0.74 pinned by
environment.yml:24; 4.13.0 installed here; algorithm unchanged).  Published
algorithm (modules/imgproc/src/canny.cpp, apertureSize=3, L2gradient=False):
  1. per channel Sobel 3x3 dx,dy (int16, BORDER_REPLICATE)
  2. mag = |dx|+|dy|; per pixel the channel with the largest mag wins (first on ties)
  3. non-maximum suppression against a magnitude map that is ZERO outside the image,
     direction sectors decided in 15-bit fixed point (tan22.5 = 13573 / 2^15)
  4. candidate iff mag > low, strong iff also mag > high (strict, thresholds floored)
  5. 8-connected hysteresis: candidates reachable from a strong pixel become edges
  6. output 255 / 0
"""
from __future__ import annotations

import numpy as np

TG22 = 13573  # (int)(0.4142135623730950488016887242097 * (1 << 15) + 0.5)


def hwc3(x: np.ndarray) -> np.ndarray:
    """all_utils/utils.py:39-55."""
    assert x.dtype == np.uint8
    if x.ndim == 2:
        x = x[:, :, None]
    h, w, c = x.shape
    assert c in (1, 3, 4)
    if c == 3:
        return x
    if c == 1:
        return np.concatenate([x, x, x], axis=2)
    color = x[:, :, 0:3].astype(np.float32)
    alpha = x[:, :, 3:4].astype(np.float32) / 255.0
    y = color * alpha + 255.0 * (1.0 - alpha)
    return y.clip(0, 255).astype(np.uint8)


def resized_shape(h: int, w: int, smaller_side_res: int) -> tuple[int, int, float]:
    """Size arithmetic of all_utils/utils.py:58-79 (the cv2.resize itself is out of
    scope for the synthetic square inputs, where it is the identity)."""
    max_res = 1200000
    H, W = float(h), float(w)
    k = float(smaller_side_res) / min(H, W)
    H *= k
    W *= k
    if H * W > max_res:
        k = np.sqrt(max_res / (H * W))
        H *= k
        W *= k
    H = int(np.round(H / 64.0)) * 64
    W = int(np.round(W / 64.0)) * 64
    return H, W, k


def sobel_argmax(img: np.ndarray):
    """Steps 1-2.  img: [H,W,C] u8 -> (dx, dy, mag) int32 [H,W] of the winning channel."""
    if img.ndim == 2:
        img = img[:, :, None]
    p = np.pad(img.astype(np.int32), ((1, 1), (1, 1), (0, 0)), mode="edge")
    tl, tc, tr = p[:-2, :-2], p[:-2, 1:-1], p[:-2, 2:]
    ml, mr = p[1:-1, :-2], p[1:-1, 2:]
    bl, bc, br = p[2:, :-2], p[2:, 1:-1], p[2:, 2:]
    dx = (tr + 2 * mr + br) - (tl + 2 * ml + bl)
    dy = (bl + 2 * bc + br) - (tl + 2 * tc + tr)
    mag = np.abs(dx) + np.abs(dy)
    idx = np.argmax(mag, axis=2)  # first maximal channel
    ii, jj = np.meshgrid(np.arange(img.shape[0]), np.arange(img.shape[1]), indexing="ij")
    return dx[ii, jj, idx], dy[ii, jj, idx], mag[ii, jj, idx]


def nms_labels(dx, dy, mag, low: int, high: int) -> np.ndarray:
    """Steps 3-4.  Returns u8 labels: 0 = not an edge, 1 = weak candidate, 2 = strong."""
    H, W = mag.shape
    m = np.pad(mag, 1)  # zero border
    c = m[1:-1, 1:-1]
    L, R = m[1:-1, :-2], m[1:-1, 2:]
    U, D = m[:-2, 1:-1], m[2:, 1:-1]
    UL, UR = m[:-2, :-2], m[:-2, 2:]
    DL, DR = m[2:, :-2], m[2:, 2:]
    ax = np.abs(dx).astype(np.int64)
    ay = np.abs(dy).astype(np.int64) << 15
    tg22x = ax * TG22
    tg67x = tg22x + (ax << 16)
    horiz = ay < tg22x
    vert = (~horiz) & (ay > tg67x)
    diag = ~(horiz | vert)
    neg = (dx ^ dy) < 0
    keep = np.zeros((H, W), bool)
    keep |= horiz & (c > L) & (c >= R)
    keep |= vert & (c > U) & (c >= D)
    keep |= diag & neg & (c > UR) & (c > DL)
    keep |= diag & (~neg) & (c > UL) & (c > DR)
    keep &= c > low
    lab = np.zeros((H, W), np.uint8)
    lab[keep] = 1
    lab[keep & (c > high)] = 2
    return lab


def hysteresis(lab: np.ndarray) -> np.ndarray:
    """Step 5-6: 8-connected flood from label 2 through label 1 -> u8 0/255."""
    H, W = lab.shape
    strong = lab == 2
    cand = lab >= 1
    while True:
        p = np.pad(strong, 1)
        grown = (
            p[:-2, :-2] | p[:-2, 1:-1] | p[:-2, 2:] | p[1:-1, :-2] | p[1:-1, 2:] | p[2:, :-2] | p[2:, 1:-1] | p[2:, 2:]
        )
        new = strong | (grown & cand)
        if (new == strong).all():
            break
        strong = new
    return strong.astype(np.uint8) * 255


def canny(img: np.ndarray, low_threshold, high_threshold) -> np.ndarray:
    """cv2.Canny(img, low, high) for u8 [H,W] / [H,W,C] input -> u8 [H,W] in {0,255}."""
    low, high = float(low_threshold), float(high_threshold)
    if low > high:
        low, high = high, low
    low_i, high_i = int(np.floor(low)), int(np.floor(high))
    dx, dy, mag = sobel_argmax(img)
    return hysteresis(nms_labels(dx, dy, mag, low_i, high_i))


def generate_canny_np(img: np.ndarray, low_threshold, high_threshold, image_resolution: int) -> np.ndarray:
    """all_utils/utils.py:87-109 for inputs whose resize is the identity -> u8 [H,W,3]."""
    x = hwc3(np.asarray(img).astype(np.uint8))
    H, W, _ = resized_shape(x.shape[0], x.shape[1], image_resolution)
    if (H, W) != x.shape[:2]:
        raise ValueError("oracle covers the identity-resize case only (synthetic inputs are already HxW % 64 == 0)")
    return hwc3(canny(x, low_threshold, high_threshold))
