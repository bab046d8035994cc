"""CPU restatement of ``cv2.resize(u8 HWC, (W, H), interpolation=cv2.INTER_AREA | cv2.INTER_LANCZOS4)`` -- the calls the reference's
``utils.resize_image`` (all_utils/utils.py:58-79) makes: INTER_AREA for every source that is at least 512 px on its short side (k <= 1),
INTER_LANCZOS4 for smaller ones (k > 1); controlnet_aux's ``resize_image`` likewise.

TEST INFRASTRUCTURE.  OpenCV is a third-party dependency of the reference (``opencv-python==4.8.0.74``, environment.yml:24; 4.13.0 is what
is installed here and on the GPU box); this file restates the three code paths ``modules/imgproc/src/resize.cpp`` takes for 8-bit
INTER_AREA, from the published algorithm, and is PINNED: tests/test_resize_area_cpu.py compares it bit for bit with the installed
``cv2.resize`` itself on every path.

  * both scale factors >= 1 and integers  -> ResizeAreaFast: integer block sums; 2 x 2 blocks ``(s + 2) >> 2``, otherwise
    ``cvRound(float(s) * float(1 / area))``;
  * both >= 1, not both integers         -> ResizeArea: per-axis tables of (source index, float weight) from ``computeResizeAreaTab``,
    ``buf[dx] += S[sx] * alpha`` over the row, ``sum[dx] (+)= beta * buf[dx]`` over the rows, ``cvRound`` at the end (float32, no FMA);
  * one axis < 1 (the x64 rounding of resize_image can turn ONE axis into a slight up-scale) -> the fixed-point bilinear kernels with
    INTER_AREA's coefficient rule: ``sx = floor(dx * scale)``, ``fx = (dx + 1) - (sx + 1) / scale`` clipped to [0, 1), 11-bit weights,
    ``((b0 * (r0 >> 4)) >> 16) + ((b1 * (r1 >> 4)) >> 16) + 2) >> 2``."""
from __future__ import annotations

import math

import numpy as np


def area_tab(ssize: int, dsize: int):
    """computeResizeAreaTab: list of (dst index, src index, float32 weight)."""
    scale = ssize / dsize
    tab = []
    for dx in range(dsize):
        fsx1 = dx * scale
        fsx2 = fsx1 + scale
        cell = min(scale, ssize - fsx1)
        sx1, sx2 = math.ceil(fsx1), math.floor(fsx2)
        sx2 = min(sx2, ssize - 1)
        sx1 = min(sx1, sx2)
        if sx1 - fsx1 > 1e-3:
            tab.append((dx, sx1 - 1, np.float32((sx1 - fsx1) / cell)))
        for sx in range(sx1, sx2):
            tab.append((dx, sx, np.float32(1.0 / cell)))
        if fsx2 - sx2 > 1e-3:
            tab.append((dx, sx2, np.float32(min(min(fsx2 - sx2, 1.0), cell) / cell)))
    return tab


def _area_float(src, dw, dh):
    sh, sw, cn = src.shape
    xtab, ytab = area_tab(sw, dw), area_tab(sh, dh)
    dst = np.zeros((dh, dw, cn), np.uint8)
    srcf = src.astype(np.float32)
    buf = np.zeros((dw, cn), np.float32)
    acc = np.zeros((dw, cn), np.float32)
    prev = ytab[0][0]
    for dy, sy, beta in ytab:
        buf[:] = 0
        for dx, sx, alpha in xtab:
            buf[dx] = buf[dx] + srcf[sy, sx] * alpha
        if dy != prev:
            dst[prev] = np.clip(np.rint(acc), 0, 255).astype(np.uint8)
            acc = (beta * buf).astype(np.float32)
            prev = dy
        else:
            acc = (acc + beta * buf).astype(np.float32)
    dst[prev] = np.clip(np.rint(acc), 0, 255).astype(np.uint8)
    return dst


def _area_fast(src, dw, dh):
    sh, sw, cn = src.shape
    ix, iy = sw // dw, sh // dh
    blk = src.astype(np.int32)[: dh * iy, : dw * ix].reshape(dh, iy, dw, ix, cn).sum(axis=(1, 3))
    if ix == 2 and iy == 2:
        return ((blk + 2) >> 2).astype(np.uint8)
    v = blk.astype(np.float32) * np.float32(1.0 / (ix * iy))
    return np.clip(np.rint(v), 0, 255).astype(np.uint8)


def linear_area_tab(ssize: int, dsize: int):
    """The bilinear tables resize() builds when interpolation == INTER_AREA reaches the general path: (offsets, 11-bit weights, dmax)."""
    scale, inv = ssize / dsize, dsize / ssize
    ofs = np.zeros(dsize, np.int64)
    w = np.zeros((dsize, 2), np.int64)
    dmax = dsize
    for d in range(dsize):
        s = math.floor(d * scale)
        f = np.float32((d + 1) - (s + 1) * inv)
        f = np.float32(0.0) if f <= 0 else np.float32(f - math.floor(f))
        if s < 0:
            f, s = np.float32(0.0), 0
        if s + 1 >= ssize:
            dmax = min(dmax, d)
            if s >= ssize - 1:
                f, s = np.float32(0.0), ssize - 1
        ofs[d] = s
        w[d, 0] = int(np.clip(np.rint(np.float32((np.float32(1.0) - f) * np.float32(2048))), -32768, 32767))
        w[d, 1] = int(np.clip(np.rint(np.float32(f * np.float32(2048))), -32768, 32767))
    return ofs, w, dmax


def _area_linear(src, dw, dh):
    sh, sw, cn = src.shape
    xofs, xw, xmax = linear_area_tab(sw, dw)
    yofs, yw, _ = linear_area_tab(sh, dh)
    S = src.astype(np.int64)
    H = np.zeros((sh, dw, cn), np.int64)
    for d in range(dw):
        s = xofs[d]
        H[:, d] = S[:, s] * xw[d, 0] + S[:, s + 1] * xw[d, 1] if d < xmax else S[:, s] * 2048
    out = np.zeros((dh, dw, cn), np.uint8)
    for d in range(dh):
        s0, s1 = min(max(yofs[d], 0), sh - 1), min(max(yofs[d] + 1, 0), sh - 1)
        v = (((int(yw[d, 0]) * (H[s0] >> 4)) >> 16) + ((int(yw[d, 1]) * (H[s1] >> 4)) >> 16) + 2) >> 2
        out[d] = np.clip(v, 0, 255).astype(np.uint8)
    return out


def resize_area(src: np.ndarray, dw: int, dh: int) -> np.ndarray:
    """cv2.resize(src, (dw, dh), interpolation=cv2.INTER_AREA) for uint8 HWC."""
    assert src.dtype == np.uint8 and src.ndim == 3
    sh, sw, _ = src.shape
    if (sh, sw) == (dh, dw):
        return src.copy()
    sx, sy = sw / dw, sh / dh
    if sx >= 1 and sy >= 1:
        if sx == int(sx) and sy == int(sy):
            return _area_fast(src, dw, dh)
        return _area_float(src, dw, dh)
    return _area_linear(src, dw, dh)


_S45 = 0.70710678118654752440084436210485
_CS = ((1, 0), (-_S45, -_S45), (0, 1), (_S45, -_S45), (-1, 0), (_S45, _S45), (0, -1), (-_S45, _S45))


def lanczos4_coeffs(x) -> np.ndarray:
    """interpolateLanczos4: eight float32 weights for the fractional position x in [0, 1)."""
    x = np.float32(x)
    co = np.zeros(8, np.float32)
    x3 = np.float32(x + np.float32(3))
    y0 = -float(x3) * math.pi * 0.25
    s0, c0 = math.sin(y0), math.cos(y0)
    sm = np.float32(0)
    for i in range(8):
        y0_ = np.float32(x3 - np.float32(i))
        if abs(y0_) >= np.float32(1e-6):
            y = -float(y0_) * math.pi * 0.25
            co[i] = np.float32((_CS[i][0] * s0 + _CS[i][1] * c0) / (y * y))
        else:
            co[i] = np.float32(1e30)
        sm = np.float32(sm + co[i])
    return (co * np.float32(np.float32(1.0) / sm)).astype(np.float32)


def lanczos4_tab(ssize: int, dsize: int):
    """(floor of the source coordinate, 11-bit weights of taps ofs - 3 .. ofs + 4) per destination index."""
    scale = ssize / dsize
    ofs = np.zeros(dsize, np.int64)
    w = np.zeros((dsize, 8), np.int64)
    for d in range(dsize):
        f = np.float32((d + 0.5) * scale - 0.5)
        s = math.floor(f)
        ofs[d] = s
        w[d] = np.clip(np.rint(lanczos4_coeffs(np.float32(f - np.float32(s))) * np.float32(2048)), -32768, 32767).astype(np.int64)
    return ofs, w


def resize_lanczos4(src: np.ndarray, dw: int, dh: int) -> np.ndarray:
    """cv2.resize(src, (dw, dh), interpolation=cv2.INTER_LANCZOS4) for uint8 HWC: HResizeLanczos4 (int sums, replicated border) then
    VResizeLanczos4 with FixedPtCast<int, uchar, 22>."""
    assert src.dtype == np.uint8 and src.ndim == 3
    sh, sw, cn = src.shape
    if (sh, sw) == (dh, dw):
        return src.copy()
    xofs, xw = lanczos4_tab(sw, dw)
    yofs, yw = lanczos4_tab(sh, dh)
    S = src.astype(np.int64)
    H = np.zeros((sh, dw, cn), np.int64)
    for d in range(dw):
        H[:, d] = sum(S[:, min(max(xofs[d] - 3 + k, 0), sw - 1)] * xw[d, k] for k in range(8))
    out = np.zeros((dh, dw, cn), np.uint8)
    for d in range(dh):
        acc = sum(H[min(max(yofs[d] - 3 + k, 0), sh - 1)] * yw[d, k] for k in range(8))
        out[d] = np.clip((acc + (1 << 21)) >> 22, 0, 255).astype(np.uint8)
    return out
