"""ctypes loader for the plain-C oracle (oracle/liboracle.so).  TEST INFRASTRUCTURE."""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build(force: bool = False) -> str:
    so = os.path.join(_HERE, "liboracle.so")
    srcs = [os.path.join(_HERE, f) for f in ("canny_ref.c", "pil_resize_ref.c")]
    if force or not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        subprocess.check_call(["make", "-C", _HERE, "-B", "liboracle.so"], stdout=subprocess.DEVNULL)
    return so


def lib():
    global _LIB
    if _LIB is None:
        _LIB = ctypes.CDLL(build())
    return _LIB


def canny(img: np.ndarray, low: int, high: int) -> np.ndarray:
    """img u8 [H,W,C] or [N,H,W,C] -> u8 [H,W] / [N,H,W]."""
    img = np.ascontiguousarray(img, dtype=np.uint8)
    single = img.ndim == 3
    if single:
        img = img[None]
    n, h, w, c = img.shape
    out = np.empty((n, h, w), np.uint8)
    rc = lib().oracle_canny_u8_batch(
        img.ctypes.data_as(ctypes.c_void_p), n, h, w, c, int(low), int(high), out.ctypes.data_as(ctypes.c_void_p)
    )
    if rc:
        raise MemoryError("oracle_canny_u8_batch failed")
    return out[0] if single else out


def pil_resize(img: np.ndarray, out_h: int, out_w: int, filt: str) -> np.ndarray:
    """img u8 [H,W,C] -> u8 [out_h,out_w,C], bit-exact PIL Image.resize (bilinear|bicubic, antialias)."""
    img = np.ascontiguousarray(img, dtype=np.uint8)
    h, w, c = img.shape
    out = np.empty((out_h, out_w, c), np.uint8)
    rc = lib().oracle_pil_resize_u8(
        img.ctypes.data_as(ctypes.c_void_p), h, w, c, out.ctypes.data_as(ctypes.c_void_p), out_h, out_w,
        {"bilinear": 0, "bicubic": 1}[filt],
    )
    if rc:
        raise RuntimeError("oracle_pil_resize_u8 failed")
    return out
