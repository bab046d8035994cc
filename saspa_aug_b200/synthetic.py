"""Deterministic synthetic inputs: no datasets, weights or tokenizer
vocabularies exist offline, so benchmarks and parity tests use these generators.
Only the generators are committed, never the data."""
from __future__ import annotations

import numpy as np


def synthetic_source(seed: int, h: int = 512, w: int = 512, kind: str = "blobs") -> np.ndarray:
    """u8 [h,w,3] source image.  ``blobs`` (default) = smooth low-frequency field +
    filled circles + light blur (2-3 % Canny edge pixels at 120/200); ``noise`` =
    i.i.d. bytes (Canny worst case); ``smooth`` = the low-frequency field only."""
    import cv2

    rng = np.random.default_rng(seed)
    if kind == "noise":
        return rng.integers(0, 256, size=(h, w, 3), dtype=np.uint8)
    low = rng.integers(0, 256, size=(max(h // 16, 2), max(w // 16, 2), 3), dtype=np.uint8)
    img = cv2.resize(low, (w, h), interpolation=cv2.INTER_CUBIC)
    if kind == "smooth":
        return np.ascontiguousarray(img)
    for _ in range(40):
        cx, cy = int(rng.integers(0, w)), int(rng.integers(0, h))
        r = int(rng.integers(6, max(min(h, w) // 8, 7)))
        color = tuple(int(v) for v in rng.integers(0, 256, size=3))
        cv2.circle(img, (cx, cy), r, color, thickness=-1)
    img = cv2.GaussianBlur(img, (5, 5), 1.2)
    return np.ascontiguousarray(img)


def synthetic_token_ids(seed: int, batch: int = 1, vocab: int = 49408, length: int = 77):
    """CLIP-style ids: BOS 49406 at 0, random words, EOS 49407 at L in [8,40], EOS padding."""
    import torch

    g = torch.Generator().manual_seed(seed)
    ids = torch.randint(0, vocab - 2, (batch, length), generator=g)
    ids[:, 0] = vocab - 2
    for b in range(batch):
        L = int(torch.randint(8, 41, (1,), generator=g))
        ids[b, L:] = vocab - 1
    return ids
