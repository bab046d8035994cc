"""CPU fp32 restatement of the HED conditioning detector the reference offers as the second ControlNet flavour:
``HEDdetector.from_pretrained('lllyasviel/ControlNet')`` at run_aug/run_aug.py:311-312 and ``control_image = hed_detector(orig_img)``
at :438-439 (CONTROLNET == "hed", ControlNet checkpoint ``lllyasviel/sd-controlnet-hed`` :66).

TEST INFRASTRUCTURE.  The arithmetic lives in the un-vendored dependency ``controlnet-aux==0.0.5`` (environment.yml:23; not installed
offline, no wheel, no network) and the reference holds no test or golden image for it => **parity unpinned**: this file restates the
published detector of the ControlNet annotators (``ControlNetHED_Apache2``: a 13-convolution VGG-shaped trunk, ReLU after every
convolution, 2x2 max pooling in front of blocks 2-5, a 1x1 projection to one channel per block; state-dict layout of
``ControlNetHED.pth``: ``norm``, ``block{1..5}.convs.{i}.{weight,bias}``, ``block{1..5}.projection.{weight,bias}``) and the numpy / OpenCV
steps of ``HEDdetector.__call__`` that follow it.  Random-init weights stand in for the checkpoint."""
from __future__ import annotations

import cv2
import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

BLOCKS = ((3, 64, 2), (64, 128, 2), (128, 256, 3), (256, 512, 3), (512, 512, 3))  # (in, out, convolutions) of block1..block5


class DoubleConvBlock(nn.Module):
    def __init__(self, cin, cout, layers):
        super().__init__()
        self.convs = nn.Sequential(*[nn.Conv2d(cin if i == 0 else cout, cout, 3, 1, 1) for i in range(layers)])
        self.projection = nn.Conv2d(cout, 1, 1, 1, 0)

    def forward(self, x, down_sampling=False):
        h = F.max_pool2d(x, 2, 2) if down_sampling else x
        for conv in self.convs:
            h = F.relu(conv(h))
        return h, self.projection(h)


class ControlNetHED(nn.Module):
    def __init__(self):
        super().__init__()
        self.norm = nn.Parameter(torch.zeros(1, 3, 1, 1))
        for k, (cin, cout, layers) in enumerate(BLOCKS, 1):
            setattr(self, f"block{k}", DoubleConvBlock(cin, cout, layers))

    def forward(self, x):
        h = x - self.norm
        sides = []
        for k in range(1, 6):
            h, p = getattr(self, f"block{k}")(h, down_sampling=k > 1)
            sides.append(p)
        return sides


def HWC3(x):
    """controlnet_aux.util.HWC3 (same function as the reference's all_utils/utils.py:39-55)."""
    assert x.dtype == np.uint8
    if x.ndim == 2:
        x = x[:, :, None]
    H, W, C = x.shape
    assert C in (1, 3, 4)
    if C == 3:
        return x
    if C == 1:
        return np.concatenate([x, x, x], axis=2)
    color = x[:, :, 0:3].astype(np.float32)
    alpha = x[:, :, 3:4].astype(np.float32) / 255.0
    return (color * alpha + 255.0 * (1.0 - alpha)).clip(0, 255).astype(np.uint8)


def resize_image(input_image, resolution):
    """controlnet_aux.util.resize_image: min side -> resolution, both sides rounded to multiples of 64 (no area cap, unlike the reference's
    own all_utils/utils.py:58-79)."""
    H, W, C = input_image.shape
    k = float(resolution) / min(float(H), float(W))
    H = int(np.round(float(H) * k / 64.0)) * 64
    W = int(np.round(float(W) * k / 64.0)) * 64
    return cv2.resize(input_image, (W, H), interpolation=cv2.INTER_LANCZOS4 if k > 1 else cv2.INTER_AREA)


def safe_step(x, step=2):
    y = x.astype(np.float32) * float(step + 1)
    return y.astype(np.int32).astype(np.float32) / float(step)


def fuse_sides(sides, H, W, safe=False):
    """The numpy / OpenCV tail of HEDdetector.__call__: five float32 maps [h_k, w_k] -> uint8 [H, W]."""
    edges = [cv2.resize(np.ascontiguousarray(e, dtype=np.float32), (W, H), interpolation=cv2.INTER_LINEAR) for e in sides]
    edges = np.stack(edges, axis=2)
    edge = 1 / (1 + np.exp(-np.mean(edges, axis=2).astype(np.float64)))
    if safe:
        edge = safe_step(edge)
    return (edge * 255.0).clip(0, 255).astype(np.uint8)


@torch.no_grad()
def hed_detect(net: ControlNetHED, input_image: np.ndarray, detect_resolution=512, image_resolution=512, safe=False, return_sides=False):
    """HEDdetector.__call__(input_image, detect_resolution=512, image_resolution=512, safe=False, output_type="np") -> uint8 [H, W, 3]."""
    input_image = resize_image(HWC3(np.asarray(input_image, dtype=np.uint8)), detect_resolution)
    H, W, _ = input_image.shape
    x = torch.from_numpy(input_image.copy()).float().permute(2, 0, 1)[None]
    sides = [e.numpy().astype(np.float32)[0, 0] for e in net(x)]
    detected = HWC3(fuse_sides(sides, H, W, safe))
    img = resize_image(input_image, image_resolution)
    H2, W2, _ = img.shape
    detected = cv2.resize(detected, (W2, H2), interpolation=cv2.INTER_LINEAR)
    return (detected, sides) if return_sides else detected
