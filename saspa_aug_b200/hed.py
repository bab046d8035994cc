"""HED conditioning for the ``CONTROLNET == "hed"`` flavour of the generation path: drop-in for ``controlnet_aux.HEDdetector`` as the
reference uses it (run_aug/run_aug.py:20 import, :311-312 ``HEDdetector.from_pretrained('lllyasviel/ControlNet')``, :438-439
``control_image = hed_detector(orig_img)``; the matching ControlNet checkpoint is ``lllyasviel/sd-controlnet-hed``, :66).

Same call surface (``from_pretrained(path_or_id, filename=None, cache_dir=None)``, ``.to(device)``, ``__call__(input_image,
detect_resolution=512, image_resolution=512, safe=False, output_type="pil", scribble=False)``) and the same state-dict layout
(``ControlNetHED.pth``: ``norm``, ``block{k}.convs.{i}``, ``block{k}.projection``), computed on the B200: the VGG-shaped trunk on the
tcgen05 implicit-GEMM convolutions with the ReLU in their epilogue (the 3-channel stem on the small-channel kernel), 2x2 max pooling on
saspa_pool2d_nhwc_bf16, the 1x1 projections with fp32 output, and the detector's resize / mean / sigmoid / quantise tail as ONE kernel
(saspa_hed_fuse_u8, csrc/hed.cu).  ``detect_batch`` is the batched entry the sharded driver uses (all sources of a micro-batch in one
pass; the reference runs the detector per prompt on one image).  controlnet_aux is un-vendored and absent offline: oracle/hed.py
restates it, parity unpinned."""
from __future__ import annotations

from pathlib import Path
from typing import Dict, List, Optional

import numpy as np
import torch

from . import nn, ops
from .checkpoints import HED_BLOCKS
from .ops import ACT_RELU

HED_FILENAMES = ("ControlNetHED.pth", "annotator/ckpts/ControlNetHED.pth")


def resize_image(input_image: np.ndarray, resolution: int) -> np.ndarray:
    """controlnet_aux.util.resize_image (host-side; min side -> resolution, sides rounded to multiples of 64, no area cap)."""
    import cv2

    H, W, _ = input_image.shape
    k = float(resolution) / min(float(H), float(W))
    H2 = int(np.round(float(H) * k / 64.0)) * 64
    W2 = int(np.round(float(W) * k / 64.0)) * 64
    if (H2, W2) == (H, W):
        return input_image
    return cv2.resize(input_image, (W2, H2), interpolation=cv2.INTER_LANCZOS4 if k > 1 else cv2.INTER_AREA)


class HEDNetwork:
    """ControlNetHED_Apache2 on the device: u8 RGB [n,H,W,3] -> five fp32 side outputs [n,H/2^k,W/2^k,1]."""

    def __init__(self, sd: Dict[str, torch.Tensor], device):
        self.device = torch.device(device)
        self.norm = tuple(float(v) for v in sd["norm"].reshape(3).float())
        self.blocks: List[List[nn.Conv]] = []
        self.proj: List[nn.Conv] = []
        for k, (_, _, layers) in enumerate(HED_BLOCKS, 1):
            self.blocks.append([nn.Conv(sd, f"block{k}.convs.{i}", self.device) for i in range(layers)])
            self.proj.append(nn.Conv(sd, f"block{k}.projection", self.device))

    def __call__(self, img_u8: torch.Tensor) -> List[torch.Tensor]:
        n, H, W, c = img_u8.shape
        assert img_u8.dtype == torch.uint8 and c == 3 and img_u8.is_cuda
        assert H % 16 == 0 and W % 16 == 0, "HED runs on detect_resolution-resized images (sides are multiples of 64)"
        # x - norm on the 0..255 scale: (x / 255 - norm / 255) / (1 / 255)
        h = ops.crop_normalize(img_u8.contiguous(), 0, 0, H, W, tuple(v / 255.0 for v in self.norm), (1.0 / 255.0,) * 3, out_c=3)
        sides = []
        for k, convs in enumerate(self.blocks):
            if k > 0:
                h = ops.pool2d(h, 2, 2, 0, True)
            for conv in convs:
                h = conv(h, act=ACT_RELU)
            sides.append(self.proj[k](h, out_fp32=True))
        return sides


class HEDdetector:
    def __init__(self, netNetwork: HEDNetwork):
        self.netNetwork = netNetwork

    @classmethod
    def from_state_dict(cls, sd: Dict[str, torch.Tensor], device="cuda") -> "HEDdetector":
        return cls(HEDNetwork(sd, device))

    @classmethod
    def from_pretrained(cls, pretrained_model_or_path, filename: Optional[str] = None, cache_dir=None, device="cuda") -> "HEDdetector":
        """A directory, or an HF id resolved on local storage ($SASPA_MODEL_ROOT / the hub cache: there is no network on the box)."""
        from . import checkpoint_io as cio

        names = (filename,) if filename else HED_FILENAMES
        ids = [str(pretrained_model_or_path)]
        if str(pretrained_model_or_path) == "lllyasviel/ControlNet":
            ids.append("lllyasviel/Annotators")  # where ControlNetHED.pth is published
        tried = []
        for mid in ids:
            root = Path(cache_dir) / mid if cache_dir and (Path(cache_dir) / mid).is_dir() else cio.resolve_model_dir(mid)
            for nm in names:
                tried.append(f"{root or mid}/{nm}")
                if root is not None and (Path(root) / nm).is_file():
                    sd = torch.load(Path(root) / nm, map_location="cpu", weights_only=True)
                    return cls.from_state_dict(sd, device)
        raise FileNotFoundError(f"HED checkpoint not on local storage (tried {', '.join(tried)})")

    def to(self, device):
        if torch.device(device) != self.netNetwork.device:
            raise NotImplementedError("build the detector on its device (from_pretrained(..., device=...))")
        return self

    def detect_batch(self, img_u8: torch.Tensor, safe: bool = False) -> torch.Tensor:
        """u8 RGB [n,H,W,3] on the device, already at the detect resolution -> u8 control maps [n,H,W,3]."""
        _, H, W, _ = img_u8.shape
        return ops.hed_fuse(self.netNetwork(img_u8), H, W, safe=safe, out_channels=3)

    def __call__(self, input_image, detect_resolution=512, image_resolution=512, safe=False, output_type="pil", scribble=False, **kwargs):
        from PIL import Image

        from .run_aug import HWC3

        if scribble:
            raise NotImplementedError("scribble=True (nms + blur + threshold) is not on the reference's path (run_aug.py:439 uses the defaults)")
        if "return_pil" in kwargs:
            output_type = "pil" if kwargs["return_pil"] else "np"
        if not isinstance(input_image, np.ndarray):
            input_image = np.array(input_image, dtype=np.uint8)
        input_image = resize_image(HWC3(input_image), detect_resolution)
        t = torch.from_numpy(np.ascontiguousarray(input_image))[None].to(self.netNetwork.device)
        detected = self.detect_batch(t, safe=safe)[0].cpu().numpy()
        out_hw = resize_image(input_image, image_resolution).shape[:2]
        if out_hw != detected.shape[:2]:
            raise NotImplementedError("image_resolution != detect_resolution: the reference calls the detector with both at 512")
        return Image.fromarray(detected) if output_type == "pil" else detected
