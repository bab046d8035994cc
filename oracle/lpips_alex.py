"""CPU fp32 restatement of the LPIPS distance the reference's optional filter uses: ``lpips.LPIPS(net='alex')`` at
all_utils/utils.py:269-270 and ``calc_lpips_distance`` at :576-590 (images -> "L" -> "RGB", PIL resize to 256x256, ToTensor, x*2-1).

TEST INFRASTRUCTURE.  The arithmetic lives in the un-vendored dependency ``lpips`` (richzhang/PerceptualSimilarity, pip package lpips
0.1.4; not installed offline, the reference itself wraps the import in try/except at all_utils/utils.py:20-23) => **parity unpinned**:
this file restates the published algorithm -- ScalingLayer, torchvision AlexNet ``features`` sliced at the five ReLUs, unit-normalised
channel vectors, squared difference, non-negative 1x1 ``lin`` weights, spatial mean, sum over layers -- and random-init weights stand in
for the ImageNet AlexNet + the learned linear heads."""
from __future__ import annotations

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

SHIFT = (-0.030, -0.088, -0.188)
SCALE = (0.458, 0.448, 0.450)
CHANNELS = (64, 192, 384, 256, 256)


class LPIPSAlex(nn.Module):
    def __init__(self):
        super().__init__()
        # torchvision.models.alexnet().features indices: conv 0, 3, 6, 8, 10 (ReLU after each, MaxPool(3, 2) after the first two)
        self.convs = nn.ModuleList([nn.Conv2d(3, 64, 11, 4, 2), nn.Conv2d(64, 192, 5, 1, 2), nn.Conv2d(192, 384, 3, 1, 1), nn.Conv2d(384, 256, 3, 1, 1),
                                    nn.Conv2d(256, 256, 3, 1, 1)])
        self.lins = nn.ParameterList([nn.Parameter(torch.rand(1, c, 1, 1) / c) for c in CHANNELS])  # lin{k}.model.1.weight, >= 0
        self.register_buffer("shift", torch.tensor(SHIFT).view(1, 3, 1, 1))
        self.register_buffer("scale", torch.tensor(SCALE).view(1, 3, 1, 1))

    def features(self, x):
        out = []
        h = (x - self.shift) / self.scale
        for k, conv in enumerate(self.convs):
            if k in (1, 2):
                h = F.max_pool2d(h, 3, 2)
            h = F.relu(conv(h))
            out.append(h)
        return out

    @torch.no_grad()
    def forward(self, in0, in1):
        """in0, in1 fp32 [n,3,H,W] in [-1,1] -> distance [n]."""
        total = 0.0
        for f0, f1, w in zip(self.features(in0), self.features(in1), self.lins):
            n0 = f0 / (f0.pow(2).sum(dim=1, keepdim=True).sqrt() + 1e-10)
            n1 = f1 / (f1.pow(2).sum(dim=1, keepdim=True).sqrt() + 1e-10)
            total = total + ((n0 - n1).pow(2) * w).sum(dim=1).mean(dim=(1, 2))
        return total


def state_dict_keys():
    """(name, shape) of the lpips 0.1.4 state dict restricted to what the distance reads: net.slice{k}.{idx}.{weight,bias} + lin{k}.model.1.weight."""
    idx = (0, 3, 6, 8, 10)
    shapes = [(64, 3, 11, 11), (192, 64, 5, 5), (384, 192, 3, 3), (256, 384, 3, 3), (256, 256, 3, 3)]
    out = []
    for k, (i, s) in enumerate(zip(idx, shapes)):
        out += [(f"net.slice{k + 1}.{i}.weight", s), (f"net.slice{k + 1}.{i}.bias", (s[0],))]
    out += [(f"lin{k}.model.1.weight", (1, c, 1, 1)) for k, c in enumerate(CHANNELS)]
    return out


def random_state_dict(seed: int):
    g = torch.Generator().manual_seed(seed)
    sd = {}
    for name, shape in state_dict_keys():
        if name.startswith("lin"):
            sd[name] = torch.rand(shape, generator=g) / shape[1]
        elif name.endswith("bias"):
            sd[name] = 0.05 * torch.randn(shape, generator=g)
        else:
            fan_in = shape[1] * shape[2] * shape[3]
            sd[name] = torch.randn(shape, generator=g) * (2.0 / fan_in) ** 0.5
    return sd


def load(model: LPIPSAlex, sd):
    idx = (0, 3, 6, 8, 10)
    for k, conv in enumerate(model.convs):
        conv.weight.data.copy_(sd[f"net.slice{k + 1}.{idx[k]}.weight"])
        conv.bias.data.copy_(sd[f"net.slice{k + 1}.{idx[k]}.bias"])
    for k, p in enumerate(model.lins):
        p.data.copy_(sd[f"lin{k}.model.1.weight"])
    return model.eval()


def preprocess(img_u8: np.ndarray, resize=(256, 256)) -> torch.Tensor:
    """calc_lpips_distance's image path (all_utils/utils.py:577-586): PIL "L" -> "RGB" -> resize (PIL default: bicubic) -> ToTensor -> *2-1."""
    from PIL import Image

    im = Image.fromarray(img_u8).convert("L").convert("RGB")
    if resize:
        im = im.resize(resize)
    x = torch.from_numpy(np.asarray(im).astype(np.float32) / 255.0).permute(2, 0, 1)
    return x * 2 - 1
