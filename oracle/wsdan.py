"""CPU restatement of the eval-mode forward of the reference's baseline classifier
``WSDAN_CAL`` (fgvc/models/cal.py:131-213; BAP :44-86; trunk = ResNet.get_features fgvc/models/resnet.py:168-178,
Bottleneck :61-104, layer4 stride 1 :118-119; attention head BasicConv2d fgvc/models/inception.py:374-384)
plus the filter rule ``correct_label in logits.topk(k)[1]`` (all_utils/utils.py:357-365).

TEST INFRASTRUCTURE.  Parity PINNED: tests/test_filter_oracle_cpu.py checks this restatement against the
reference's own WSDAN_CAL imported from /root/reference (same state dict, same input) and against the
committed golden logits (tests/golden/wsdan_golden.npz, made by tests/golden/make_filter_golden.py).
Parameter names follow the reference state dict so its checkpoints load unchanged."""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F

EPSILON = 1e-6
LAYERS = {"resnet50": (3, 4, 6, 3), "resnet101": (3, 4, 23, 3)}


class Bottleneck(nn.Module):
    def __init__(self, inplanes, planes, stride, downsample):
        super().__init__()
        self.conv1 = nn.Conv2d(inplanes, planes, 1, bias=False)
        self.bn1 = nn.BatchNorm2d(planes)
        self.conv2 = nn.Conv2d(planes, planes, 3, stride=stride, padding=1, bias=False)
        self.bn2 = nn.BatchNorm2d(planes)
        self.conv3 = nn.Conv2d(planes, planes * 4, 1, bias=False)
        self.bn3 = nn.BatchNorm2d(planes * 4)
        self.downsample = downsample

    def forward(self, x):
        o = F.relu(self.bn1(self.conv1(x)))
        o = F.relu(self.bn2(self.conv2(o)))
        o = self.bn3(self.conv3(o))
        r = x if self.downsample is None else self.downsample(x)
        return F.relu(o + r)


class AttentionHead(nn.Module):
    def __init__(self, cin, m):
        super().__init__()
        self.conv = nn.Conv2d(cin, m, 1, bias=False)
        self.bn = nn.BatchNorm2d(m, eps=0.001)

    def forward(self, x):
        return F.relu(self.bn(self.conv(x)))


class WSDANOracle(nn.Module):
    def __init__(self, num_classes: int, net: str = "resnet50", M: int = 32):
        super().__init__()
        mods = [nn.Conv2d(3, 64, 7, stride=2, padding=3, bias=False), nn.BatchNorm2d(64), nn.ReLU(), nn.MaxPool2d(3, 2, 1)]
        inplanes = 64
        for li, (planes, blocks) in enumerate(zip((64, 128, 256, 512), LAYERS[net])):
            stride = 1 if li in (0, 3) else 2
            layer = []
            for b in range(blocks):
                s = stride if b == 0 else 1
                ds = None
                if b == 0 and (s != 1 or inplanes != planes * 4):
                    ds = nn.Sequential(nn.Conv2d(inplanes, planes * 4, 1, stride=s, bias=False), nn.BatchNorm2d(planes * 4))
                layer.append(Bottleneck(inplanes, planes, s, ds))
                inplanes = planes * 4
            mods.append(nn.Sequential(*layer))
        self.features = nn.Sequential(*mods)
        self.attentions = AttentionHead(2048, M)
        self.fc = nn.Linear(M * 2048, num_classes, bias=False)

    def forward(self, x):
        f = self.features(x)
        a = self.attentions(f)
        b, c, h, w = f.shape
        fm = (torch.einsum("imjk,injk->imn", a, f) / float(h * w)).view(b, -1)
        fm = torch.sign(fm) * torch.sqrt(torch.abs(fm) + EPSILON)
        fm = F.normalize(fm, dim=-1)
        return self.fc(fm * 100.0)


def in_topk(logits: torch.Tensor, labels, k: int):
    """all_utils/utils.py:363: `correct_label in logits.topk(conf_top_k)[1]`."""
    top = logits.topk(k, dim=-1)[1]
    return torch.tensor([int(int(l) in top[i].tolist()) for i, l in enumerate(labels)], dtype=torch.uint8)
