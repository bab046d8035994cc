"""CPU fp32 restatement of diffusers 0.32.2 ``StableDiffusionSafetyChecker`` (pipelines/stable_diffusion/safety_checker.py) and of the
``run_safety_checker`` step of the SD v1.5 pipelines the reference builds without ``safety_checker=None`` (run_aug/run_aug.py:200-211),
plus the ``CLIPImageProcessor`` preprocessing that feeds it (shortest side -> 224 PIL bicubic, center crop, /255, CLIP mean/std).

TEST INFRASTRUCTURE.  **Parity unpinned** at the reference boundary (diffusers is an un-vendored dependency, environment.yml:17, and the
reference holds no test for it); the vision tower is the installed ``transformers.CLIPVisionModel`` itself, the decision loop follows the
published source line by line (including ``round(., 3)`` and the 0.01 special-care adjustment)."""
from __future__ import annotations

from typing import List

import numpy as np
import torch
import torch.nn as nn

CLIP_MEAN = (0.48145466, 0.4578275, 0.40821073)
CLIP_STD = (0.26862954, 0.26130258, 0.27577711)


def clip_image_processor(images_u8: np.ndarray, size: int = 224) -> torch.Tensor:
    """[n,H,W,3] u8 -> [n,3,size,size] fp32."""
    from PIL import Image

    out = []
    for a in images_u8:
        h, w = a.shape[:2]
        oh, ow = (size, int(size * w / h)) if h <= w else (int(size * h / w), size)
        r = np.asarray(Image.fromarray(a).resize((ow, oh), resample=Image.BICUBIC)) if (oh, ow) != (h, w) else a
        cy, cx = int(round((oh - size) / 2.0)), int(round((ow - size) / 2.0))
        x = r[cy:cy + size, cx:cx + size].astype(np.float32) / 255.0
        out.append((x - np.asarray(CLIP_MEAN, np.float32)) / np.asarray(CLIP_STD, np.float32))
    return torch.from_numpy(np.stack(out)).permute(0, 3, 1, 2).contiguous()


def cosine_distance(image_embeds, text_embeds):
    return torch.mm(nn.functional.normalize(image_embeds), nn.functional.normalize(text_embeds).t())


class SafetyCheckerOracle(nn.Module):
    def __init__(self, width=1024, layers=24, patch=14, res=224, proj=768, n_concepts=17, n_special=3):
        super().__init__()
        from transformers import CLIPVisionConfig, CLIPVisionModel

        cfg = CLIPVisionConfig(hidden_size=width, intermediate_size=4 * width, num_hidden_layers=layers, num_attention_heads=max(1, width // 64),
                               image_size=res, patch_size=patch, hidden_act="quick_gelu", projection_dim=proj)
        self.res = res
        self.vision_model = CLIPVisionModel(cfg)
        self.visual_projection = nn.Linear(width, proj, bias=False)
        self.concept_embeds = nn.Parameter(torch.ones(n_concepts, proj), requires_grad=False)
        self.special_care_embeds = nn.Parameter(torch.ones(n_special, proj), requires_grad=False)
        self.concept_embeds_weights = nn.Parameter(torch.ones(n_concepts), requires_grad=False)
        self.special_care_embeds_weights = nn.Parameter(torch.ones(n_special), requires_grad=False)

    @torch.no_grad()
    def cosines(self, clip_input):
        pooled = self.vision_model(clip_input)[1]
        emb = self.visual_projection(pooled)
        return cosine_distance(emb, self.special_care_embeds), cosine_distance(emb, self.concept_embeds)

    @torch.no_grad()
    def forward(self, clip_input, images: np.ndarray):
        special_cos_dist, cos_dist = (t.cpu().float().numpy() for t in self.cosines(clip_input))
        result = []
        for i in range(clip_input.shape[0]):
            img = {"special_scores": {}, "special_care": [], "concept_scores": {}, "bad_concepts": []}
            adjustment = 0.0
            for c in range(len(special_cos_dist[0])):
                img["special_scores"][c] = round(float(special_cos_dist[i][c]) - self.special_care_embeds_weights[c].item() + adjustment, 3)
                if img["special_scores"][c] > 0:
                    img["special_care"].append({c, img["special_scores"][c]})
                    adjustment = 0.01
            for c in range(len(cos_dist[0])):
                img["concept_scores"][c] = round(float(cos_dist[i][c]) - self.concept_embeds_weights[c].item() + adjustment, 3)
                if img["concept_scores"][c] > 0:
                    img["bad_concepts"].append(c)
            result.append(img)
        has_nsfw: List[bool] = [len(r["bad_concepts"]) > 0 for r in result]
        images = images.copy()
        for idx, flag in enumerate(has_nsfw):
            if flag:
                images[idx] = np.zeros(images[idx].shape, images.dtype)
        return images, has_nsfw, result
