"""Dataset-utils interface the hot path consumes (reference all_utils/dataset_utils.py:28-177 ``BaseUtils`` and the
``DS_UTILS_DICT`` registry :547-554): ``original_images_paths``, ``get_image_path_to_class_id_dict``,
``get_basic_prompt``, ``num_classes``, ``meta_class``, plus model loading for the filter.  Real-dataset parsing
(FGVC-Aircraft, Cars, DTD, CompCars, CUB) is out of scope (no datasets exist offline); a
synthetic dataset with the same interface drives benchmarks and tests, and real datasets can register here."""
from __future__ import annotations

import os
from pathlib import Path
from typing import Callable, Dict, List

import numpy as np


class BaseUtils:
    name = "base"
    meta_class = "object"
    num_classes = 0

    def __init__(self, print_func: Callable = print):
        self.print_func = print_func
        self.original_images_paths: List[str] = []

    def get_classes(self) -> List[str]:
        raise NotImplementedError

    def get_image_path_to_class_id_dict(self) -> Dict[str, int]:
        raise NotImplementedError

    def get_basic_prompt(self) -> str:
        """used for semantic filtering (dataset_utils.py:62-64)"""
        return f"a photo of a {self.meta_class}"

    def load_filter_models(self, ds_utils, device):
        raise NotImplementedError

    # ---- used only by the optional filters (all_utils/utils.py:269-300, :323-328) -----------------------------------------
    clip_filtering_suffix = ""  # e.g. ", a type of aircraft" (planes), ", a type of car" (cars), ", a type of a bird" (cub)

    def get_clip_filtering_prompts(self):
        """all_utils/utils.py:276-296: one prompt per class name."""
        return [f"a photo of a {name}{self.clip_filtering_suffix}." for name in self.get_classes()]

    def get_class_index_for_image(self, image_path: str) -> int:
        """Index of the source image's class in get_classes() (the reference goes through image-stem / image-path -> class-string
        dicts, :383-388)."""
        return int(self.get_image_path_to_class_id_dict()[image_path])

    def get_baseline_conf_threshold(self):
        """{str(class id): logit threshold} of the ALIA confidence filter (dataset_utils get_baseline_conf_threshold)."""
        raise NotImplementedError


class SyntheticUtils(BaseUtils):
    """N synthetic 512x512 sources written under ``root`` as PNGs, label_i = i % num_classes."""
    name = "synthetic"
    meta_class = "airplane"
    num_classes = 100

    def __init__(self, print_func: Callable = print, root: str = None, n_images: int = 16, seed_base: int = 0, size=(512, 512),
                 wsdan_seed: int = 4242, clip_seed: int = 777, net: str = "resnet50", clip_model: str = "RN50"):
        super().__init__(print_func)
        self.root = Path(root or os.environ.get("SASPA_SYNTHETIC_ROOT", "/tmp/saspa_synthetic"))
        self.n_images, self.seed_base, self.size = n_images, seed_base, size
        self.wsdan_seed, self.clip_seed, self.net = wsdan_seed, clip_seed, net
        self.alia_threshold = 1.0  # logit threshold of the synthetic ALIA confidence filter
        self.clip_model = clip_model  # "RN50" (the reference, all_utils/utils.py:253) | "ViT-L/14" (BASELINE config 5)
        self.images_path = self.root / "images"
        self.original_images_paths = [str(self.images_path / f"syn_{seed_base + i:07d}.png") for i in range(n_images)]

    def materialize(self):
        from PIL import Image

        from .synthetic import synthetic_source

        self.images_path.mkdir(parents=True, exist_ok=True)
        for i, p in enumerate(self.original_images_paths):
            if not os.path.exists(p):
                Image.fromarray(synthetic_source(self.seed_base + i, *self.size)).save(p)
        return self

    def get_classes(self):
        return [f"class_{i}" for i in range(self.num_classes)]

    def get_image_path_to_class_id_dict(self):
        return {p: i % self.num_classes for i, p in enumerate(self.original_images_paths)}

    def get_baseline_conf_threshold(self):
        return {str(c): self.alia_threshold for c in range(self.num_classes)}

    def load_filter_models(self, ds_utils, device):
        """Random-init WSDAN_CAL + CLIP (RN50 | ViT-L/14) of the reference architectures (no checkpoints offline)."""
        from . import checkpoints as ck
        from .filter_nets import CLIPRN50, CLIPViT, WSDANClassifier
        from .pipelines import SyntheticTokenizer

        wsd = ck.random_filter_state_dict(ck.wsdan_shapes(self.num_classes, self.net), self.wsdan_seed)
        if self.clip_model == "RN50":
            clip = CLIPRN50(ck.random_filter_state_dict(ck.clip_rn50_shapes(), self.clip_seed), device)
        elif self.clip_model == "ViT-L/14":
            clip = CLIPViT(ck.random_filter_state_dict(ck.clip_vit_shapes(), self.clip_seed), device)
        else:
            raise ValueError(f"clip_model {self.clip_model!r}: RN50 and ViT-L/14 are built")
        return WSDANClassifier(wsd, self.num_classes, self.net, device), clip, SyntheticTokenizer()


DS_UTILS_DICT: Dict[str, Callable] = {"synthetic": SyntheticUtils}
