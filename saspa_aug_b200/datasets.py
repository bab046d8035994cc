"""Dataset-utils interface the hot path consumes: host-side mirror of the reference's ``all_utils/dataset_utils.py`` ``BaseUtils``
(:28-177) and its ``DS_UTILS_DICT`` registry (:547-554).  Same constructor (``split``, ``root_path``, ``print_func``), same attribute
and method names (``original_images_paths``, ``get_classes``, ``num_classes`` property, ``get_image_path_to_class_{id,str}_dict``,
``get_image_stem_to_class_str_dict``, ``get_basic_prompt``, ``get_image_path_with_same_class``, ``get_transform``,
``load_baseline_model``, ``get_baseline_conf_threshold``), so a subclass written against the reference registers here unchanged.

What differs is what ``load_baseline_model`` / ``get_transform`` return: the B200 classifier (``filter_nets.WSDANClassifier``, same
``.pth`` rules -- exactly one checkpoint, ``_orig_mod.`` keys, ResNet-101 first then ResNet-50) and a transform DESCRIPTION the batched
device-side pre-processing follows, instead of a torch module and a torchvision pipeline.

Real-dataset parsing (FGVC-Aircraft, Cars, DTD, CompCars, CUB) is out of scope (no datasets exist offline); a
synthetic dataset with the same interface drives benchmarks and tests."""
from __future__ import annotations

import os
from pathlib import Path
from typing import Callable, Dict, List, Optional


DATASETS_SUPPORTED = ["planes", "cars", "dtd", "compcars-parts", "cub", "planes_biased", "synthetic"]


class BaseUtils:
    def __init__(self, split="train", root_path: str = "", print_func: Callable = print):
        self.name: str = ""
        self.meta_class: str = ""
        self.root_path = Path(root_path)
        self.split = split
        self.print_func = print_func
        self.original_images_paths: List[str] = []
        self.baseline_model_cp = ""

    # ---- the reference's abstract surface (dataset_utils.py:38-65) -------------------------------------------------------
    def get_classes(self) -> List[str]:
        "return a list of all classes in the dataset, in string format"
        raise NotImplementedError

    @property
    def num_classes(self) -> int:
        return len(self.get_classes())

    def get_image_path_to_class_str_dict(self) -> Dict[str, str]:
        raise NotImplementedError

    def get_image_stem_to_class_str_dict(self) -> Dict[str, str]:
        raise NotImplementedError

    def get_image_path_to_class_id_dict(self, split="train") -> Dict[str, int]:
        raise NotImplementedError

    def get_basic_prompt(self) -> str:
        """used for semantic filtering"""
        raise NotImplementedError

    def get_image_path_with_same_class(self, image_path: str) -> List[str]:
        """dataset_utils.py:67-76: every image of the input image's class (BLIP-Diffusion subject images, run_aug.py:446)."""
        if self.name in ["planes", "cars"]:
            image_path = Path(image_path).stem
        class_str = self.image_path_to_class_str_dict[image_path]
        same = [path for path, cls in self.image_path_to_class_str_dict.items() if cls == class_str]
        if self.name in ["planes", "cars"]:
            same = [str(self.images_path / f"{path}.jpg") for path in same]
        return same

    # ---- filter models (dataset_utils.py:78-115) -----------------------------------------------------------------------------
    def get_transform(self, resize=(224, 224)) -> dict:
        """The reference returns torchvision's Resize(resize / 0.875) -> CenterCrop(resize) -> ToTensor -> Normalize(ImageNet)
        (:78-85).  Here the same recipe as data: the filter applies it on the device, batched and bit-exact for the u8 part
        (saspa_resize_pil_u8 + saspa_crop_normalize_bf16)."""
        return {"resize": (int(resize[0] / 0.875), int(resize[1] / 0.875)), "center_crop": tuple(resize), "interpolation": "bilinear",
                "mean": (0.485, 0.456, 0.406), "std": (0.229, 0.224, 0.225)}

    def checkpoint_folder(self) -> Path:
        """all_utils/checkpoints/<name> in the reference tree (:88-89); $SASPA_CHECKPOINTS overrides the root."""
        name = "compcars" if "compcars" in self.name else self.name
        return Path(os.environ.get("SASPA_CHECKPOINTS", Path(__file__).parent / "checkpoints")) / name

    def load_baseline_model(self, resize=(224, 224), device="cuda"):
        """dataset_utils.py:87-115: exactly one ``*.pth`` under the dataset's checkpoint folder, ``checkpoint['state_dict']``,
        ``_orig_mod.`` prefixes of torch.compile'd training runs, ResNet-101 tried first and ResNet-50 on a mismatch
        -> (WSDANClassifier on the device, transform description)."""
        from . import checkpoint_io

        cps = list(self.checkpoint_folder().glob("*.pth"))
        assert len(cps) == 1, f"Found {len(cps)} checkpoints in {self.checkpoint_folder()}. Expected 1"
        self.baseline_model_cp = cps[0]
        model = checkpoint_io.load_wsdan_checkpoint(self.baseline_model_cp, len(self.get_classes()), device)
        return model, self.get_transform(resize=resize)

    def get_baseline_conf_threshold(self) -> Dict[str, float]:
        """{str(class id): logit threshold} of the ALIA confidence filter (dataset_utils.py:117-146: mean logit of the correct class
        over the training images, cached as alia_confidence_thresholds/<name>.json)."""
        import json

        json_path = Path(f"alia_confidence_thresholds/{self.name}.json")
        if json_path.exists():
            return json.load(open(json_path, "r"))
        raise NotImplementedError(f"{json_path} not found: compute it with the baseline classifier over the training images")

    # ---- used only by the optional per-class CLIP filter (all_utils/utils.py:272-303, :383-389) ---------------------------------
    clip_filtering_suffix = ""  # ", a type of aircraft" (planes), ", a type of car" (cars), ", a type of a bird" (cub), ...

    def get_clip_filtering_prompts(self):
        return [f"a photo of a {name}{self.clip_filtering_suffix}." for name in self.get_classes()]

    def get_class_index_for_image(self, image_path: str) -> int:
        """Index in get_classes() of the source's class (the reference: image stem / path -> class string -> classnames.index)."""
        return self.get_classes().index(self.get_image_path_to_class_str_dict()[image_path])


class SyntheticUtils(BaseUtils):
    """N synthetic sources written under ``root`` as PNGs, label_i = i % num_classes, random-init filter nets
    of the reference architectures (no checkpoints exist offline)."""

    def __init__(self, split="train", root_path: Optional[str] = None, print_func: Callable = print, root: Optional[str] = None, n_images: int = 16,
                 seed_base: int = 0, size=(512, 512), wsdan_seed: int = 4242, clip_seed: int = 777, net: str = "resnet50", clip_model: str = "RN50",
                 n_classes: int = 100, sizes=None, names: Optional[List[str]] = None, labels: Optional[List[int]] = None):
        root = root or root_path or os.environ.get("SASPA_SYNTHETIC_ROOT", "/tmp/saspa_synthetic")
        super().__init__(split, root, print_func)
        self.name = "synthetic"
        self.meta_class = "airplane"
        self.root = self.root_path
        self.n_images, self.seed_base, self.size = n_images, seed_base, size
        self.sizes = sizes  # optional list of (h, w) cycled over the sources (mixed aspect ratios, like real FGVC data)
        self.wsdan_seed, self.clip_seed, self.net = wsdan_seed, clip_seed, net
        self.n_classes = n_classes
        self.alia_threshold = 1.0  # logit threshold of the synthetic ALIA confidence filter
        self.clip_model = clip_model  # "RN50" (the reference, all_utils/utils.py:253) | "ViT-L/14" (BASELINE config 5)
        self.lpips_seed = 31
        self.images_path = self.root / "images"
        names = names or [f"syn_{seed_base + i:07d}.png" for i in range(n_images)]
        self.original_images_paths = [str(self.images_path / n) for n in names]
        self._labels = labels
        self.image_path_to_class_str_dict = self.get_image_stem_to_class_str_dict()

    def size_of(self, i: int):
        return self.sizes[i % len(self.sizes)] if self.sizes else self.size

    def materialize(self):
        from PIL import Image

        from .synthetic import synthetic_source

        self.images_path.mkdir(parents=True, exist_ok=True)
        for i, p in enumerate(self.original_images_paths):
            if not os.path.exists(p):
                Image.fromarray(synthetic_source(self.seed_base + i, *self.size_of(i))).save(p)
        return self

    def label_of(self, i: int) -> int:
        return int(self._labels[i]) if self._labels is not None else i % self.n_classes

    def get_classes(self):
        return [f"class_{i}" for i in range(self.n_classes)]

    def get_image_path_to_class_id_dict(self, split="train"):
        return {p: self.label_of(i) for i, p in enumerate(self.original_images_paths)}

    def get_image_path_to_class_str_dict(self):
        return {p: f"class_{self.label_of(i)}" for i, p in enumerate(self.original_images_paths)}

    def get_image_stem_to_class_str_dict(self):
        return {Path(p).stem: f"class_{self.label_of(i)}" for i, p in enumerate(self.original_images_paths)}

    def get_image_path_with_same_class(self, image_path: str):
        d = self.get_image_path_to_class_str_dict()
        return [p for p, c in d.items() if c == d[image_path]]

    def get_basic_prompt(self):
        return f"a photo of an {self.meta_class}"

    def get_baseline_conf_threshold(self):
        return {str(c): self.alia_threshold for c in range(self.n_classes)}

    def load_filter_models(self, ds_utils, device):
        """Random-init WSDAN_CAL + CLIP (RN50 | ViT-L/14) of the reference architectures."""
        from . import checkpoints as ck
        from .filter_nets import CLIPRN50, CLIPViT, WSDANClassifier
        from .pipelines import SyntheticTokenizer

        wsd = ck.random_filter_state_dict(ck.wsdan_shapes(self.n_classes, self.net), self.wsdan_seed)
        if self.clip_model == "RN50":
            clip = CLIPRN50(ck.random_filter_state_dict(ck.clip_rn50_shapes(), self.clip_seed), device)
        elif self.clip_model == "ViT-L/14":
            clip = CLIPViT(ck.random_filter_state_dict(ck.clip_vit_shapes(), self.clip_seed), device)
        else:
            raise ValueError(f"clip_model {self.clip_model!r}: RN50 and ViT-L/14 are built")
        return WSDANClassifier(wsd, self.n_classes, self.net, device), clip, SyntheticTokenizer()

    def load_lpips(self, device):
        """Random-init LPIPS-AlexNet (lpips key layout) for the optional lpips_min / lpips_max filter."""
        from . import checkpoints as ck
        from .filter_nets import LPIPSAlex

        return LPIPSAlex(ck.random_lpips_state_dict(self.lpips_seed), device)


DS_UTILS_DICT: Dict[str, Callable] = {"synthetic": SyntheticUtils}
