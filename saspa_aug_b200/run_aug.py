"""Sharded generation driver: host-side mirror of the reference's ``run_aug/run_aug.py`` for the ControlNet (canny | hed) path.

Kept from the reference (names / defaults / behaviour): the module-level constants as ``AugConfig`` fields
(run_aug.py:513-577), ``init_pipeline`` (:128-230), ``pass_thorugh_pipe`` (:233-279), the output-folder and file naming
(:668-692, :377, :429, :442), prompt post-processing (:342-346, :380-394), skip-if-exists resume (:430-432), the error
policy (RuntimeError logged, loop ends; :493-500) and the final filter + JSON call (:721-733).

B200-first differences: one process per GPU; sources are partitioned ``index % world == rank`` with NO collective in the
loop; Canny runs once per source on the GPU (the reference recomputes it per prompt on the CPU, :436-437); prompts of a
micro-batch are generated together; PNG encoding runs in a thread pool; the only collective is one gather of the
per-image filter records before rank 0 writes the JSON.
"""
from __future__ import annotations

import logging
import os
import random
from concurrent.futures import ThreadPoolExecutor
from dataclasses import dataclass
from pathlib import Path
from typing import List, Optional, Sequence

import numpy as np

NEGATIVE_PROMPT = ("over-exposure, under-exposure, saturated, duplicate, out of frame, lowres, cropped, worst quality, low quality, jpeg artifacts, morbid, "
                   "mutilated, out of frame, ugly, bad anatomy, bad proportions, deformed, blurry, duplicate")  # the run constant of run_aug.py:47
MAX_FILENAME_LENGTH = 40
MAX_PROMPT_LENGTH = 150


@dataclass
class AugConfig:
    """The reference's module-level constants (run_aug.py:513-556) with their shipped defaults, except DATASET (the reference ships
    "planes"; no dataset exists offline, so the synthetic stand-in is the default) and the fields below the rule that are this
    driver's own (RNG_MODE, MICRO_BATCH)."""
    DATASET: str = "synthetic"
    BASE_MODEL: str = "sd_v1.5"
    CONTROLNET: Optional[str] = "canny"
    SDEDIT: int = 0
    NUM_PER_IMAGE: int = 2
    SEED: int = 1
    RESOLUTION: int = 512
    GUIDANCE_SCALE: float = 7.5
    NUM_INFERENCE_STEPS: int = 30
    SDEDIT_STRENGTH: float = 0.85
    LOW_THRESHOLD_CANNY: int = 120
    HIGH_THRESHOLD_CANNY: int = 200
    CONTROLNET_CONDITIONING_SCALE: float = 0.75
    PROMPT_TYPE: str = "gpt-meta_class"
    PROMPT_WITH_SUB_CLASS: bool = True
    USE_ARTISTIC_PROMPTS: Optional[bool] = None  # None = the reference's rule: True iff BASE_MODEL == "sd_v1.5" (run_aug.py:530)
    ARTISTIC_PROMPTS_PROB: float = 0.5
    USE_CAMERA_VARIATIONS_PROMPTS: bool = False
    CAMERA_VAIRATIONS_PROB: float = 0.5  # (sic, run_aug.py:533)
    STYLE_IMG_FROM_DIFF_IMG: bool = True  # BLIP-Diffusion: the subject image is another image of the source's class (run_aug.py:548)
    NEGATIVE_PROMPT: Optional[str] = NEGATIVE_PROMPT
    SEMANTIC_FILTERING: int = 1
    MODEL_CONFIDENCE_BASED_FILTERING: int = 1
    CONF_TOP_K: int = 10
    SAMPLER: str = "ddim"
    RNG_MODE: str = "per_item"  # "per_item" (partition independent) | "reference_order" (replays the global generator)
    MICRO_BATCH: int = 16
    DEVICE_RESIZE: bool = False  # loader threads resize the sources on the GPU (bit-exact INTER_AREA / INTER_LANCZOS4) instead of with cv2

    def __post_init__(self):
        if self.USE_ARTISTIC_PROMPTS is None:
            self.USE_ARTISTIC_PROMPTS = self.BASE_MODEL in ("sd_v1.5", "tiny")

    def apply_dataset_rules(self):
        """run_aug.py:560-577, in the reference's order."""
        if "cars" in self.DATASET.lower():
            self.NUM_INFERENCE_STEPS = 50
        if self.DATASET.lower() == "cub":
            self.BASE_MODEL = "sd_xl-turbo"
        if self.BASE_MODEL == "sd_xl-turbo":
            self.GUIDANCE_SCALE, self.NUM_INFERENCE_STEPS, self.NEGATIVE_PROMPT = 0.0, 2, None
        if self.SDEDIT:
            assert self.NUM_INFERENCE_STEPS * self.SDEDIT_STRENGTH >= 1, "NUM_INFERENCE_STEPS * SDEDIT_STRENGTH must be >= 1"
        return self


def output_folder(ds_root: str, cfg: AugConfig) -> str:
    """run_aug.py:668-692."""
    prompt_str = cfg.PROMPT_TYPE
    if cfg.PROMPT_WITH_SUB_CLASS:
        prompt_str += "_prompt_w_sub_class"
    if cfg.USE_ARTISTIC_PROMPTS:
        prompt_str += f"_artistic_prompts_p_{cfg.ARTISTIC_PROMPTS_PROB}"
    if cfg.USE_CAMERA_VARIATIONS_PROMPTS:
        prompt_str += f"_camera_variations_p_{cfg.CAMERA_VAIRATIONS_PROB}"
    if "blip_diffusion" in cfg.BASE_MODEL and cfg.STYLE_IMG_FROM_DIFF_IMG:
        prompt_str += "_style_img_from_diff_img"
    base = f"regular/{cfg.BASE_MODEL}"
    if cfg.SDEDIT:
        base += f"-SDEdit_strength_{cfg.SDEDIT_STRENGTH}"
    if cfg.CONTROLNET:
        base = base.replace("regular/", "controlnet/")
    return f"{ds_root}/aug_data/{base}/{cfg.CONTROLNET}/{prompt_str}_seed_{cfg.SEED}/images"


def aug_file_name(image_stem: str, prompt: str, i: int) -> str:
    """run_aug.py:429."""
    return f"{image_stem[:MAX_FILENAME_LENGTH]}_prompt_{prompt.replace('/', '-')}_{i}.png"


def resized_hw(h: int, w: int, smaller_side_res: int):
    """Size rule of all_utils/utils.py:58-79: min side -> res, area cap 1.2 MP, both sides rounded to multiples of 64 -> (H, W, k)."""
    MAX_RES_SIZE = 1200000
    H, W = float(h), float(w)
    k = float(smaller_side_res) / min(H, W)
    H *= k
    W *= k
    if H * W > MAX_RES_SIZE:
        k = np.sqrt(MAX_RES_SIZE / (H * W))
        H *= k
        W *= k
    return int(np.round(H / 64.0)) * 64, int(np.round(W / 64.0)) * 64, k


def set_seed(seed: int) -> None:
    """all_utils/utils.py:32-36: the reference seeds ``random``, ``np.random`` and torch once per run (run_aug.py:588).  The sharded driver
    does not depend on this global state (``replay_prompt_draws`` / ``reference_order_noise`` re-create the streams from ``cfg.SEED`` on
    every rank); the function exists for callers that drive ``pass_thorugh_pipe`` themselves, as the reference's loop does."""
    import random

    import torch

    random.seed(seed)
    np.random.seed(seed)
    torch.manual_seed(seed)
    if torch.cuda.is_available():
        torch.cuda.manual_seed_all(seed)


def resize_image(input_image: np.ndarray, smaller_side_res: int) -> np.ndarray:
    """all_utils/utils.py:58-79 (host-side pre-processing; identity for sources already at HxW % 64 == 0, min side = res)."""
    import cv2

    H, W, k = resized_hw(input_image.shape[0], input_image.shape[1], smaller_side_res)
    if (H, W) == input_image.shape[:2]:
        return input_image
    return cv2.resize(input_image, (W, H), interpolation=cv2.INTER_LANCZOS4 if k > 1 else cv2.INTER_AREA)


def resize_image_device(img_u8, smaller_side_res: int):
    """``resize_image`` for a u8 HWC image that is already on the device: same size rule, same pixels as cv2 -- INTER_AREA for k <= 1
    (saspa_resize_area_u8), INTER_LANCZOS4 for k > 1 (saspa_resize_lanczos4_u8), both bit-exact against the installed OpenCV."""
    from . import ops

    H, W, k = resized_hw(int(img_u8.shape[0]), int(img_u8.shape[1]), smaller_side_res)
    if (H, W) == tuple(img_u8.shape[:2]):
        return img_u8
    return (ops.resize_lanczos4 if k > 1 else ops.resize_area)(img_u8.contiguous(), H, W)


def HWC3(x: np.ndarray) -> np.ndarray:
    """all_utils/utils.py:39-55."""
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


def generate_canny(cond_image_input, low_threshold, high_threshold, image_resolution):
    """Drop-in for all_utils/utils.py:102-109: PIL | ndarray -> PIL RGB edge map (0/255), computed by saspa_canny_u8."""
    import torch
    from PIL import Image

    from . import ops

    img = resize_image(HWC3(np.array(cond_image_input).astype(np.uint8)), image_resolution)
    t = torch.from_numpy(np.ascontiguousarray(img)[None]).cuda()
    edges, _ = ops.canny(t, int(low_threshold), int(high_threshold), out_channels=3)
    return Image.fromarray(edges[0].cpu().numpy())


BASE_MODEL_DICT = {  # run_aug.py:53-62 (the entries whose ControlNet-canny pipelines are built here)
    "sd_v1.5": "runwayml/stable-diffusion-v1-5",
    "sd_xl": "stabilityai/stable-diffusion-xl-base-1.0",
    "sd_xl-turbo": "stabilityai/sdxl-turbo",
    "blip_diffusion": "Salesforce/blipdiffusion-controlnet",  # :181 -- the ControlNet flavour has its own repo, ControlNet included
}
CONTROLNET_DICT_SD = {"canny": "lllyasviel/control_v11p_sd15_canny", "hed": "lllyasviel/sd-controlnet-hed"}  # :64-67
HED_ANNOTATOR = "lllyasviel/ControlNet"                                       # :312 HEDdetector.from_pretrained(...)
CONTROLNET_DICT_SD_XL = {"canny": "diffusers/controlnet-canny-sdxl-1.0"}    # :69-72
SDXL_VAE = "madebyollin/sdxl-vae-fp16-fix"                                    # :189


def _local_checkpoints(base_model: str, controlnet: str, model_dirs: Optional[dict]):
    """The local directories of the HF repos the reference's init_pipeline downloads (run_aug.py:141,:184,:189), or None when the base
    model is not on local storage (then the caller falls back to random init -- there is no network here)."""
    from . import checkpoint_io as cio

    md = dict(model_dirs or {})
    base = md.get("base") or (cio.resolve_model_dir(BASE_MODEL_DICT[base_model]) if base_model in BASE_MODEL_DICT else None)
    if base is None:
        return None
    xl = base_model in ("sd_xl", "sd_xl-turbo")
    cn = md.get("controlnet")
    if cn is None and base_model != "blip_diffusion":
        cn = cio.resolve_model_dir((CONTROLNET_DICT_SD_XL if xl else CONTROLNET_DICT_SD)[controlnet])
        if cn is None:
            raise FileNotFoundError(f"{base_model} is on local storage ({base}) but its ControlNet "
                                    f"{(CONTROLNET_DICT_SD_XL if xl else CONTROLNET_DICT_SD)[controlnet]} is not")
    vae = md.get("vae") or (cio.resolve_model_dir(SDXL_VAE) if xl else None)
    return cio.load_pipeline_dir(base, cn, vae)


def init_pipeline(base_model, controlnet, SDEdit, use_compile=False, sampler="ddim", state_dicts=None, device="cuda", model_dirs=None):
    """run_aug.py:128-230 for the ControlNet-canny pipelines (SD v1.5, SD-XL(-turbo), BLIP-Diffusion).  ``use_compile`` is accepted
    and ignored: nothing is traced, the kernels are launched directly.
    Weights, in this order: ``state_dicts`` (diffusers-keyed dicts); the HF repos of run_aug.py:53-72 found on LOCAL storage
    (``model_dirs`` = {"base", "controlnet", "vae"} paths, else $SASPA_MODEL_ROOT / the HF hub cache: safetensors or .bin weights, the
    components' config.json, the CLIP BPE tokenizer files, scheduler_config.json); else deterministic random init of the named
    architecture (logged) -- no checkpoint exists in the build environment."""
    from .pipelines import (SaspaBlipControlNetPipeline, SaspaControlNetPipeline, SaspaSDXLControlNetPipeline, blip_configs, random_state_dicts,
                            sdxl_configs)

    assert sampler in ["ddim", "unipcmultistep"]
    assert controlnet in ("canny", "hed"), "ControlNet conditioning: canny | hed (run_aug.py:132)"
    if controlnet == "hed" and base_model not in ("sd_v1.5", "tiny"):
        raise KeyError(f"no HED ControlNet for {base_model!r} (run_aug.py:69-72 lists canny only for SD-XL; BLIP-Diffusion ships its canny ControlNet)")
    loaded = _local_checkpoints(base_model, controlnet, model_dirs) if state_dicts is None else None
    cfgs = loaded["configs"] if loaded else {}
    toks = {k: loaded.get(k) for k in ("tokenizer", "tokenizer_2")} if loaded else {}
    if loaded is None and state_dicts is None and base_model in BASE_MODEL_DICT:
        logging.warning(f"init_pipeline({base_model!r}): no local checkpoint found, using deterministic RANDOM weights of that architecture")
    if base_model in ("sd_xl-turbo", "sd_xl", "tiny_xl"):
        # run_aug.py:188-199 (+ :223-228: for sd_xl-turbo the scheduler is rebuilt from the turbo config => trailing spacing)
        cfg = "tiny_xl" if base_model == "tiny_xl" else "sdxl"
        sds = state_dicts or loaded or random_state_dicts(cfg, 1234)
        u, v, t1, t2 = sdxl_configs(cfg)
        u, v, t1, t2 = cfgs.get("unet", u), cfgs.get("vae", v), cfgs.get("text", t1), cfgs.get("text2", t2)
        turbo = base_model != "sd_xl"
        smp = ("unipc_sdxl_turbo" if sampler == "unipcmultistep" else "ddim_sdxl_turbo") if turbo else ("unipc" if sampler == "unipcmultistep" else "ddim")
        return SaspaSDXLControlNetPipeline.from_state_dicts(sds["unet"], sds["controlnet"], sds["vae"], sds["text"], sds["text2"], unet_cfg=u, vae_cfg=v,
                                                            text_cfg=t1, text2_cfg=t2, sampler=smp, device=device, img2img=bool(SDEdit),
                                                            tokenizer=toks.get("tokenizer"), tokenizer_2=toks.get("tokenizer_2"))
    if base_model in ("blip_diffusion", "tiny_blip"):
        # run_aug.py:185-187: BlipDiffusionControlNetPipeline; the sampler stays the checkpoint's PNDM (run_aug.py:217 skips the swap)
        cfg = "tiny_blip" if base_model == "tiny_blip" else "blip"
        sds = state_dicts or loaded or random_state_dicts(cfg, 1234)
        u, v, t, q = blip_configs(cfg)
        u, v, t = cfgs.get("unet", u), cfgs.get("vae", v), cfgs.get("text", t)
        return SaspaBlipControlNetPipeline.from_state_dicts(sds["unet"], sds["controlnet"], sds["vae"], sds["text"], sds["qformer"], unet_cfg=u, vae_cfg=v,
                                                            text_cfg=t, qformer_cfg=q, device=device, tokenizer=toks.get("tokenizer"))
    if base_model not in ("sd_v1.5", "tiny"):
        raise NotImplementedError(f"base_model {base_model!r}: SD v1.5, SD-XL(-turbo) and BLIP-Diffusion ControlNet paths are built (see DESIGN.md)")
    cfg = "tiny" if base_model == "tiny" else "sd15"
    sds = state_dicts or loaded or random_state_dicts(cfg, 1234)
    from . import checkpoints as ck

    kw = {}
    if cfg == "tiny":
        kw = dict(unet_cfg=ck.UNetConfig.tiny(), vae_cfg=ck.VAEConfig.tiny(), text_cfg=ck.CLIPTextConfig.tiny())
    if cfgs:
        kw = dict(unet_cfg=cfgs["unet"], vae_cfg=cfgs["vae"], text_cfg=cfgs["text"])
    smp = "unipc" if sampler == "unipcmultistep" else "ddim"
    pipe = SaspaControlNetPipeline.from_state_dicts(sds["unet"], sds["controlnet"], sds["vae"], sds["text"], sampler=smp, device=device, img2img=bool(SDEdit),
                                                    tokenizer=toks.get("tokenizer"), **kw)
    if controlnet == "hed":  # run_aug.py:311-312: the detector rides with the pipeline (generate() reads pipe.hed_detector)
        from .hed import HEDdetector

        if sds.get("hed") is not None:
            pipe.hed_detector = HEDdetector.from_state_dict(sds["hed"], device)
        elif loaded:
            pipe.hed_detector = HEDdetector.from_pretrained(HED_ANNOTATOR, device=device)
        else:
            pipe.hed_detector = HEDdetector.from_state_dict(ck.random_hed_state_dict(4321), device)
    return pipe


def pass_thorugh_pipe(base_model, pipe, prompt, orig_img, SDEdit, SDEdit_strength, num_inference_steps, generator, guidance_scale, control_cond_scale,
                      negative_prompt=NEGATIVE_PROMPT, control_image=None, blip_src_category=None, blip_target_category=None):
    """run_aug.py:233-279 (same kwargs assembly, same return)."""
    pipe_args = {"prompt": str(prompt), "num_inference_steps": num_inference_steps, "generator": generator, "guidance_scale": guidance_scale,
                 "negative_prompt": negative_prompt}
    blip = "blip" in base_model
    if blip:  # run_aug.py:243-250
        pipe_args.update(reference_image=orig_img, source_subject_category=blip_src_category, target_subject_category=blip_target_category,
                         height=orig_img.size[1], width=orig_img.size[0], neg_prompt=NEGATIVE_PROMPT)
        del pipe_args["negative_prompt"]
    if control_image is not None:
        if SDEdit:
            pipe_args["control_image"] = control_image
            pipe_args["controlnet_conditioning_scale"] = control_cond_scale
        elif blip:  # run_aug.py:268-271 (the parameter really is spelled "condtioning_image")
            pipe_args["condtioning_image"] = control_image
            pipe_args["height"] = control_image.size[1]
            pipe_args["width"] = control_image.size[0]
        else:
            pipe_args["image"] = control_image
            pipe_args["controlnet_conditioning_scale"] = control_cond_scale
    if SDEdit:
        pipe_args["image"] = orig_img
        pipe_args["strength"] = SDEdit_strength
    return pipe(**pipe_args).images[0]


def shard_indices(n: int, rank: int, world: int) -> List[int]:
    """rank r of W takes source indices {i : i % W == r}; both augmentations of a source stay on one rank."""
    return [i for i in range(n) if i % world == rank]


def item_seed(seed: int, index: int, i: int) -> int:
    """Per-item generator seed, independent of the partitioning (RNG_MODE = per_item)."""
    return (seed * 1_000_003 + index * 131 + i) % (2 ** 31 - 1)


def reference_order_noise(cfg: AugConfig, source_hw: Sequence, draws: Sequence[Sequence], mine, latent_channels: int = 4, dtype=None):
    """RNG_MODE = "reference_order" (SURVEY.md 8e): the reference threads ONE generator -- ``torch.manual_seed(SEED)``, the global CPU
    generator, run_aug.py:324 -- through every pipeline call in dataset order (:464), and diffusers draws from it on the CPU in the
    pipeline dtype (fp16 after ``.to(DEVICE, torch.float16)``, :323): per call the VAE-posterior noise first when SDEdit is on, then the
    latent noise.  This rank-independent pre-pass replays exactly that stream (a private generator with the same seed yields the same
    values) and keeps the draws of the items in ``mine``; items the reference would skip (existing output, :430-432) draw nothing.
      source_hw  [(H, W)] of every source AFTER resize_image;  draws: ``replay_prompt_draws`` rows (``.skipped`` per item, or plain
      strings = nothing skipped);  mine: set of source indices of this rank
    -> {(index, i): (noise fp32 [1,C,H/8,W/8], posterior noise fp32 | None)}"""
    import torch

    dtype = dtype or torch.float16
    g = torch.Generator().manual_seed(cfg.SEED)
    out = {}
    for index, (H, W) in enumerate(source_hw):
        shape = (1, latent_channels, H // 8, W // 8)
        for i, d in enumerate(draws[index]):
            if getattr(d, "skipped", False):
                continue
            post = torch.randn(shape, generator=g, dtype=dtype) if cfg.SDEDIT else None
            noise = torch.randn(shape, generator=g, dtype=dtype)
            if index in mine:
                out[(index, i)] = (noise.float(), post.float() if post is not None else None)
    return out


class PromptDraw:
    """What the reference's inner loop settles for one (source, i) before the pipeline call (run_aug.py:385-456)."""
    __slots__ = ("prompt", "subject_path", "skipped")

    def __init__(self, prompt: str, subject_path: Optional[str] = None, skipped: bool = False):
        self.prompt, self.subject_path, self.skipped = prompt, subject_path, skipped


def replay_prompt_draws(prompts, paths: Sequence[str], cfg: AugConfig, ds_utils=None, skip=None) -> List[List[PromptDraw]]:
    """Rank-independent pre-pass replaying, in dataset order, every draw the reference's loop makes from its two GLOBAL streams
    (``utils.set_seed(SEED)`` seeds both, run_aug.py:588): ``np.random`` -- ``choice(prompts, NUM_PER_IMAGE)`` per source (:382), the
    artistic / camera suffix (:394,:397) -- and ``random`` -- ``random.random()`` of the artistic test (evaluated only when its first
    clause is false: Python short-circuits ``(i % 2 == 0 and P == 0.5) or (random.random() < P and P != 0.5)``, :391), of the camera test
    (:396), and BLIP-Diffusion's ``random.choice`` of a same-class subject image (:446, drawn only for items that are NOT skipped).
    Then the prompt rewriting: compcars-parts prefix (:386-389), sub-class insertion per dataset (:399-427).

      prompts  list of prompt strings already truncated to MAX_PROMPT_LENGTH (:308,:345), or a callable (index, path) -> list for the
               per-image prompt types (captions / txt2sentence-per_class, :361-367)
      skip     (index, i, prompt) -> bool: output already exists (:430-432); decides whether the BLIP subject draw happens
    """
    from .prompts import ARTISTIC_PROMPTS, IMAGE_VARIATIONS_PROMPTS

    np_rs = np.random.RandomState(cfg.SEED)
    py_rs = random.Random(cfg.SEED)
    blip = "blip_diffusion" in cfg.BASE_MODEL or cfg.BASE_MODEL == "tiny_blip"
    stem_keyed = cfg.DATASET in ("planes", "cars", "planes_biased", "synthetic")  # run_aug.py:698
    classes = None
    if cfg.PROMPT_WITH_SUB_CLASS:
        if ds_utils is None:
            raise ValueError("PROMPT_WITH_SUB_CLASS needs ds_utils (image -> class string, run_aug.py:698)")
        classes = ds_utils.get_image_stem_to_class_str_dict() if stem_keyed else ds_utils.get_image_path_to_class_str_dict()
    fixed = None if callable(prompts) else list(prompts)
    out: List[List[PromptDraw]] = []
    for index, source_image_path in enumerate(paths):
        cur = list(prompts(index, source_image_path)) if fixed is None else fixed
        cur = [p[:-1] if p[-1] == "." else p for p in cur]  # :380 (re-applied per source on the same list, as the reference does)
        if fixed is not None:
            fixed = cur
        image_stem = Path(source_image_path).stem
        sampled = np_rs.choice(cur, cfg.NUM_PER_IMAGE)
        row = []
        for i, prompt in enumerate(sampled):
            prompt = str(prompt)
            if cfg.DATASET == "compcars-parts":
                part = source_image_path.split("/")[-2]
                prompt = f"{ds_utils.get_basic_prompt(part=part)} {prompt}"
            P = cfg.ARTISTIC_PROMPTS_PROB
            if cfg.USE_ARTISTIC_PROMPTS and ((i % 2 == 0 and P == 0.5) or (py_rs.random() < P and P != 0.5)):
                prompt = f"{prompt}, {np_rs.choice(ARTISTIC_PROMPTS)}"
            elif cfg.USE_CAMERA_VARIATIONS_PROMPTS and py_rs.random() < cfg.CAMERA_VAIRATIONS_PROB:
                prompt = f"{prompt}, {np_rs.choice(IMAGE_VARIATIONS_PROMPTS)} photo"
            if cfg.PROMPT_WITH_SUB_CLASS:
                if cfg.DATASET in ("planes", "planes_biased"):
                    prompt = prompt.replace("airplane", f"{classes[image_stem]} airplane")
                elif cfg.DATASET == "cars":
                    prompt = prompt.replace("car", f"{classes[image_stem]} car")
                elif cfg.DATASET == "dtd":
                    prompt = f"{prompt} with a {classes[source_image_path]} texture"
                elif cfg.DATASET in ("compcars", "compcars-parts"):
                    prompt = prompt.replace("car", f"{classes[source_image_path]} car")
                elif cfg.DATASET == "cub":
                    prompt = prompt.replace("bird", f"{classes[source_image_path]} bird")
                elif cfg.DATASET == "synthetic":  # the stand-in dataset follows the planes rule with its own meta class
                    prompt = prompt.replace(ds_utils.meta_class, f"{classes[image_stem]} {ds_utils.meta_class}")
                else:
                    raise NotImplementedError
            skipped = bool(skip(index, i, prompt)) if skip is not None else False
            subject = None
            if blip and cfg.STYLE_IMG_FROM_DIFF_IMG and not skipped:
                subject = py_rs.choice(ds_utils.get_image_path_with_same_class(source_image_path))
            row.append(PromptDraw(prompt, subject, skipped))
        out.append(row)
    return out


def sample_prompts(prompts, n_sources, cfg: AugConfig, ds_utils=None) -> List[List[str]]:
    """The prompt strings of ``replay_prompt_draws`` (``n_sources``: a count -- sources are then anonymous, which is only valid without
    PROMPT_WITH_SUB_CLASS -- or the list of source paths)."""
    if isinstance(n_sources, int):
        if cfg.PROMPT_WITH_SUB_CLASS or cfg.DATASET == "compcars-parts":
            raise ValueError("sample_prompts(count) cannot rewrite prompts per source: pass the source paths")
        paths = [f"src_{k}" for k in range(n_sources)]
    else:
        paths = list(n_sources)
    prompts = prompts if callable(prompts) else [p.strip()[:MAX_PROMPT_LENGTH] for p in prompts]
    c = cfg
    uses_py_stream = cfg.USE_CAMERA_VARIATIONS_PROMPTS or (cfg.USE_ARTISTIC_PROMPTS and cfg.ARTISTIC_PROMPTS_PROB != 0.5)
    if "blip" in cfg.BASE_MODEL and cfg.STYLE_IMG_FROM_DIFF_IMG and ds_utils is None and not uses_py_stream:
        # the subject draw shares the `random` stream with the suffix tests; when those never read it, it cannot change a prompt
        import copy

        c = copy.copy(cfg)
        c.STYLE_IMG_FROM_DIFF_IMG = False
    return [[d.prompt for d in row] for row in replay_prompt_draws(prompts, paths, c, ds_utils)]


@dataclass
class WorkItem:
    index: int            # source index in ds_utils.original_images_paths
    i: int                # augmentation index of that source
    prompt: str
    out_path: str
    subject_path: Optional[str] = None  # BLIP-Diffusion subject image (None = the source itself)


def source_hw(path: str, resolution: int):
    """(H, W) of a source after resize_image, from the image header only."""
    from PIL import Image

    with Image.open(path) as im:
        w, h = im.size
    return resized_hw(h, w, resolution)[:2]


def plan_work(cfg: AugConfig, paths: Sequence[str], draws, out_dir: str, mine: Sequence[int], hw_of):
    """Host-side plan of this rank's shard: which (source, i) still have to be generated (resume: existing outputs are kept,
    run_aug.py:430-432) and how they are cut into micro-batches.  FGVC sources are not square -- resize_image keeps the aspect ratio
    (512x704, 512x768, ...) -- and a launch list needs one latent shape, so work is bucketed by the resized (H, W) of its source and each
    bucket is cut into chunks of at most MICRO_BATCH: every micro-batch is uniform by construction, whatever was already on disk.
    -> (existing [(index, i, path)], chunks [[WorkItem]])"""
    existing, buckets = [], {}
    for index in mine:
        stem = Path(paths[index]).stem
        for i, d in enumerate(draws[index]):
            out_path = str(Path(out_dir) / aug_file_name(stem, d.prompt, i))
            if d.skipped or os.path.exists(out_path):
                logging.info(f"Skipping {out_path} as it already exists")
                existing.append((index, i, out_path))
                continue
            buckets.setdefault(tuple(hw_of(index)), []).append(WorkItem(index, i, d.prompt, out_path, d.subject_path))
    chunks = []
    for items in buckets.values():  # insertion order = first appearance in dataset order
        for b0 in range(0, len(items), cfg.MICRO_BATCH):
            chunks.append(items[b0 : b0 + cfg.MICRO_BATCH])
    return existing, chunks


def generate(cfg: AugConfig, ds_utils, pipe, prompts, out_dir: str, rank: int = 0, world: int = 1, io_threads: int = 8, prefetch: int = 2):
    """The generation loop (run_aug.py:357-471) for this rank's shard.  Returns the list of (source index, i, path).
    Works for the three ControlNet base models: SD v1.5 / SD-XL(-turbo) (text + added conditioning from the pipeline's own
    ``_encode_call``) and BLIP-Diffusion (subject image per run_aug.py:444-456, subject categories = ``ds_utils.meta_class``;
    conditioning scale 1.0, no SDEdit).

    Streaming: nothing but file names is held for the whole shard.  Sources of the next ``prefetch`` micro-batches are decoded and
    resized by loader threads while the GPU runs the current one; the u8 result leaves through a pinned buffer with an asynchronous
    copy whose completion a writer thread waits on before handing rows to the PNG pool, so the launching thread never blocks on D2H
    or on encoding.  Every save is checked: a failed one is logged and dropped from the returned list."""
    import queue
    import threading

    import torch
    from PIL import Image

    from . import ops

    Path(out_dir).mkdir(parents=True, exist_ok=True)
    paths = ds_utils.original_images_paths
    prompt_list = prompts if callable(prompts) else [p.strip()[:MAX_PROMPT_LENGTH] for p in prompts]
    mine = shard_indices(len(paths), rank, world)
    if cfg.RNG_MODE not in ("per_item", "reference_order"):
        raise ValueError(f"RNG_MODE {cfg.RNG_MODE!r}: per_item | reference_order")

    def exists(index, i, prompt):
        return (Path(out_dir) / aug_file_name(Path(paths[index]).stem, prompt, i)).exists()

    draws = replay_prompt_draws(prompt_list, paths, cfg, ds_utils, skip=exists)
    hw_cache = {}

    def hw_of(index):
        if index not in hw_cache:
            hw_cache[index] = source_hw(paths[index], cfg.RESOLUTION)
        return hw_cache[index]

    ref_noise = None
    if cfg.RNG_MODE == "reference_order":
        ref_noise = reference_order_noise(cfg, [hw_of(k) for k in range(len(paths))], draws, set(mine), pipe.vae_cfg.latent_channels)
    existing, chunks = plan_work(cfg, paths, draws, out_dir, mine, hw_of)
    written = list(existing)
    dev = pipe.device
    blip = "blip" in cfg.BASE_MODEL
    assert not (blip and cfg.SDEDIT), "BLIP-Diffusion has no img2img ControlNet pipeline in the reference (run_aug.py:185-187)"
    if cfg.CONTROLNET == "hed" and getattr(pipe, "hed_detector", None) is None:
        raise ValueError('CONTROLNET == "hed" needs pipe.hed_detector (init_pipeline(base_model, "hed", ...) attaches it)')
    do_cfg = cfg.GUIDANCE_SCALE > 1.0
    pool = ThreadPoolExecutor(max_workers=io_threads)
    loaders = ThreadPoolExecutor(max_workers=max(2, io_threads // 2))
    saves = []  # (future, (index, i, out_path) | None)

    def load(path):
        img = np.array(Image.open(path).convert("RGB"))
        if cfg.DEVICE_RESIZE:
            # same pixels as cv2 (saspa_resize_area_u8), on this loader thread's own stream so it never queues behind the denoising kernels
            side = torch.cuda.Stream(device=dev)
            with torch.cuda.stream(side):
                out = resize_image_device(torch.from_numpy(img).to(dev, non_blocking=True), cfg.RESOLUTION).cpu()
            side.synchronize()
            return out.numpy()
        return resize_image(img, cfg.RESOLUTION)

    def submit_loads(chunk):
        """-> {path: future} for every distinct image the chunk reads (sources, BLIP subjects)."""
        want = {paths[w.index] for w in chunk} | {w.subject_path for w in chunk if w.subject_path}
        return {p: loaders.submit(load, p) for p in sorted(want)}

    done_q: "queue.Queue" = queue.Queue()

    def writer():
        while True:
            item = done_q.get()
            if item is None:
                return
            ev, pinned, chunk = item
            ev.synchronize()
            arr = pinned.numpy()
            for w, a in zip(chunk, arr):
                saves.append((pool.submit(Image.fromarray(a.copy()).save, w.out_path), (w.index, w.i, w.out_path)))

    wt = threading.Thread(target=writer, daemon=True)
    wt.start()
    num_errors = 0
    saved_sources = set()
    try:
        pending = [submit_loads(c) for c in chunks[:prefetch]]
        for ci, chunk in enumerate(chunks):
            loaded = {p: f.result() for p, f in pending.pop(0).items()}
            if ci + prefetch < len(chunks):
                pending.append(submit_loads(chunks[ci + prefetch]))
            uniq = sorted({w.index for w in chunk})
            for u in uniq:  # "<stem>_source.png" (run_aug.py:377-378)
                src_out = os.path.join(out_dir, f"{Path(paths[u]).stem[:MAX_FILENAME_LENGTH]}_source.png")
                if u not in saved_sources and not os.path.exists(src_out):
                    saves.append((pool.submit(Image.fromarray(loaded[paths[u]]).save, src_out), None))
                saved_sources.add(u)
            src_t = torch.from_numpy(np.stack([loaded[paths[u]] for u in uniq])).to(dev)
            if cfg.CONTROLNET == "hed":  # run_aug.py:438-439, once per source instead of once per prompt
                edges = pipe.hed_detector.detect_batch(src_t)
                ctrl = ops.crop_normalize(edges, 0, 0, src_t.shape[1], src_t.shape[2], (0.0, 0.0, 0.0), (1.0, 1.0, 1.0), out_c=3)
            else:
                edges, ctrl = ops.canny(src_t, cfg.LOW_THRESHOLD_CANNY, cfg.HIGH_THRESHOLD_CANNY, out_channels=3, want_ctrl=True)
            sel = torch.tensor([uniq.index(w.index) for w in chunk], device=dev)
            for u_i, u in enumerate(uniq):
                if u < 10:  # first 10 control images are saved (run_aug.py:441-442)
                    cp = os.path.join(out_dir, f"{Path(paths[u]).stem[:MAX_FILENAME_LENGTH]}_control.png")
                    if not os.path.exists(cp):
                        saves.append((pool.submit(Image.fromarray(edges[u_i].cpu().numpy()).save, cp), None))
            H, W = src_t.shape[1:3]
            extra = {}
            if blip:
                subject = getattr(ds_utils, "meta_class", "object")
                if any(w.subject_path for w in chunk):  # another image of the source's class is the subject (run_aug.py:445-454)
                    subj = []
                    for w in chunk:
                        a = loaded[w.subject_path] if w.subject_path else loaded[paths[w.index]]
                        sp = os.path.join(out_dir, f"{Path(paths[w.index]).stem[:MAX_FILENAME_LENGTH]}_subject_{w.i}.png")
                        saves.append((pool.submit(Image.fromarray(a).save, sp), None))
                        subj.append(torch.from_numpy(a))
                    shapes = {tuple(t.shape) for t in subj}
                    if len(shapes) == 1:
                        ref_u8 = torch.stack(subj).to(dev)
                    else:  # subjects of other aspect ratios: BlipImageProcessor squashes every reference image to 224x224 anyway
                        S = pipe.qformer.cfg.image_size
                        ref_u8 = torch.cat([ops.resize_pil(t[None].to(dev).contiguous(), S, S, "bicubic") for t in subj], 0)
                else:
                    ref_u8 = src_t.index_select(0, sel).contiguous()
                extra = dict(reference_u8=ref_u8, source_subject=subject, target_subject=subject)
            text, neg, added = pipe._encode_call([w.prompt for w in chunk], None, cfg.NEGATIVE_PROMPT, None, do_cfg, H, W, **extra)
            shape = (1, pipe.vae_cfg.latent_channels, H // 8, W // 8)
            noise, post = [], []
            for w in chunk:
                if ref_noise is not None:
                    n_ref, p_ref = ref_noise[(w.index, w.i)]
                    noise.append(n_ref)
                    if cfg.SDEDIT:
                        post.append(p_ref)
                    continue
                g = torch.Generator().manual_seed(item_seed(cfg.SEED, w.index, w.i))
                if cfg.SDEDIT:
                    post.append(torch.randn(shape, generator=g))
                noise.append(torch.randn(shape, generator=g))
            noise = torch.cat(noise).to(dev)
            post = torch.cat(post).to(dev) if cfg.SDEDIT else None
            try:
                imgs = pipe.generate_batch(text, neg, None, src_t.index_select(0, sel) if cfg.SDEDIT else None,
                                           noise=noise, noise_posterior=post, num_inference_steps=cfg.NUM_INFERENCE_STEPS, guidance_scale=cfg.GUIDANCE_SCALE,
                                           strength=cfg.SDEDIT_STRENGTH, controlnet_conditioning_scale=1.0 if blip else cfg.CONTROLNET_CONDITIONING_SCALE,
                                           control_bf16=ctrl.index_select(0, sel), added=added)
            except RuntimeError as e:  # the reference logs OOM-style errors and leaves the loop (run_aug.py:493-500)
                logging.exception(e)
                num_errors += 1
                break
            pinned = torch.empty(imgs.shape, dtype=torch.uint8, pin_memory=True)
            pinned.copy_(imgs, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record()
            done_q.put((ev, pinned, chunk))
    finally:
        done_q.put(None)
        wt.join()
        loaders.shutdown(wait=True, cancel_futures=True)
        pool.shutdown(wait=True)
    for fut, rec in saves:
        err = fut.exception()
        if err is not None:
            logging.error(f"saving {rec[2] if rec else 'an auxiliary image'} failed: {err!r}")
        elif rec is not None:
            written.append(rec)
    logging.info(f"Done Generating ({len(written)} files on rank {rank}, {num_errors} errors)")
    return written


def verify_written(written, max_delete: int = 50):
    """PIL-verify THIS rank's files before they are decoded for the filter and unlink the corrupt ones (a truncated PNG of a killed
    run is accepted by the resume check, which only tests existence).  Same policy as the reference's
    check_folder_of_images_with_pil (all_utils/utils.py:681-703): delete, keep going, stop deleting after ``max_delete``."""
    from PIL import Image

    good, deleted = [], 0
    for rec in written:
        try:
            with Image.open(rec[2]) as im:
                im.verify()
            good.append(rec)
        except Exception:
            if deleted < max_delete:
                logging.info(f"image {rec[2]} is corrupted, deleting")
                try:
                    os.remove(rec[2])
                except OSError:
                    pass
                deleted += 1
    return good


def gather_records(records: np.ndarray, rank: int, world: int, failed: bool = False):
    """The ONE collective of the path: fixed-size per-image filter records -> every rank (torch.distributed all_gather over
    NCCL/NVLink on GPUs, gloo in CPU tests).  records: int32 [n_local, k]; ranks may hold different n_local.  ``failed`` rides along
    with the sizes: a rank whose generate / filter stage raised still enters the collective (so nobody hangs) and every rank then
    raises the same RuntimeError."""
    import torch
    import torch.distributed as dist

    if world == 1:
        if failed:
            raise RuntimeError("the generate / filter stage failed on this rank (see the log)")
        return records
    backend = dist.get_backend()
    dev = torch.device("cuda", torch.cuda.current_device()) if backend == "nccl" else torch.device("cpu")
    n = torch.tensor([records.shape[0], 1 if failed else 0], dtype=torch.int64, device=dev)
    sizes = [torch.zeros_like(n) for _ in range(world)]
    dist.all_gather(sizes, n)
    bad = [r for r, sz in enumerate(sizes) if int(sz[1].item())]
    if bad:
        raise RuntimeError(f"the generate / filter stage failed on rank(s) {bad}; no JSON is written")
    m = int(max(sz[0].item() for sz in sizes))
    k = records.shape[1]
    buf = torch.full((m, k), -1, dtype=torch.int32, device=dev)
    if records.shape[0]:
        buf[: records.shape[0]] = torch.from_numpy(records).to(dev)
    out = [torch.empty_like(buf) for _ in range(world)]
    dist.all_gather(out, buf)
    return np.concatenate([o[: int(sz[0].item())].cpu().numpy() for o, sz in zip(out, sizes)], 0)


_FILTER_CACHE: dict = {}


def _shard_filter(cfg: AugConfig, ds_utils, device, filter_models, batch_size):
    """The AugmentationFilter of this process (nets are built once: rank 0 reuses them for the extra substring matches)."""
    from .filter_nets import AugmentationFilter
    from .filtering import SEMANTIC_NEGATIVE_PROMPTS, load_filter_models

    key = (id(ds_utils), str(device), bool(cfg.SEMANTIC_FILTERING), bool(cfg.MODEL_CONFIDENCE_BASED_FILTERING), cfg.CONF_TOP_K, id(filter_models))
    if key not in _FILTER_CACHE:
        _FILTER_CACHE.clear()
        classifier, clip, tokenizer = (filter_models or load_filter_models)(ds_utils, device)
        prompt_ids = tokenizer([ds_utils.get_basic_prompt()] + SEMANTIC_NEGATIVE_PROMPTS) if cfg.SEMANTIC_FILTERING else None
        _FILTER_CACHE[key] = AugmentationFilter(classifier if cfg.MODEL_CONFIDENCE_BASED_FILTERING else None, clip if cfg.SEMANTIC_FILTERING else None,
                                                prompt_ids, min(cfg.CONF_TOP_K, ds_utils.num_classes), micro_batch=batch_size)
    return _FILTER_CACHE[key]


def filter_shard(cfg: AugConfig, ds_utils, written, device="cuda", batch_size: int = 64, filter_models=None) -> np.ndarray:
    """Runs the two enabled filters (run_aug.py:551-556) on THIS rank's augmentations -> fixed-size records int32 [n, 4] =
    (source index, aug index i, in_topk, semantic) for the final gather.  ``written`` must already be verified (``verify_written``)."""
    import torch

    from .filtering import _load_batches

    rec = np.ones((len(written), 4), np.int32)
    for k, (index, i, _) in enumerate(written):
        rec[k, 0], rec[k, 1] = index, i
    if not written or not (cfg.SEMANTIC_FILTERING or cfg.MODEL_CONFIDENCE_BASED_FILTERING):
        return rec
    flt = _shard_filter(cfg, ds_utils, device, filter_models, batch_size)
    label_of = ds_utils.get_image_path_to_class_id_dict() if cfg.MODEL_CONFIDENCE_BASED_FILTERING else None
    paths = ds_utils.original_images_paths
    dev = torch.device(device)
    for idx, imgs in _load_batches([w[2] for w in written], batch_size):
        # strict lookup, as the reference's image_path_to_class_id_dict[image_path] (utils.py:359): a source without a label is an error
        labels = torch.tensor([int(label_of[paths[written[k][0]]]) if label_of is not None else 0 for k in idx], dtype=torch.int32, device=dev)
        out = flt(torch.from_numpy(imgs).to(dev), labels)
        rec[idx, 2] = out["in_topk"].cpu().numpy()
        rec[idx, 3] = out["semantic"].cpu().numpy()
    return rec


def run_sharded(cfg: AugConfig, ds_utils, prompts, ds_root: str, pipe=None, device=None, filter_models=None,
                generate_fn=None, filter_fn=None):
    """End to end for one rank of a one-process-per-GPU job (RANK / WORLD_SIZE / LOCAL_RANK from torchrun; single process otherwise):
    generate this rank's shard (run_aug.py:357-471) -> verify + filter it on this GPU -> ONE collective (all-gather of the per-image
    filter records, NCCL over NVLink on GPUs) -> rank 0 writes the aug JSON through the reference-compatible writer (run_aug.py:721-733).
    Returns (json_path | None on ranks > 0, stats).  ``generate_fn`` / ``filter_fn`` replace the two GPU stages (defaults: ``generate``,
    ``filter_shard``); the CPU test of the multi-rank host logic (gloo, world size 2) injects stand-ins.
    A stage that raises on one rank does not strand the others in the collective: the failure travels with the record counts and
    every rank raises after the all-gather.  ``SASPA_DIST_BACKEND`` overrides the backend (gloo lets two ranks share one GPU in tests)."""
    import time

    import torch
    import torch.distributed as dist

    from . import filtering

    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if device is None:
        device = f"cuda:{local}"
    if torch.cuda.is_available():
        torch.cuda.set_device(torch.device(device))
    if world > 1 and not dist.is_initialized():
        dist.init_process_group(os.environ.get("SASPA_DIST_BACKEND") or ("nccl" if torch.cuda.is_available() else "gloo"))
    out_dir = output_folder(ds_root, cfg)
    t0 = time.perf_counter()
    written, rec, failed = [], np.zeros((0, 4), np.int32), False
    t1 = t2 = t0
    try:
        if pipe is None and generate_fn is None:
            pipe = init_pipeline(cfg.BASE_MODEL, cfg.CONTROLNET, cfg.SDEDIT, sampler=cfg.SAMPLER, device=device)
        written = (generate_fn or generate)(cfg, ds_utils, pipe, prompts, out_dir, rank=rank, world=world)
        if torch.cuda.is_available():
            torch.cuda.synchronize()
        t1 = time.perf_counter()
        written = verify_written(written)
        rec = (filter_fn or filter_shard)(cfg, ds_utils, written, device=device, filter_models=filter_models)
        t2 = time.perf_counter()
    except Exception as e:  # noqa: BLE001 -- reported through the collective so that no rank hangs
        logging.exception(e)
        failed = True
    allrec = gather_records(rec, rank, world, failed=failed)
    t3 = time.perf_counter()
    stats = {"rank": rank, "world": world, "generated": len(written), "generate_s": t1 - t0, "filter_s": t2 - t1, "gather_s": t3 - t2}
    json_path = None
    if rank == 0:
        paths = ds_utils.original_images_paths
        # (source index, i) -> file path: every rank used the same rank-independent prompt replay, so rank 0 can rebuild the names
        prompt_list = prompts if callable(prompts) else [p.strip()[:MAX_PROMPT_LENGTH] for p in prompts]
        draws = replay_prompt_draws(prompt_list, paths, cfg, ds_utils, skip=lambda *a: True)  # names only: no subject draws needed
        decisions = {}
        for r in allrec:
            index, i = int(r[0]), int(r[1])
            name = aug_file_name(Path(paths[index]).stem, draws[index][i].prompt, i)
            decisions[(Path(paths[index]).name, str(Path(out_dir) / name))] = (int(r[2]), int(r[3]))

        def extra(pairs):
            """(source, augmentation) pairs the substring rule of utils.py:352-354 adds beyond each rank's own records (one source's stem
            inside another's file name; files left by earlier runs): scored here, against THAT source's label."""
            by_name = {Path(p).name: k for k, p in enumerate(paths)}
            w = [(by_name[name], -1, path) for name, path in pairs]
            ok = {rec[2] for rec in verify_written(w)}  # files nobody generated in this run (leftovers): same verify-or-delete policy
            w = [rec for rec in w if rec[2] in ok]
            rr = (filter_fn or filter_shard)(cfg, ds_utils, w, device=device, filter_models=filter_models)
            return {(Path(paths[rec[0]]).name, rec[2]): (int(x[2]), int(x[3])) for rec, x in zip(w, rr)}

        json_path = filtering.create_json_of_image_name_to_augmented_images_paths(
            cfg.DATASET, out_dir, semantic_filtering=bool(cfg.SEMANTIC_FILTERING), model_confidence_based_filtering=bool(cfg.MODEL_CONFIDENCE_BASED_FILTERING),
            conf_top_k=cfg.CONF_TOP_K, init_log=False, ds_utils=ds_utils, decisions=decisions, missing_decisions=extra, assume_verified=True)
        stats.update(records=int(allrec.shape[0]), kept=int(sum(1 for a, b in decisions.values() if a and b)), json_s=time.perf_counter() - t3)
    if world > 1:
        dist.barrier()
    return json_path, stats
