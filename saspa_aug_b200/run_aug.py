"""Sharded generation driver: host-side mirror of the reference's ``run_aug/run_aug.py`` for the ControlNet-canny path.

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
from dataclasses import dataclass, field
from pathlib import Path
from typing import List, Optional, Sequence

import numpy as np

NEGATIVE_PROMPT = ("over-exposure, under-exposure, saturated, duplicate, out of frame, lowres, cropped, worst quality, low quality, jpeg artifacts, morbid, "
                   "mutilated, out of frame, ugly, bad anatomy, bad proportions, deformed, blurry, duplicate")  # the run constant of run_aug.py:47
MAX_FILENAME_LENGTH = 40
MAX_PROMPT_LENGTH = 150


@dataclass
class AugConfig:
    """The reference's module-level constants (run_aug.py:513-556) with their shipped defaults."""
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
    PROMPT_WITH_SUB_CLASS: bool = False
    USE_ARTISTIC_PROMPTS: bool = False
    ARTISTIC_PROMPTS_PROB: float = 0.5
    SEMANTIC_FILTERING: int = 1
    MODEL_CONFIDENCE_BASED_FILTERING: int = 1
    CONF_TOP_K: int = 10
    SAMPLER: str = "ddim"
    RNG_MODE: str = "per_item"  # "per_item" (partition independent) | "reference_order" (replays the global generator)
    MICRO_BATCH: int = 16

    def apply_dataset_rules(self):
        """run_aug.py:560-577."""
        if self.DATASET == "cars":
            self.NUM_INFERENCE_STEPS = 50
        if self.BASE_MODEL == "sd_xl-turbo":
            self.GUIDANCE_SCALE, self.NUM_INFERENCE_STEPS = 0.0, 2
        if self.SDEDIT:
            assert self.NUM_INFERENCE_STEPS * self.SDEDIT_STRENGTH >= 1, "NUM_INFERENCE_STEPS * SDEDIT_STRENGTH must be >= 1"
        return self


ARTISTIC_PROMPTS = ["a painting", "a sketch", "a watercolor", "an oil painting", "a pencil drawing"]  # stand-in list (prompts_engineering assets are data)


def output_folder(ds_root: str, cfg: AugConfig) -> str:
    """run_aug.py:668-692."""
    base = f"{cfg.BASE_MODEL}-SDEdit_strength_{cfg.SDEDIT_STRENGTH}" if cfg.SDEDIT else cfg.BASE_MODEL
    kind = "controlnet" if cfg.CONTROLNET else "regular"
    prompt_str = cfg.PROMPT_TYPE
    if cfg.PROMPT_WITH_SUB_CLASS:
        prompt_str += "_prompt_w_sub_class"
    if cfg.USE_ARTISTIC_PROMPTS:
        prompt_str += f"_artistic_prompts_p_{cfg.ARTISTIC_PROMPTS_PROB}"
    return str(Path(ds_root) / "aug_data" / kind / base / str(cfg.CONTROLNET) / f"{prompt_str}_seed_{cfg.SEED}" / "images")


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


def resize_image(input_image: np.ndarray, smaller_side_res: int) -> np.ndarray:
    """all_utils/utils.py:58-79 (host-side pre-processing; identity for sources already at HxW % 64 == 0, min side = res)."""
    import cv2

    H, W, k = resized_hw(input_image.shape[0], input_image.shape[1], smaller_side_res)
    if (H, W) == input_image.shape[:2]:
        return input_image
    return cv2.resize(input_image, (W, H), interpolation=cv2.INTER_LANCZOS4 if k > 1 else cv2.INTER_AREA)


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


def init_pipeline(base_model, controlnet, SDEdit, use_compile=False, sampler="ddim", state_dicts=None, device="cuda"):
    """run_aug.py:128-230 for the ControlNet-canny pipelines (SD v1.5, SD-XL(-turbo), BLIP-Diffusion).  ``use_compile`` is accepted
    and ignored: nothing is traced, the kernels are launched directly (the step is GPU-bound: ~550 launches per 94 ms at micro-batch
    32).  Weights: ``state_dicts`` (diffusers-keyed) or deterministic random init -- no checkpoints exist offline."""
    from .pipelines import (SaspaBlipControlNetPipeline, SaspaControlNetPipeline, SaspaSDXLControlNetPipeline, blip_configs, random_state_dicts,
                            sdxl_configs)

    assert sampler in ["ddim", "unipcmultistep"]
    assert controlnet in ("canny",), "only the canny ControlNet is on the hot path"
    if base_model in ("sd_xl-turbo", "sd_xl", "tiny_xl"):
        # run_aug.py:188-199 (+ :223-228: for sd_xl-turbo the scheduler is rebuilt from the turbo config => trailing spacing)
        cfg = "tiny_xl" if base_model == "tiny_xl" else "sdxl"
        sds = state_dicts or random_state_dicts(cfg, 1234)
        u, v, t1, t2 = sdxl_configs(cfg)
        turbo = base_model != "sd_xl"
        smp = ("unipc_sdxl_turbo" if sampler == "unipcmultistep" else "ddim_sdxl_turbo") if turbo else ("unipc" if sampler == "unipcmultistep" else "ddim")
        return SaspaSDXLControlNetPipeline.from_state_dicts(sds["unet"], sds["controlnet"], sds["vae"], sds["text"], sds["text2"], unet_cfg=u, vae_cfg=v,
                                                            text_cfg=t1, text2_cfg=t2, sampler=smp, device=device, img2img=bool(SDEdit))
    if base_model in ("blip_diffusion", "tiny_blip"):
        # run_aug.py:185-187: BlipDiffusionControlNetPipeline; the sampler stays the checkpoint's PNDM (run_aug.py:217 skips the swap)
        cfg = "tiny_blip" if base_model == "tiny_blip" else "blip"
        sds = state_dicts or random_state_dicts(cfg, 1234)
        u, v, t, q = blip_configs(cfg)
        return SaspaBlipControlNetPipeline.from_state_dicts(sds["unet"], sds["controlnet"], sds["vae"], sds["text"], sds["qformer"], unet_cfg=u, vae_cfg=v,
                                                            text_cfg=t, qformer_cfg=q, device=device)
    if base_model not in ("sd_v1.5", "tiny"):
        raise NotImplementedError(f"base_model {base_model!r}: SD v1.5, SD-XL(-turbo) and BLIP-Diffusion ControlNet paths are built (see DESIGN.md)")
    cfg = "tiny" if base_model == "tiny" else "sd15"
    sds = state_dicts or random_state_dicts(cfg, 1234)
    from . import checkpoints as ck

    kw = {}
    if cfg == "tiny":
        kw = dict(unet_cfg=ck.UNetConfig.tiny(), vae_cfg=ck.VAEConfig.tiny(), text_cfg=ck.CLIPTextConfig.tiny())
    smp = "unipc" if sampler == "unipcmultistep" else "ddim"
    return SaspaControlNetPipeline.from_state_dicts(sds["unet"], sds["controlnet"], sds["vae"], sds["text"], sampler=smp, device=device, img2img=bool(SDEdit), **kw)


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


def reference_order_noise(cfg: AugConfig, source_hw: Sequence, sampled: Sequence[Sequence[str]], skip, mine, latent_channels: int = 4,
                          dtype=None):
    """RNG_MODE = "reference_order" (SURVEY.md 8e): the reference threads ONE generator -- ``torch.manual_seed(SEED)``, the global CPU
    generator, run_aug.py:324 -- through every pipeline call in dataset order (:464), and diffusers draws from it on the CPU in the
    pipeline dtype (fp16 after ``.to(DEVICE, torch.float16)``, :323): per call the VAE-posterior noise first when SDEdit is on, then the
    latent noise.  This rank-independent pre-pass replays exactly that stream (a private generator with the same seed yields the same
    values) and keeps the draws of the items in ``mine``; items the reference would skip (existing output, :430-432) draw nothing.
      source_hw  [(H, W)] of every source AFTER resize_image;  skip(index, i) -> bool;  mine: set of source indices of this rank
    -> {(index, i): (noise fp32 [1,C,H/8,W/8], posterior noise fp32 | None)}"""
    import torch

    dtype = dtype or torch.float16
    g = torch.Generator().manual_seed(cfg.SEED)
    out = {}
    for index, (H, W) in enumerate(source_hw):
        shape = (1, latent_channels, H // 8, W // 8)
        for i in range(len(sampled[index])):
            if skip(index, i):
                continue
            post = torch.randn(shape, generator=g, dtype=dtype) if cfg.SDEDIT else None
            noise = torch.randn(shape, generator=g, dtype=dtype)
            if index in mine:
                out[(index, i)] = (noise.float(), post.float() if post is not None else None)
    return out


def sample_prompts(prompts: Sequence[str], n_sources: int, cfg: AugConfig) -> List[List[str]]:
    """Rank-independent pre-pass replaying the reference's sequential global draws (run_aug.py:382, :391-394):
    np.random.seed(SEED) then one np.random.choice(prompts, NUM_PER_IMAGE) per source in dataset order."""
    rs = np.random.RandomState(cfg.SEED)
    prompts = [p.strip()[:MAX_PROMPT_LENGTH] for p in prompts]
    prompts = [p[:-1] if p and p[-1] == "." else p for p in prompts]
    out = []
    for _ in range(n_sources):
        chosen = list(rs.choice(prompts, cfg.NUM_PER_IMAGE))
        for i in range(len(chosen)):
            if cfg.USE_ARTISTIC_PROMPTS and i % 2 == 0 and cfg.ARTISTIC_PROMPTS_PROB == 0.5:
                chosen[i] = f"{chosen[i]}, {rs.choice(ARTISTIC_PROMPTS)}"
        out.append([str(c) for c in chosen])
    return out


def generate(cfg: AugConfig, ds_utils, pipe, prompts: Sequence[str], out_dir: str, rank: int = 0, world: int = 1, io_threads: int = 8):
    """The generation loop (run_aug.py:357-471) for this rank's shard.  Returns the list of (source index, i, path).
    Works for the three ControlNet base models: SD v1.5 / SD-XL(-turbo) (text + added conditioning from the pipeline's own
    ``_encode_call``) and BLIP-Diffusion (reference image = the source, subject categories = ``ds_utils.meta_class``,
    run_aug.py:444-456; conditioning scale 1.0, no SDEdit)."""
    import torch
    from PIL import Image

    from . import ops

    Path(out_dir).mkdir(parents=True, exist_ok=True)
    paths = ds_utils.original_images_paths
    sampled = sample_prompts(prompts, len(paths), cfg)
    mine = shard_indices(len(paths), rank, world)
    ref_noise = None
    if cfg.RNG_MODE == "reference_order":
        def _hw(p):
            with Image.open(p) as im:  # header only
                w, h = im.size
            return resized_hw(h, w, cfg.RESOLUTION)[:2]

        ref_noise = reference_order_noise(cfg, [_hw(p) for p in paths], sampled,
                                          lambda index, i: (Path(out_dir) / aug_file_name(Path(paths[index]).stem, sampled[index][i], i)).exists(),
                                          set(mine), pipe.vae_cfg.latent_channels)
    elif cfg.RNG_MODE != "per_item":
        raise ValueError(f"RNG_MODE {cfg.RNG_MODE!r}: per_item | reference_order")
    pool = ThreadPoolExecutor(max_workers=io_threads)
    written = []
    work = []  # (index, i, prompt, output_path)
    sources = {}
    num_errors = 0
    try:
        for index in mine:
            stem = Path(paths[index]).stem
            img = resize_image(np.array(Image.open(paths[index]).convert("RGB")), cfg.RESOLUTION)
            src_out = os.path.join(out_dir, f"{stem[:MAX_FILENAME_LENGTH]}_source.png")
            if not os.path.exists(src_out):
                pool.submit(Image.fromarray(img).save, src_out)
            for i, prompt in enumerate(sampled[index]):
                out_path = Path(out_dir) / aug_file_name(stem, prompt, i)
                if out_path.exists():  # resume (run_aug.py:430-432)
                    logging.info(f"Skipping {out_path} as it already exists")
                    written.append((index, i, str(out_path)))
                    continue
                sources[index] = img
                work.append((index, i, prompt, str(out_path)))
        dev = pipe.device
        blip = "blip" in cfg.BASE_MODEL
        assert not (blip and cfg.SDEDIT), "BLIP-Diffusion has no img2img ControlNet pipeline in the reference (run_aug.py:185-187)"
        do_cfg = cfg.GUIDANCE_SCALE > 1.0
        for b0 in range(0, len(work), cfg.MICRO_BATCH):
            chunk = work[b0 : b0 + cfg.MICRO_BATCH]
            uniq = sorted({w[0] for w in chunk})
            shapes = {sources[u].shape for u in uniq}
            if len(shapes) != 1:  # mixed resolutions: fall back to per-source batches
                raise RuntimeError("mixed source resolutions in one micro-batch; set MICRO_BATCH = NUM_PER_IMAGE for non-uniform datasets")
            src_t = torch.from_numpy(np.stack([sources[u] for u in uniq])).to(dev)
            edges, ctrl = ops.canny(src_t, cfg.LOW_THRESHOLD_CANNY, cfg.HIGH_THRESHOLD_CANNY, out_channels=3, want_ctrl=True)
            sel = torch.tensor([uniq.index(w[0]) for w in chunk], device=dev)
            for u_i, u in enumerate(uniq):
                if u < 10:  # first 10 control images are saved (run_aug.py:441-442)
                    cp = os.path.join(out_dir, f"{Path(paths[u]).stem[:MAX_FILENAME_LENGTH]}_control.png")
                    if not os.path.exists(cp):
                        pool.submit(Image.fromarray(edges[u_i].cpu().numpy()).save, cp)
            H, W = src_t.shape[1:3]
            extra = {}
            if blip:
                subject = getattr(ds_utils, "meta_class", "object")
                extra = dict(reference_u8=src_t.index_select(0, sel).contiguous(), source_subject=subject, target_subject=subject)
            text, neg, added = pipe._encode_call([w[2] for w in chunk], None, NEGATIVE_PROMPT, None, do_cfg, H, W, **extra)
            shape = (1, pipe.vae_cfg.latent_channels, H // 8, W // 8)
            noise, post = [], []
            for (index, i, _, _) in chunk:
                if ref_noise is not None:
                    n_ref, p_ref = ref_noise[(index, i)]
                    noise.append(n_ref)
                    if cfg.SDEDIT:
                        post.append(p_ref)
                    continue
                g = torch.Generator().manual_seed(item_seed(cfg.SEED, index, i))
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
            arr = imgs.cpu().numpy()
            for (index, i, _, out_path), a in zip(chunk, arr):
                pool.submit(Image.fromarray(a).save, out_path)
                written.append((index, i, out_path))
    finally:
        pool.shutdown(wait=True)
    logging.info(f"Done Generating ({len(written)} files on rank {rank}, {num_errors} errors)")
    return written


def gather_records(records: np.ndarray, rank: int, world: int):
    """The ONE collective of the path: fixed-size per-image filter records -> rank 0 (torch.distributed all_gather over
    NCCL/NVLink on GPUs, gloo in CPU tests).  records: int32 [n_local, k]; ranks may hold different n_local."""
    import torch
    import torch.distributed as dist

    if world == 1:
        return records
    backend = dist.get_backend()
    dev = torch.device("cuda", torch.cuda.current_device()) if backend == "nccl" else torch.device("cpu")
    n = torch.tensor([records.shape[0]], dtype=torch.int64, device=dev)
    sizes = [torch.zeros_like(n) for _ in range(world)]
    dist.all_gather(sizes, n)
    m = int(max(s.item() for s in sizes))
    k = records.shape[1]
    buf = torch.full((m, k), -1, dtype=torch.int32, device=dev)
    if records.shape[0]:
        buf[: records.shape[0]] = torch.from_numpy(records).to(dev)
    out = [torch.empty_like(buf) for _ in range(world)]
    dist.all_gather(out, buf)
    return np.concatenate([o[: int(s.item())].cpu().numpy() for o, s in zip(out, sizes)], 0)


def filter_shard(cfg: AugConfig, ds_utils, written, device="cuda", batch_size: int = 64, filter_models=None) -> np.ndarray:
    """Runs the two enabled filters (run_aug.py:551-556) on THIS rank's augmentations -> fixed-size records int32 [n, 4] =
    (source index, aug index i, in_topk, semantic) for the final gather."""
    import torch

    from .filter_nets import AugmentationFilter
    from .filtering import SEMANTIC_NEGATIVE_PROMPTS, _load_batches

    rec = np.ones((len(written), 4), np.int32)
    for k, (index, i, _) in enumerate(written):
        rec[k, 0], rec[k, 1] = index, i
    if not written or not (cfg.SEMANTIC_FILTERING or cfg.MODEL_CONFIDENCE_BASED_FILTERING):
        return rec
    classifier, clip, tokenizer = (filter_models or ds_utils.load_filter_models)(ds_utils, device)
    prompt_ids = tokenizer([ds_utils.get_basic_prompt()] + SEMANTIC_NEGATIVE_PROMPTS) if cfg.SEMANTIC_FILTERING else None
    flt = AugmentationFilter(classifier if cfg.MODEL_CONFIDENCE_BASED_FILTERING else None, clip if cfg.SEMANTIC_FILTERING else None, prompt_ids,
                             min(cfg.CONF_TOP_K, ds_utils.num_classes), micro_batch=batch_size)
    label_of = ds_utils.get_image_path_to_class_id_dict() if cfg.MODEL_CONFIDENCE_BASED_FILTERING else {}
    paths = ds_utils.original_images_paths
    dev = torch.device(device)
    for idx, imgs in _load_batches([w[2] for w in written], batch_size):
        labels = torch.tensor([int(label_of.get(paths[written[k][0]], 0)) for k in idx], dtype=torch.int32, device=dev)
        out = flt(torch.from_numpy(imgs).to(dev), labels)
        rec[idx, 2] = out["in_topk"].cpu().numpy()
        rec[idx, 3] = out["semantic"].cpu().numpy()
    return rec


def run_sharded(cfg: AugConfig, ds_utils, prompts: Sequence[str], ds_root: str, pipe=None, device=None, filter_models=None,
                generate_fn=None, filter_fn=None):
    """End to end for one rank of a one-process-per-GPU job (RANK / WORLD_SIZE / LOCAL_RANK from torchrun; single process otherwise):
    generate this rank's shard (run_aug.py:357-471) -> filter it on this GPU -> ONE collective (all-gather of the per-image filter
    records, NCCL over NVLink on GPUs) -> rank 0 writes the aug JSON through the reference-compatible writer (run_aug.py:721-733).
    Returns (json_path | None on ranks > 0, stats).  ``generate_fn`` / ``filter_fn`` replace the two GPU stages (defaults: ``generate``,
    ``filter_shard``); the CPU test of the multi-rank host logic (gloo, world size 2) injects stand-ins."""
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
        dist.init_process_group("nccl" if torch.cuda.is_available() else "gloo")
    if pipe is None and generate_fn is None:
        pipe = init_pipeline(cfg.BASE_MODEL, cfg.CONTROLNET, cfg.SDEDIT, sampler=cfg.SAMPLER, device=device)
    out_dir = output_folder(ds_root, cfg)
    t0 = time.perf_counter()
    written = (generate_fn or generate)(cfg, ds_utils, pipe, prompts, out_dir, rank=rank, world=world)
    t1 = time.perf_counter()
    rec = (filter_fn or filter_shard)(cfg, ds_utils, written, device=device, filter_models=filter_models)
    t2 = time.perf_counter()
    allrec = gather_records(rec, rank, world)
    stats = {"rank": rank, "world": world, "generated": len(written), "generate_s": t1 - t0, "filter_s": t2 - t1}
    json_path = None
    if rank == 0:
        paths = ds_utils.original_images_paths
        by_key = {(int(r[0]), int(r[1])): (int(r[2]), int(r[3])) for r in allrec}
        # (source index, i) -> file path: every rank used the same rank-independent prompt draw, so rank 0 can rebuild the names
        sampled = sample_prompts(prompts, len(paths), cfg)
        decisions = {}
        for (index, i), d in by_key.items():
            decisions[str(Path(out_dir) / aug_file_name(Path(paths[index]).stem, sampled[index][i], i))] = d
        json_path = filtering.create_json_of_image_name_to_augmented_images_paths(
            cfg.DATASET, out_dir, semantic_filtering=bool(cfg.SEMANTIC_FILTERING), model_confidence_based_filtering=bool(cfg.MODEL_CONFIDENCE_BASED_FILTERING),
            conf_top_k=cfg.CONF_TOP_K, init_log=False, ds_utils=ds_utils, decisions=decisions)
        stats.update(records=int(allrec.shape[0]), kept=int(sum(1 for a, b in by_key.values() if a and b)))
    if world > 1:
        dist.barrier()
    return json_path, stats
