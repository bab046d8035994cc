"""Checkpoint ingestion for the drop-in constructors.

The reference builds its pipelines with ``from_pretrained`` of HF repos (run_aug/run_aug.py:53-72, :184-211) and its baseline classifier
from exactly one ``.pth`` (all_utils/dataset_utils.py:87-115).  This module reads the same artefacts from LOCAL storage (there is no
network): a diffusers model directory (``unet/``, ``vae/``, ``text_encoder/`` [, ``text_encoder_2/``, ``qformer/``], each with a
``config.json`` and ``*.safetensors`` / ``*.bin`` weights, ``tokenizer/vocab.json`` + ``merges.txt``), a ControlNet directory, a WSDAN_CAL
``.pth``, an openai-clip checkpoint.  Weights are handed to the model classes as plain diffusers/transformers-keyed state dicts -- the
re-layout for the B200 kernels happens in their constructors (nn.py, filter_nets.py).

The safetensors container is read here directly (8-byte little-endian header length, JSON header of dtype / shape / data_offsets, raw
little-endian tensor bytes): no dependency, memory-mapped, zero-copy until a tensor is moved to the device.
"""
from __future__ import annotations

import json
import os
import struct
from pathlib import Path
from typing import Dict, Optional

import numpy as np
import torch

from . import checkpoints as ck

SD = Dict[str, torch.Tensor]

_ST_DTYPES = {"F64": (torch.float64, 8), "F32": (torch.float32, 4), "F16": (torch.float16, 2), "BF16": (torch.bfloat16, 2), "I64": (torch.int64, 8),
              "I32": (torch.int32, 4), "I16": (torch.int16, 2), "I8": (torch.int8, 1), "U8": (torch.uint8, 1), "BOOL": (torch.bool, 1)}
_ST_NAMES = {v[0]: k for k, v in _ST_DTYPES.items()}


def read_safetensors(path) -> SD:
    """-> {name: tensor} (CPU, views into one memory map of the file)."""
    path = str(path)
    with open(path, "rb") as f:
        head = f.read(8)
        if len(head) != 8:
            raise ValueError(f"{path}: not a safetensors file (truncated header)")
        (n,) = struct.unpack("<Q", head)
        if n <= 0 or n > 100 * 1024 * 1024:
            raise ValueError(f"{path}: implausible safetensors header length {n}")
        meta = json.loads(f.read(n).decode("utf-8"))
    base = 8 + n
    size = os.path.getsize(path)
    buf = np.memmap(path, dtype=np.uint8, mode="c")  # copy-on-write: the file is never modified, torch gets a writable view
    out: SD = {}
    for name, info in meta.items():
        if name == "__metadata__":
            continue
        if info["dtype"] not in _ST_DTYPES:
            raise ValueError(f"{path}: tensor {name!r} has unsupported dtype {info['dtype']}")
        dt, esz = _ST_DTYPES[info["dtype"]]
        b, e = info["data_offsets"]
        shape = tuple(info["shape"])
        numel = int(np.prod(shape)) if shape else 1
        if e - b != numel * esz or base + e > size:
            raise ValueError(f"{path}: tensor {name!r} offsets {b}:{e} do not match shape {shape} x {esz} B (file {size} B)")
        raw = torch.from_numpy(buf[base + b: base + e]) if e > b else torch.empty(0, dtype=torch.uint8)
        out[name] = raw.view(dt).reshape(shape)
    return out


def write_safetensors(path, sd: SD, metadata: Optional[dict] = None) -> None:
    """Writer for tests / conversions (same container)."""
    header, off, blobs = {}, 0, []
    if metadata:
        header["__metadata__"] = {str(k): str(v) for k, v in metadata.items()}
    for name, t in sd.items():
        t = t.detach().cpu().contiguous()
        raw = t.view(torch.uint8).numpy().tobytes() if t.numel() else b""
        header[name] = {"dtype": _ST_NAMES[t.dtype], "shape": list(t.shape), "data_offsets": [off, off + len(raw)]}
        off += len(raw)
        blobs.append(raw)
    h = json.dumps(header, separators=(",", ":")).encode("utf-8")
    h += b" " * ((8 - len(h) % 8) % 8)
    with open(path, "wb") as f:
        f.write(struct.pack("<Q", len(h)))
        f.write(h)
        for raw in blobs:
            f.write(raw)


def load_state_dict_file(path) -> SD:
    path = str(path)
    if path.endswith(".safetensors"):
        return read_safetensors(path)
    obj = torch.load(path, map_location="cpu", weights_only=True)
    if isinstance(obj, dict) and "state_dict" in obj and isinstance(obj["state_dict"], dict):
        obj = obj["state_dict"]
    return obj


_WEIGHT_NAMES = ["diffusion_pytorch_model.fp16.safetensors", "diffusion_pytorch_model.safetensors", "model.fp16.safetensors", "model.safetensors",
                 "diffusion_pytorch_model.fp16.bin", "diffusion_pytorch_model.bin", "pytorch_model.fp16.bin", "pytorch_model.bin"]


def find_weights(folder, prefer_fp16: bool = True) -> Path:
    """The weight file ``from_pretrained`` would pick in a component folder (fp16 variant first when asked, safetensors before .bin)."""
    folder = Path(folder)
    names = _WEIGHT_NAMES if prefer_fp16 else [n for n in _WEIGHT_NAMES if ".fp16." not in n] + [n for n in _WEIGHT_NAMES if ".fp16." in n]
    for n in names:
        if (folder / n).exists():
            return folder / n
    raise FileNotFoundError(f"no weight file ({', '.join(_WEIGHT_NAMES[:4])}, ...) under {folder}")


def resolve_model_dir(model_id: str) -> Optional[Path]:
    """Local directory of an HF model id, or None.  Looked up, in order: the id itself as a path; $SASPA_MODEL_ROOT/<id> and
    $SASPA_MODEL_ROOT/<id with '/' -> '--'>; the HF hub cache ($HF_HUB_CACHE | $HF_HOME/hub | ~/.cache/huggingface/hub) layout
    models--org--name/snapshots/<rev>/."""
    if os.path.isdir(model_id):
        return Path(model_id)
    roots = []
    if os.environ.get("SASPA_MODEL_ROOT"):
        r = Path(os.environ["SASPA_MODEL_ROOT"])
        roots += [r / model_id, r / model_id.replace("/", "--")]
    for p in roots:
        if p.is_dir():
            return p
    hub = os.environ.get("HF_HUB_CACHE") or os.path.join(os.environ.get("HF_HOME", os.path.expanduser("~/.cache/huggingface")), "hub")
    snaps = Path(hub) / ("models--" + model_id.replace("/", "--")) / "snapshots"
    if snaps.is_dir():
        revs = sorted(p for p in snaps.iterdir() if p.is_dir())
        if revs:
            return revs[-1]
    return None


# ---- config.json -> config dataclasses -------------------------------------------------------------------------------------
def _tuple(v, n):
    return tuple(v) if isinstance(v, (list, tuple)) else (v,) * n


def unet_config_from_json(path, controlnet_json=None) -> ck.UNetConfig:
    """diffusers UNet2DConditionModel config.json (note its quirk: ``attention_head_dim`` holds the number of HEADS for SD v1.x / SDXL
    when ``num_attention_heads`` is null)."""
    c = json.load(open(path))
    nlev = len(c["block_out_channels"])
    heads = c.get("num_attention_heads") or c["attention_head_dim"]
    kw = dict(in_channels=c["in_channels"], out_channels=c.get("out_channels", 4), block_out_channels=tuple(c["block_out_channels"]),
              down_block_types=tuple(c["down_block_types"]), up_block_types=tuple(c["up_block_types"]), layers_per_block=c["layers_per_block"],
              transformer_layers_per_block=_tuple(c.get("transformer_layers_per_block", 1), nlev), num_attention_heads=_tuple(heads, nlev),
              cross_attention_dim=c["cross_attention_dim"], norm_num_groups=c["norm_num_groups"], norm_eps=c["norm_eps"],
              use_linear_projection=bool(c.get("use_linear_projection", False)), flip_sin_to_cos=bool(c.get("flip_sin_to_cos", True)),
              freq_shift=float(c.get("freq_shift", 0)), addition_embed_type=c.get("addition_embed_type"),
              addition_time_embed_dim=c.get("addition_time_embed_dim") or 256,
              projection_class_embeddings_input_dim=c.get("projection_class_embeddings_input_dim") or 2816)
    if controlnet_json is not None:
        cc = json.load(open(controlnet_json))
        kw["conditioning_embedding_out_channels"] = tuple(cc.get("conditioning_embedding_out_channels", (16, 32, 96, 256)))
    return ck.UNetConfig(**kw)


def vae_config_from_json(path) -> ck.VAEConfig:
    c = json.load(open(path))
    return ck.VAEConfig(in_channels=c["in_channels"], out_channels=c["out_channels"], latent_channels=c["latent_channels"],
                        block_out_channels=tuple(c["block_out_channels"]), layers_per_block=c["layers_per_block"], norm_num_groups=c["norm_num_groups"],
                        scaling_factor=float(c.get("scaling_factor", 0.18215)))


def clip_text_config_from_json(path) -> ck.CLIPTextConfig:
    c = json.load(open(path))
    with_proj = "CLIPTextModelWithProjection" in (c.get("architectures") or [])
    return ck.CLIPTextConfig(vocab_size=c["vocab_size"], hidden_size=c["hidden_size"], intermediate_size=c["intermediate_size"],
                             num_hidden_layers=c["num_hidden_layers"], num_attention_heads=c["num_attention_heads"],
                             max_position_embeddings=c["max_position_embeddings"], hidden_act=c.get("hidden_act", "quick_gelu"),
                             layer_norm_eps=c.get("layer_norm_eps", 1e-5), projection_dim=c.get("projection_dim", 0) if with_proj else 0)


def scheduler_kwargs_from_json(path) -> dict:
    """The fields of scheduler_config.json the three schedulers read (``X.from_config(pipe.scheduler.config)``, run_aug.py:219-228)."""
    c = json.load(open(path))
    keep = ("num_train_timesteps", "beta_start", "beta_end", "beta_schedule", "steps_offset", "timestep_spacing", "clip_sample", "set_alpha_to_one",
            "prediction_type", "skip_prk_steps")
    return {k: c[k] for k in keep if k in c}


def load_pipeline_dir(base_dir, controlnet_dir=None, vae_dir=None, prefer_fp16: bool = True) -> dict:
    """Reads a diffusers pipeline directory (+ a separate ControlNet directory, + an override VAE directory as the SDXL path uses,
    run_aug.py:189) -> {"unet": sd, "controlnet": sd | None, "vae": sd, "text": sd, ["text2": sd], ["qformer": sd], "configs": {...},
    "tokenizer": CLIPBPETokenizer | None, "tokenizer_2": ..., "scheduler": kwargs}."""
    from .tokenizer import CLIPBPETokenizer

    base = Path(base_dir)
    out = {"configs": {}}
    cn_json = None
    cn_dir = Path(controlnet_dir) if controlnet_dir is not None else (base / "controlnet" if (base / "controlnet").is_dir() else None)
    if cn_dir is not None:
        cn_json = cn_dir / "config.json"
        out["controlnet"] = load_state_dict_file(find_weights(cn_dir, prefer_fp16))
    else:
        out["controlnet"] = None
    out["unet"] = load_state_dict_file(find_weights(base / "unet", prefer_fp16))
    out["configs"]["unet"] = unet_config_from_json(base / "unet" / "config.json", cn_json)
    vdir = Path(vae_dir) if vae_dir is not None else base / "vae"
    out["vae"] = load_state_dict_file(find_weights(vdir, prefer_fp16))
    out["configs"]["vae"] = vae_config_from_json(vdir / "config.json")
    out["text"] = load_state_dict_file(find_weights(base / "text_encoder", prefer_fp16))
    out["configs"]["text"] = clip_text_config_from_json(base / "text_encoder" / "config.json")
    if (base / "text_encoder_2").is_dir():
        out["text2"] = load_state_dict_file(find_weights(base / "text_encoder_2", prefer_fp16))
        out["configs"]["text2"] = clip_text_config_from_json(base / "text_encoder_2" / "config.json")
    if (base / "qformer").is_dir():
        out["qformer"] = load_state_dict_file(find_weights(base / "qformer", prefer_fp16))
    for key, sub in (("tokenizer", "tokenizer"), ("tokenizer_2", "tokenizer_2")):
        t = base / sub
        out[key] = None
        if (t / "vocab.json").exists() and (t / "merges.txt").exists():
            pad = None
            if (t / "special_tokens_map.json").exists():
                pad_tok = json.load(open(t / "special_tokens_map.json")).get("pad_token")
                pad = pad_tok["content"] if isinstance(pad_tok, dict) else pad_tok
            out[key] = CLIPBPETokenizer.from_files(t / "vocab.json", t / "merges.txt", pad_token=pad)
    out["scheduler"] = scheduler_kwargs_from_json(base / "scheduler" / "scheduler_config.json") if (base / "scheduler" / "scheduler_config.json").exists() else {}
    return out


# ---- baseline classifier (all_utils/dataset_utils.py:87-115) ------------------------------------------------------------------
def _strip_compile_prefix(sd: SD) -> SD:
    return {k.replace("_orig_mod.", ""): v for k, v in sd.items()}


def wsdan_net_of(sd: SD, num_classes: int) -> str:
    """Which trunk a WSDAN_CAL state dict was trained with.  The reference tries ``WSDAN_CAL(num_classes)`` (ResNet-101, cal.py:132)
    and, when ``load_state_dict`` raises, ResNet-50 (dataset_utils.py:99-109): the same order, decided by strict key / shape agreement."""
    keys = {k: tuple(v.shape) for k, v in sd.items() if not k.endswith("num_batches_tracked")}
    why = {}
    for net in ("resnet101", "resnet50"):
        want = dict(ck.wsdan_shapes(num_classes, net))
        missing = [k for k in want if k not in keys]
        unexpected = [k for k in keys if k not in want]
        wrong = [k for k in want if k in keys and tuple(want[k]) != keys[k]]
        if not missing and not unexpected and not wrong:
            return net
        why[net] = f"{len(missing)} missing (e.g. {missing[:1]}), {len(unexpected)} unexpected (e.g. {unexpected[:1]}), {len(wrong)} wrong shapes (e.g. {wrong[:1]})"
    raise RuntimeError(f"state dict matches neither WSDAN_CAL trunk for {num_classes} classes: {why}")


def load_wsdan_checkpoint(path, num_classes: int, device="cuda"):
    """``torch.load(cp)['state_dict']`` -> WSDANClassifier.  (weights_only=False as the reference: its checkpoints pickle the optimizer
    / logs next to the state dict.)"""
    from .filter_nets import WSDANClassifier

    checkpoint = torch.load(str(path), map_location="cpu", weights_only=False)
    sd = _strip_compile_prefix(checkpoint["state_dict"])
    return WSDANClassifier(sd, num_classes, wsdan_net_of(sd, num_classes), device)


def wsdan_from_reference_model(model_or_sd, num_classes: int, device="cuda"):
    """A torch ``WSDAN_CAL`` module (what the reference's load_baseline_model returns) or its state dict -> WSDANClassifier."""
    from .filter_nets import WSDANClassifier

    sd = model_or_sd if isinstance(model_or_sd, dict) else model_or_sd.state_dict()
    sd = _strip_compile_prefix({k: v.detach().cpu() for k, v in sd.items()})
    return WSDANClassifier(sd, num_classes, wsdan_net_of(sd, num_classes), device)


def load_clip_checkpoint(path, device="cuda"):
    """openai-clip ``RN50.pt`` / ``ViT-L-14.pt`` (TorchScript archive, or a plain state dict) -> CLIPRN50 | CLIPViT."""
    from .filter_nets import CLIPRN50, CLIPViT

    try:
        sd = torch.jit.load(str(path), map_location="cpu").state_dict()
    except RuntimeError:
        sd = load_state_dict_file(path)
    sd = {k: v for k, v in sd.items() if k not in ("input_resolution", "context_length", "vocab_size")}
    return (CLIPViT if "visual.proj" in sd else CLIPRN50)(sd, device)


def filter_models_from_reference_interface(ds_utils, device):
    """(classifier, clip, tokenizer) for a dataset-utils class written against the REFERENCE's BaseUtils (it offers
    ``load_baseline_model`` and nothing B200-specific)."""
    from .filter_nets import WSDANClassifier
    from .tokenizer import CLIPBPETokenizer

    model, _ = ds_utils.load_baseline_model()
    clf = model if isinstance(model, WSDANClassifier) else wsdan_from_reference_model(model, len(ds_utils.get_classes()), device)
    if hasattr(ds_utils, "load_clip"):
        clip, tok = ds_utils.load_clip(device)
        return clf, clip, tok
    cp, vocab = os.environ.get("SASPA_CLIP_RN50"), os.environ.get("SASPA_CLIP_BPE")
    if not cp or not vocab:
        raise FileNotFoundError("CLIP RN50 is needed for the semantic filter (all_utils/utils.py:253): set $SASPA_CLIP_RN50 to the openai-clip checkpoint "
                                "and $SASPA_CLIP_BPE to bpe_simple_vocab_16e6.txt.gz, or give the dataset class a load_clip(device) method")
    return clf, load_clip_checkpoint(cp, device), CLIPBPETokenizer.from_openai_bpe(vocab)
