"""Deterministic augmentation folder shared by the golden generator (tests/golden/make_filter_json_golden.py, which runs the
REFERENCE's own create_json_of_image_name_to_augmented_images_paths on it) and by the GPU test that must reproduce the
reference's JSON byte for byte.  Only seeds live here; every file is regenerated from them.

The folder exercises the corners of all_utils/utils.py:246-355:
  * substring matching: stems "syn_1" / "syn_10" / "syn_11" -- the augmentations of syn_10 and syn_11 also match source syn_1 and are
    scored there against syn_1's label;
  * excluded names (_source. / _control.), an unrelated leftover file, a truncated PNG that check_folder_of_images_with_pil deletes;
  * non-square augmentations (CLIP's Resize(224) keeps the aspect ratio, the classifier's Resize((256,256)) does not).
"""
from __future__ import annotations

import os
from pathlib import Path

import numpy as np

N_SOURCES = 12
NUM_CLASSES = 16
NUM_PER_IMAGE = 2
WSDAN_SEED = 4242
CLIP_SEED = 797
SIZES = [(160, 160), (128, 192), (192, 128)]
KINDS = ("blobs", "smooth", "noise", "flat", "checker", "stripes", "gradient")  # the last four spread the random nets' outputs
# Random-init nets separate images only weakly, so most images sit within bf16 rounding of a keep/drop boundary and "decisions identical"
# would be a coin toss.  The augmentation images (candidate index c -> seed 700 + c, SIZES[c % 3], KINDS[(c // 3) % 7]) and the source labels
# below were SELECTED by tests/golden/make_filter_json_golden.py --select from the reference nets' own fp32 outputs so that every decision
# has a margin the bf16 forward does not flip (CLIP |l0 - max other| > 0.03, classifier |logit[label] - top-k boundary| > 0.27) while both
# outcomes of both filters occur.  Regenerate with --select if the synthetic generators or seeds change.
AUG_CANDIDATES = [20, 2, 39, 5, 41, 11, 62, 13, 63, 14, 88, 15, 89, 17, 123, 21, 129, 24, 146, 26, 151, 31, 166, 32]
LABELS = [3, 3, 4, 7, 3, 4, 3, 3, 4, 3, 3, 4]
PROMPTS = ["an airplane on a runway at dusk", "an airplane above the clouds, a painting of van gogh", "an airplane parked near a hangar"]


def source_names():
    return [f"syn_{i}.png" for i in range(N_SOURCES)]


def labels():
    return list(LABELS) if LABELS is not None else [(5 * i + 3) % NUM_CLASSES for i in range(N_SOURCES)]


def candidate_image(c: int):
    from saspa_aug_b200.synthetic import synthetic_source

    h, w = SIZES[c % len(SIZES)]
    kind = KINDS[(c // 3) % len(KINDS)]
    if kind in ("blobs", "smooth", "noise"):
        return synthetic_source(700 + c, h, w, kind=kind)
    rng = np.random.default_rng(700 + c)
    col = rng.integers(0, 256, size=(2, 3))
    yy, xx = np.mgrid[0:h, 0:w]
    if kind == "flat":
        m = np.zeros((h, w))
    elif kind == "checker":
        p = int(rng.integers(4, 24))
        m = ((yy // p + xx // p) % 2).astype(np.float64)
    elif kind == "stripes":
        p = int(rng.integers(3, 16))
        m = ((xx if c % 2 else yy) // p % 2).astype(np.float64)
    else:  # gradient
        m = (yy / h + xx / w) / 2.0
    img = col[0][None, None, :] * (1.0 - m[..., None]) + col[1][None, None, :] * m[..., None]
    return np.ascontiguousarray(img.round().astype(np.uint8))


def aug_candidate(k: int) -> int:
    """Candidate index of augmentation k = source * NUM_PER_IMAGE + j."""
    return AUG_CANDIDATES[k] if AUG_CANDIDATES is not None else k


def aug_name(stem: str, prompt: str, i: int) -> str:
    return f"{stem[:40]}_prompt_{prompt.replace('/', '-')}_{i}.png"  # run_aug.py:429


def build(root: str):
    """-> (dataset root with images/, augmentation folder .../images).  Layout mirrors run_aug.py:692."""
    from PIL import Image

    from saspa_aug_b200.synthetic import synthetic_source

    root = Path(root)
    src_dir = root / "images"
    out_dir = root / "aug_data" / "controlnet" / "sd_v1.5" / "canny" / "gpt-meta_class_seed_1" / "images"
    src_dir.mkdir(parents=True, exist_ok=True)
    out_dir.mkdir(parents=True, exist_ok=True)
    for i, name in enumerate(source_names()):
        Image.fromarray(synthetic_source(i, 160, 160)).save(src_dir / name)
        stem = Path(name).stem
        Image.fromarray(synthetic_source(i, 160, 160)).save(out_dir / f"{stem}_source.png")
        if i < 10:
            Image.fromarray(np.zeros((160, 160, 3), np.uint8)).save(out_dir / f"{stem}_control.png")
        for j in range(NUM_PER_IMAGE):
            k = i * NUM_PER_IMAGE + j
            Image.fromarray(candidate_image(aug_candidate(k))).save(out_dir / aug_name(stem, PROMPTS[k % len(PROMPTS)], j))
    Image.fromarray(synthetic_source(999, 160, 160)).save(out_dir / "zzz_leftover_0.png")  # matches no source stem
    good = out_dir / aug_name("syn_7", "a truncated file", 5)
    Image.fromarray(synthetic_source(998, 160, 160)).save(good)
    data = good.read_bytes()
    good.write_bytes(data[: len(data) // 3])  # corrupt: deleted by check_folder_of_images_with_pil before matching
    return str(root), str(out_dir)


def relativize(obj, root: str):
    """JSON bodies hold absolute paths: store them relative to the fixture root."""
    if isinstance(obj, dict):
        return {k: relativize(v, root) for k, v in obj.items()}
    if isinstance(obj, list):
        return [relativize(v, root) for v in obj]
    if isinstance(obj, str):
        return obj.replace(root.rstrip("/") + "/", "{ROOT}/")
    return obj
