"""GPU integration: sharded driver -> PNGs -> batched filter -> aug JSON (layout consumed by the reference's
AugWrapperDataset, fgvc/datasets/aug_wrapper_dataset.py:106-186), resume semantics, generate_canny drop-in."""
import json
import os
from pathlib import Path

import numpy as np
import pytest
import torch

from oracle import clib
from saspa_aug_b200 import filtering, run_aug
from saspa_aug_b200.datasets import SyntheticUtils
from saspa_aug_b200.synthetic import synthetic_source

pytestmark = pytest.mark.gpu


def test_generate_canny_dropin(cuda_device):
    from PIL import Image

    img = synthetic_source(3)
    out = run_aug.generate_canny(Image.fromarray(img), 120, 200, 512)
    assert out.mode == "RGB" and out.size == (512, 512)
    assert np.array_equal(np.array(out), np.repeat(clib.canny(img, 120, 200)[..., None], 3, 2))
    gray = run_aug.generate_canny(img[..., 0], 120, 200, 512)  # HWC3 on a 2-D input
    assert np.array_equal(np.array(gray)[..., 0], clib.canny(np.repeat(img[..., :1], 3, 2), 120, 200))


def test_driver_filter_json_roundtrip(cuda_device, tmp_path):
    ds = SyntheticUtils(root=str(tmp_path / "ds"), n_images=6, size=(128, 128)).materialize()
    cfg = run_aug.AugConfig(BASE_MODEL="tiny", RESOLUTION=128, NUM_INFERENCE_STEPS=3, MICRO_BATCH=4, USE_ARTISTIC_PROMPTS=True).apply_dataset_rules()
    pipe = run_aug.init_pipeline("tiny", "canny", cfg.SDEDIT, sampler="ddim")
    out_dir = run_aug.output_folder(str(tmp_path / "ds"), cfg)
    prompts = [f"an airplane flying over landscape {i}." for i in range(20)]
    written = run_aug.generate(cfg, ds, pipe, prompts, out_dir)
    assert len(written) == 12 and all(os.path.exists(p) for _, _, p in written)
    names = os.listdir(out_dir)
    assert sum("_source." in n for n in names) == 6 and sum("_control." in n for n in names) == 6
    # resume: nothing is regenerated, same file list
    mt = {p: os.path.getmtime(p) for _, _, p in written}
    again = run_aug.generate(cfg, ds, pipe, prompts, out_dir)
    assert sorted(p for _, _, p in again) == sorted(mt) and all(os.path.getmtime(p) == t for p, t in mt.items())
    # filter + JSON
    json_path, det = filtering.create_json_of_image_name_to_augmented_images_paths("synthetic", out_dir, semantic_filtering=1, model_confidence_based_filtering=1,
                                                                                  init_log=False, ds_utils=ds, return_details=True)
    assert json_path == str(Path(out_dir).parent / "semantic_filtering-model_confidence_based_filtering_top_10_classes-aug.json")
    d = json.load(open(json_path))
    assert list(d) == [Path(p).name for p in ds.original_images_paths]
    files = {str(Path(out_dir) / n) for n in names}
    for k, v in d.items():
        assert isinstance(v, list) and all(p in files and "_source." not in p and "_control." not in p for p in v)
    kept = sum(len(v) for v in d.values())
    assert kept == int((det["in_topk"] & det["semantic"]).sum())
    # per-item RNG: a different partitioning reproduces the same pixels
    from PIL import Image

    cfg2 = run_aug.AugConfig(BASE_MODEL="tiny", RESOLUTION=128, NUM_INFERENCE_STEPS=3, MICRO_BATCH=2, USE_ARTISTIC_PROMPTS=True).apply_dataset_rules()
    out2 = str(tmp_path / "shard1" / "images")
    w2 = run_aug.generate(cfg2, ds, pipe, prompts, out2, rank=1, world=2)
    assert sorted({i for i, _, _ in w2}) == [1, 3, 5]
    for index, i, p in w2:
        ref = [q for a, b, q in written if (a, b) == (index, i)][0]
        a, b = np.asarray(Image.open(p)).astype(int), np.asarray(Image.open(ref)).astype(int)
        # same seeds/prompts/noise => bit-identical pixels: every kernel's per-item arithmetic is independent of what else shares the
        # micro-batch (tile shapes follow per-image geometry, reductions run in a fixed order), SURVEY.md 8e "results independent of W"
        assert np.array_equal(a, b), (index, i, np.abs(a - b).mean(), np.abs(a - b).max())


@pytest.mark.parametrize("base_model", ["tiny_xl", "tiny_blip"])
def test_driver_other_base_models(cuda_device, tmp_path, base_model):
    """The sharded loop with the SD-XL(-turbo) and BLIP-Diffusion pipelines (added conditioning / subject embeddings come from the
    pipeline's own _encode_call): files are written under the reference's naming, images are non-degenerate, and a generated image
    equals the one the reference-style single call (pass_thorugh_pipe with the same per-item generator) produces."""
    from PIL import Image

    ds = SyntheticUtils(root=str(tmp_path / "ds"), n_images=3, size=(128, 128)).materialize()
    cfg = run_aug.AugConfig(BASE_MODEL=base_model, RESOLUTION=128, NUM_INFERENCE_STEPS=3, MICRO_BATCH=4)
    if base_model == "tiny_xl":
        cfg.GUIDANCE_SCALE = 0.0  # sd_xl-turbo rule (run_aug.py:567-570)
    pipe = run_aug.init_pipeline(base_model, "canny", 0, sampler="ddim")
    out_dir = run_aug.output_folder(str(tmp_path / "ds"), cfg)
    prompts = [f"an airplane parked on wet tarmac {i}." for i in range(8)]
    written = run_aug.generate(cfg, ds, pipe, prompts, out_dir)
    assert len(written) == 6 and all(os.path.exists(p) for _, _, p in written)
    for _, _, p in written:
        a = np.asarray(Image.open(p))
        assert a.shape == (128, 128, 3) and a.std() > 1.0
    # one item through the reference's per-image call path with the same per-item seed
    index, i, path = written[0]
    src = Image.open(ds.original_images_paths[index]).convert("RGB")
    canny = run_aug.generate_canny(src, cfg.LOW_THRESHOLD_CANNY, cfg.HIGH_THRESHOLD_CANNY, cfg.RESOLUTION)
    prompt = run_aug.sample_prompts(prompts, ds.original_images_paths, cfg, ds)[index][i]
    g = torch.Generator().manual_seed(run_aug.item_seed(cfg.SEED, index, i))
    one = run_aug.pass_thorugh_pipe("blip_diffusion" if "blip" in base_model else "sd_xl-turbo", pipe, prompt, src, 0, cfg.SDEDIT_STRENGTH,
                                    cfg.NUM_INFERENCE_STEPS, g, cfg.GUIDANCE_SCALE, cfg.CONTROLNET_CONDITIONING_SCALE, control_image=canny,
                                    blip_src_category=ds.meta_class, blip_target_category=ds.meta_class)
    a, b = np.asarray(one).astype(int), np.asarray(Image.open(path)).astype(int)
    assert np.array_equal(a, b), (np.abs(a - b).mean(), np.abs(a - b).max())  # batch-1 call == the item inside a micro-batch of 4


def test_reference_order_rng_matches_sequentially_threaded_generator(cuda_device, tmp_path):
    """RNG_MODE = "reference_order": the sharded loop (here as rank 1 of 2 and rank 0 of 2) produces the images a reference-style
    single-process loop produces when ONE generator (torch.manual_seed(SEED), fp16 draws after .to(device, float16)) is threaded through
    per-image pipeline calls in dataset order (run_aug.py:324, :464)."""
    from PIL import Image

    ds = SyntheticUtils(root=str(tmp_path / "ds"), n_images=4, size=(128, 128)).materialize()
    cfg = run_aug.AugConfig(BASE_MODEL="tiny", RESOLUTION=128, NUM_INFERENCE_STEPS=3, MICRO_BATCH=4, RNG_MODE="reference_order")
    pipe = run_aug.init_pipeline("tiny", "canny", 0, sampler="ddim").to("cuda", torch.float16)
    prompts = [f"an airplane above the clouds {i}." for i in range(8)]
    sampled = run_aug.sample_prompts(prompts, ds.original_images_paths, cfg, ds)
    g = torch.manual_seed(cfg.SEED)  # the reference's generator is the global CPU generator
    want = {}
    for index, p in enumerate(ds.original_images_paths):
        src = Image.open(p).convert("RGB")
        canny = run_aug.generate_canny(src, cfg.LOW_THRESHOLD_CANNY, cfg.HIGH_THRESHOLD_CANNY, cfg.RESOLUTION)
        for i, prompt in enumerate(sampled[index]):
            want[(index, i)] = np.asarray(run_aug.pass_thorugh_pipe("sd_v1.5", pipe, prompt, src, 0, cfg.SDEDIT_STRENGTH, cfg.NUM_INFERENCE_STEPS, g,
                                                                    cfg.GUIDANCE_SCALE, cfg.CONTROLNET_CONDITIONING_SCALE, control_image=canny)).astype(int)
    got = {}
    for rank in (1, 0):
        for index, i, path in run_aug.generate(cfg, ds, pipe, prompts, str(tmp_path / f"out{rank}" / "images"), rank=rank, world=2):
            got[(index, i)] = np.asarray(Image.open(path)).astype(int)
    assert set(got) == set(want)
    for k in want:
        d = np.abs(got[k] - want[k])
        assert d.max() == 0, (k, d.mean(), d.max())
    # and it differs from the per-item mode (different noise)
    cfg2 = run_aug.AugConfig(BASE_MODEL="tiny", RESOLUTION=128, NUM_INFERENCE_STEPS=3, MICRO_BATCH=4)
    other = run_aug.generate(cfg2, ds, pipe, prompts, str(tmp_path / "per_item" / "images"))
    a = np.asarray(Image.open(other[0][2])).astype(int)
    assert np.abs(a - want[(other[0][0], other[0][1])]).mean() > 2.0


_RANK_WORKER = r'''
import json, os, sys
sys.path.insert(0, sys.argv[1])
root = sys.argv[2]
import torch
from saspa_aug_b200 import run_aug
from saspa_aug_b200.datasets import SyntheticUtils

sizes = [(128, 128), (96, 128), (128, 96), (128, 128)]  # mixed aspect ratios: 128x128, 128x192 and 192x128 after resize_image(., 128)
ds = SyntheticUtils(root=root, n_images=10, sizes=sizes, n_classes=12, clip_seed=813)
if int(os.environ.get("RANK", "0")) == 0:
    ds.materialize()
cfg = run_aug.AugConfig(BASE_MODEL="tiny", RESOLUTION=128, NUM_INFERENCE_STEPS=3, MICRO_BATCH=4)
if int(os.environ.get("WORLD_SIZE", "1")) > 1:
    import torch.distributed as dist
    dist.init_process_group("gloo")
    dist.barrier()
prompts = [f"an airplane over terrain {i}." for i in range(12)]
json_path, stats = run_aug.run_sharded(cfg, ds, prompts, root, device="cuda:0")
print("STATS", json.dumps(stats))
if json_path:
    print("JSON", json_path)
'''


def test_one_rank_and_two_rank_runs_write_the_same_json_and_pixels(cuda_device, tmp_path):
    """results are independent of the world size.  The same job (10 mixed-aspect sources x 2, generate -> verify -> filter ->
    gather -> rank-0 JSON) runs once as a single process and once as two ranks (one process per rank, gloo for the one collective so
    both can share this GPU; NCCL needs a device per rank): every PNG is byte-identical and the aug JSON equal up to the root folder."""
    import json
    import subprocess
    import sys

    repo = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    w = tmp_path / "worker.py"
    w.write_text(_RANK_WORKER)
    env = dict(os.environ, SASPA_DIST_BACKEND="gloo")
    r1 = subprocess.run([sys.executable, str(w), repo, str(tmp_path / "one")], capture_output=True, text=True, timeout=600, env=env)
    assert r1.returncode == 0, r1.stdout[-3000:] + r1.stderr[-3000:]
    r2 = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1", "--master-port", "29621",
                         str(w), repo, str(tmp_path / "two")], capture_output=True, text=True, timeout=900, env=env)
    assert r2.returncode == 0, r2.stdout[-3000:] + r2.stderr[-3000:]
    j1 = [ln.split(" ", 1)[1] for ln in r1.stdout.splitlines() if ln.startswith("JSON ")]
    j2 = [ln.split(" ", 1)[1] for ln in r2.stdout.splitlines() if ln.startswith("JSON ")]
    assert len(j1) == 1 and len(j2) == 1  # rank 0 only
    d1 = json.loads(open(j1[0]).read().replace(str(tmp_path / "one"), "{ROOT}"))
    d2 = json.loads(open(j2[0]).read().replace(str(tmp_path / "two"), "{ROOT}"))
    assert list(d1) == list(d2) and len(d1) == 10
    assert {k: sorted(v) for k, v in d1.items()} == {k: sorted(v) for k, v in d2.items()}  # value order = os.listdir order of each folder
    kept = sum(len(v) for v in d1.values())
    print(f"kept {kept} of 20; stats one-rank: {[ln for ln in r1.stdout.splitlines() if ln.startswith('STATS')]}")
    f1, f2 = os.path.dirname(j1[0]) + "/images", os.path.dirname(j2[0]) + "/images"
    names = sorted(os.listdir(f1))
    assert names == sorted(os.listdir(f2)) and sum("_prompt_" in n for n in names) == 20
    from PIL import Image

    bad = []
    for n in names:
        if open(os.path.join(f1, n), "rb").read() != open(os.path.join(f2, n), "rb").read():
            a, b = np.asarray(Image.open(os.path.join(f1, n))).astype(int), np.asarray(Image.open(os.path.join(f2, n))).astype(int)
            bad.append((n, a.shape, int(np.abs(a - b).max()), float((a != b).mean())))
    assert not bad, bad
