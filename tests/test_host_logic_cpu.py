"""CPU: host-side mirror of the reference interface -- JSON naming / layout, file naming, matching rule,
sharding, and the world-size-2 gather (gloo)."""
import json
import os
import subprocess
import sys
import tempfile
from pathlib import Path

import numpy as np
import pytest

from saspa_aug_b200 import filtering, run_aug

G = os.path.join(os.path.dirname(__file__), "golden")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_json_names_match_reference_golden():
    names = json.load(open(os.path.join(G, "filter_golden.json")))["json_names"]  # produced by the reference's get_aug_json_path
    f = filtering.get_aug_json_path
    assert f("/x/y/images", semantic_filtering=1, model_confidence_based_filtering=1) == names["sem+conf"]
    assert f("/x/y/images") == names["none"]
    assert f("/x/y/images", model_confidence_based_filtering=True, conf_top_k=5) == names["conf_top5"]
    assert f("/x/y/images", lpips_min=0.1, lpips_max=0.7, clip_filtering="per_class", clip_filtering_discount=2, semantic_filtering=True,
             alia_conf_filtering=True) == names["all"]
    assert names["sem+conf"] == "/x/y/semantic_filtering-model_confidence_based_filtering_top_10_classes-aug.json"


def test_matching_rule_and_exclusions():
    srcs = ["/d/images/0001234.jpg", "/d/images/000123.jpg", "/d/images/" + "a" * 50 + ".png"]
    files = ["0001234_prompt_a photo_0.png", "0001234_source.png", "0001234_control.png", "000123_prompt_x_1.png", "a" * 40 + "_prompt_p_0.png",
             "subject_0.png"]
    files = [f for f in files if not any(s in f for s in filtering.SUBSTRINGS_TO_EXCLUDE)]
    m = filtering.match_augmentations(srcs, files, "/out/images")
    assert list(m) == ["0001234.jpg", "000123.jpg", "a" * 50 + ".png"]
    assert m["0001234.jpg"] == ["/out/images/0001234_prompt_a photo_0.png"]
    # substring semantics of the reference: "000123" also matches the "0001234_..." file
    assert m["000123.jpg"] == ["/out/images/0001234_prompt_a photo_0.png", "/out/images/000123_prompt_x_1.png"]
    assert m["a" * 50 + ".png"] == ["/out/images/" + "a" * 40 + "_prompt_p_0.png"]
    assert filtering.get_dict_of_value_counts(m) == {1: 2, 2: 1}


def test_file_and_folder_naming():
    cfg = run_aug.AugConfig(USE_ARTISTIC_PROMPTS=True, PROMPT_WITH_SUB_CLASS=True)
    assert run_aug.output_folder("/data/planes", cfg) == \
        "/data/planes/aug_data/controlnet/sd_v1.5/canny/gpt-meta_class_prompt_w_sub_class_artistic_prompts_p_0.5_seed_1/images"
    blip = run_aug.AugConfig(BASE_MODEL="blip_diffusion", USE_CAMERA_VARIATIONS_PROMPTS=True, PROMPT_WITH_SUB_CLASS=False)
    assert run_aug.output_folder("/d", blip) == \
        "/d/aug_data/controlnet/blip_diffusion/canny/gpt-meta_class_camera_variations_p_0.5_style_img_from_diff_img_seed_1/images"
    cub = run_aug.AugConfig(DATASET="cub").apply_dataset_rules()  # run_aug.py:564-571
    assert (cub.BASE_MODEL, cub.GUIDANCE_SCALE, cub.NUM_INFERENCE_STEPS, cub.NEGATIVE_PROMPT) == ("sd_xl-turbo", 0.0, 2, None)
    assert run_aug.AugConfig(DATASET="compcars-parts").apply_dataset_rules().NUM_INFERENCE_STEPS == 50  # "cars" in DATASET.lower()
    cfg2 = run_aug.AugConfig(SDEDIT=1, SDEDIT_STRENGTH=0.5, CONTROLNET=None)
    assert "/aug_data/regular/sd_v1.5-SDEdit_strength_0.5/None/" in run_aug.output_folder("/d", cfg2)
    assert run_aug.aug_file_name("x" * 60, "a/b photo", 1) == "x" * 40 + "_prompt_a-b photo_1.png"
    assert run_aug.AugConfig(BASE_MODEL="sd_xl-turbo").apply_dataset_rules().NUM_INFERENCE_STEPS == 2
    with pytest.raises(AssertionError):
        run_aug.AugConfig(SDEDIT=1, SDEDIT_STRENGTH=0.01).apply_dataset_rules()


def test_resize_image_and_hwc3_match_reference_semantics():
    img = np.zeros((375, 500, 3), np.uint8)
    assert run_aug.resize_image(img, 512).shape == (512, 704, 3)
    sq = np.random.default_rng(0).integers(0, 255, (512, 512, 3), dtype=np.uint8)
    assert run_aug.resize_image(sq, 512) is sq
    g = np.full((4, 4), 7, np.uint8)
    assert run_aug.HWC3(g).shape == (4, 4, 3)
    rgba = np.zeros((2, 2, 4), np.uint8)
    assert (run_aug.HWC3(rgba) == 255).all()  # transparent over white


def test_prompt_sampling_is_partition_independent():
    from saspa_aug_b200.datasets import SyntheticUtils

    cfg = run_aug.AugConfig()  # reference defaults: artistic suffix on even i (sd_v1.5), sub-class inserted before the meta class
    assert cfg.USE_ARTISTIC_PROMPTS and cfg.PROMPT_WITH_SUB_CLASS
    ds = SyntheticUtils(root="/nonexistent", n_images=12)
    prompts = [f"an airplane number {i}." for i in range(50)]
    a = run_aug.sample_prompts(prompts, ds.original_images_paths, cfg, ds)
    b = run_aug.sample_prompts(prompts, ds.original_images_paths, cfg, ds)
    assert a == b and len(a) == 12 and all(len(x) == 2 for x in a)
    assert all(not p.endswith(".") for x in a for p in x) and all(", a painting of" in x[0] and "," not in x[1] for x in a)
    assert all(f"an class_{k % 100} airplane number" in p for k, x in enumerate(a) for p in x)
    plain = run_aug.sample_prompts(prompts, 12, run_aug.AugConfig(PROMPT_WITH_SUB_CLASS=False))  # anonymous sources: only without per-source rewriting
    assert [[p.replace(f"class_{k % 100} ", "") for p in x] for k, x in enumerate(a)] == plain
    with pytest.raises(ValueError):
        run_aug.sample_prompts(prompts, 12, cfg)
    parts = [run_aug.shard_indices(12, r, 4) for r in range(4)]
    assert sorted(i for p in parts for i in p) == list(range(12)) and parts[1] == [1, 5, 9]
    assert run_aug.item_seed(1, 5, 0) != run_aug.item_seed(1, 5, 1)


_WORKER = r'''
import os, sys, numpy as np, torch.distributed as dist
sys.path.insert(0, sys.argv[1])
from saspa_aug_b200 import run_aug
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
dist.init_process_group("gloo")
mine = run_aug.shard_indices(11, rank, world)
rec = np.array([[i, i % 2, (i * 7) % 3] for i in mine], dtype=np.int32).reshape(-1, 3)
allrec = run_aug.gather_records(rec, rank, world)
if rank == 0:
    allrec = allrec[np.argsort(allrec[:, 0])]
    assert allrec.shape == (11, 3) and (allrec[:, 0] == np.arange(11)).all() and (allrec[:, 2] == (np.arange(11) * 7) % 3).all()
    print("GATHER_OK")
dist.destroy_process_group()
'''


def test_world_size_2_gather_gloo():
    with tempfile.TemporaryDirectory() as d:
        w = os.path.join(d, "worker.py")
        open(w, "w").write(_WORKER)
        r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
                            "--master-port", "29611", w, ROOT], capture_output=True, text=True, timeout=240)
        assert r.returncode == 0 and "GATHER_OK" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]


def test_json_writer_from_gathered_decisions(tmp_path):
    """Rank 0 of the sharded driver writes the aug JSON from the gathered filter records through the same reference-compatible
    writer (no networks, no GPU): matching by stem substring, dataset key order, listdir value order, confidence filter before the
    semantic one, file name from the enabled flags."""
    import json

    from PIL import Image

    from saspa_aug_b200 import filtering
    from saspa_aug_b200.datasets import SyntheticUtils

    ds = SyntheticUtils(root=str(tmp_path / "ds"), n_images=4, size=(32, 32)).materialize()
    cfg = run_aug.AugConfig()
    out_dir = run_aug.output_folder(str(tmp_path / "ds"), cfg)
    os.makedirs(out_dir)
    decisions, expect = {}, {}
    for index, p in enumerate(ds.original_images_paths):
        stem = os.path.splitext(os.path.basename(p))[0]
        Image.new("RGB", (8, 8)).save(os.path.join(out_dir, f"{stem}_source.png"))
        expect[os.path.basename(p)] = []
        for i in range(3):
            path = os.path.join(out_dir, run_aug.aug_file_name(stem, f"an airplane, take {i}", i))
            Image.new("RGB", (8, 8), (index * 40, i * 60, 0)).save(path)
            decisions[(os.path.basename(p), path)] = ((index + i) % 2, int(i != 1))
    for f in os.listdir(out_dir):  # listdir order is the value order
        if "_source." in f or "_control." in f:
            continue
        name = [k for k in expect if os.path.splitext(k)[0] in f][0]
        a, b = decisions[(name, os.path.join(out_dir, f))]
        if a and b:
            expect[name].append(os.path.join(out_dir, f))
    jp = filtering.create_json_of_image_name_to_augmented_images_paths("synthetic", out_dir, semantic_filtering=True, model_confidence_based_filtering=True,
                                                                       init_log=False, ds_utils=ds, decisions=decisions)
    assert os.path.basename(jp) == "semantic_filtering-model_confidence_based_filtering_top_10_classes-aug.json"
    got = json.load(open(jp))
    assert list(got) == [os.path.basename(p) for p in ds.original_images_paths] and got == expect
    # a matched pair nobody scored (a file left by an earlier run, a substring cross-match): handed to `missing_decisions`; without
    # one it is logged and left out -- never a KeyError that would abort the JSON on rank 0
    dropped = [k for k, v in decisions.items() if v == (1, 1)][0]
    missing = {k: v for k, v in decisions.items() if k != dropped}
    asked = []

    def score(pairs):
        asked.extend(pairs)
        return {pr: (1, 1) for pr in pairs}

    jp = filtering.create_json_of_image_name_to_augmented_images_paths("synthetic", out_dir, semantic_filtering=True, model_confidence_based_filtering=True,
                                                                       init_log=False, ds_utils=ds, decisions=missing, missing_decisions=score)
    assert asked == [dropped] and json.load(open(jp)) == expect
    jp = filtering.create_json_of_image_name_to_augmented_images_paths("synthetic", out_dir, semantic_filtering=True, model_confidence_based_filtering=True,
                                                                       init_log=False, ds_utils=ds, decisions=missing)
    got = json.load(open(jp))
    assert dropped[1] not in got[dropped[0]] and sum(len(v) for v in got.values()) == sum(len(v) for v in expect.values()) - 1


def test_reference_order_noise_replays_the_global_generator():
    """RNG_MODE = "reference_order": every rank replays the ONE generator the reference threads through all pipeline calls
    (torch.manual_seed(SEED) -> global CPU generator, fp16 draws on the CPU, posterior noise first under SDEdit) and keeps its own
    items; the union over ranks equals a sequential single-process replay, skipped (already existing) items draw nothing."""
    import torch

    sizes = [(512, 512), (512, 704), (576, 512), (512, 512), (512, 512)]
    for sdedit in (0, 1):
        cfg = run_aug.AugConfig(SEED=7, SDEDIT=sdedit)
        skipped = {(1, 1), (3, 0)}
        draws = [[run_aug.PromptDraw(p, None, (k, i) in skipped) for i, p in enumerate("ab")] for k in range(len(sizes))]
        # literal sequential replay, the way run_aug.py + diffusers consume the global generator
        g = torch.manual_seed(cfg.SEED)
        want = {}
        for index, (H, W) in enumerate(sizes):
            for i in range(2):
                if (index, i) in skipped:
                    continue
                shape = (1, 4, H // 8, W // 8)
                post = torch.randn(shape, generator=g, dtype=torch.float16) if sdedit else None
                want[(index, i)] = (torch.randn(shape, generator=g, dtype=torch.float16), post)
        got = {}
        for rank in range(3):
            mine = set(run_aug.shard_indices(len(sizes), rank, 3))
            part = run_aug.reference_order_noise(cfg, sizes, draws, mine)
            assert all(k[0] in mine for k in part) and not (set(part) & set(got))
            got.update(part)
        assert set(got) == set(want)
        for k, (n, p) in want.items():
            assert got[k][0].dtype == torch.float32 and torch.equal(got[k][0], n.float())
            assert (got[k][1] is None) if p is None else torch.equal(got[k][1], p.float())
    assert run_aug.resized_hw(300, 400, 512)[:2] == (512, 704) and run_aug.resized_hw(512, 512, 512)[:2] == (512, 512)


def test_pipeline_draws_noise_in_the_dtype_given_to_to():
    """`pipe.to(DEVICE, torch.float16)` (run_aug.py:323) makes diffusers draw latent noise from the caller's generator in fp16; the
    drop-in remembers the dtype so the same generator state produces the same noise (an fp32 draw consumes the generator differently)."""
    import torch

    from saspa_aug_b200.pipelines import SaspaControlNetPipeline

    p = SaspaControlNetPipeline.__new__(SaspaControlNetPipeline)
    p.noise_dtype = torch.float32
    assert p.to("cuda") is p and p.noise_dtype == torch.float32
    assert p.to("cuda", torch.float16) is p and p.noise_dtype == torch.float16
    assert p.to(torch_dtype=torch.bfloat16).noise_dtype == torch.bfloat16


from oracle import ref_import  # noqa: E402


@pytest.mark.skipif(not ref_import.available(), reason="/root/reference not present (GPU box)")
def test_host_mirrors_against_the_reference_functions_live():
    """The host-side mirrors of all_utils/utils.py run against the reference's OWN functions (imported from /root/reference, this
    container only): resize_image (:58-79, LANCZOS4 up / AREA down, x64 rounding, 1.2 MP cap), HWC3 (:39-55, gray / RGBA),
    get_dict_of_value_counts (:468-482) and get_aug_json_path (:194-218) on a flag sweep."""
    import itertools

    ru = ref_import.import_reference_utils()
    rng = np.random.default_rng(5)
    for h, w in [(300, 400), (512, 512), (1600, 1200), (720, 1280), (100, 333), (2000, 3000)]:
        img = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
        for res in (512, 768):
            want, got = ru.resize_image(img, res), run_aug.resize_image(img, res)
            assert want.shape == got.shape and np.array_equal(want, got), (h, w, res)
            assert run_aug.resized_hw(h, w, res)[:2] == want.shape[:2]
    gray = rng.integers(0, 256, (37, 41), dtype=np.uint8)
    rgba = rng.integers(0, 256, (16, 24, 4), dtype=np.uint8)
    one = rng.integers(0, 256, (9, 9, 1), dtype=np.uint8)
    for x in (gray, rgba, one, rng.integers(0, 256, (8, 8, 3), dtype=np.uint8)):
        assert np.array_equal(ru.HWC3(x), run_aug.HWC3(x))
    d = {"a": [1, 2], "b": [], "c": [3], "d": [4, 5]}
    assert dict(ru.get_dict_of_value_counts_image_name_to_num_aug_images(d)) == filtering.get_dict_of_value_counts(d)
    for lmin, lmax, cf, sem, conf, topk, hi, alia in itertools.product((None, 0.1), (None, 0.7), (False, "per_class"), (False, True), (False, True),
                                                                       (10, 5), (None, 0.9), (False, True)):
        if cf and conf:
            continue
        kw = dict(lpips_min=lmin, lpips_max=lmax, clip_filtering=cf, clip_filtering_discount=2, semantic_filtering=sem,
                  model_confidence_based_filtering=conf, conf_top_k=topk, filter_confidence_higher_than=hi, alia_conf_filtering=alia)
        assert ru.get_aug_json_path("/data/x/aug/images", **kw) == filtering.get_aug_json_path("/data/x/aug/images", **kw), kw


_SHARDED_WORKER = r'''
import json, os, sys
import numpy as np
sys.path.insert(0, sys.argv[1])
root = sys.argv[2]
from PIL import Image
from saspa_aug_b200 import run_aug
from saspa_aug_b200.datasets import SyntheticUtils

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
ds = SyntheticUtils(root=root, n_images=7, size=(32, 32))
if rank == 0:
    ds.materialize()
import torch.distributed as dist
dist.init_process_group("gloo")
dist.barrier()
cfg = run_aug.AugConfig(NUM_PER_IMAGE=2)
prompts = [f"an airplane, variation {i}." for i in range(10)]


def fake_generate(cfg, ds_utils, pipe, prompts, out_dir, rank=0, world=1):  # stands in for the GPU generation stage: same names, same shard
    os.makedirs(out_dir, exist_ok=True)
    sampled = run_aug.sample_prompts(prompts, ds_utils.original_images_paths, cfg, ds_utils)
    written = []
    for index in run_aug.shard_indices(len(ds_utils.original_images_paths), rank, world):
        stem = os.path.splitext(os.path.basename(ds_utils.original_images_paths[index]))[0]
        for i, prompt in enumerate(sampled[index]):
            path = os.path.join(out_dir, run_aug.aug_file_name(stem, prompt, i))
            Image.new("RGB", (8, 8), (index * 30, i * 100, rank * 100)).save(path)
            written.append((index, i, path))
    return written


def fake_filter(cfg, ds_utils, written, device=None, filter_models=None):  # deterministic decisions: (index + i) % 3 != 0 passes top-k, i == 0 is semantic
    rec = np.zeros((len(written), 4), np.int32)
    for k, (index, i, _) in enumerate(written):
        rec[k] = (index, i, int((index + i) % 3 != 0), int(i == 0 or index % 2 == 0))
    return rec


json_path, stats = run_aug.run_sharded(cfg, ds, prompts, root, generate_fn=fake_generate, filter_fn=fake_filter, device="cpu")
assert stats["generated"] == 2 * len(run_aug.shard_indices(7, rank, world))
if rank == 0:
    d = json.load(open(json_path))
    assert list(d) == [os.path.basename(p) for p in ds.original_images_paths]
    sampled = run_aug.sample_prompts(prompts, ds.original_images_paths, cfg, ds)
    kept = 0
    for index, p in enumerate(ds.original_images_paths):
        stem = os.path.splitext(os.path.basename(p))[0]
        want = {os.path.join(run_aug.output_folder(root, cfg), run_aug.aug_file_name(stem, sampled[index][i], i)) for i in range(2)
                if (index + i) % 3 != 0 and (i == 0 or index % 2 == 0)}
        assert set(d[os.path.basename(p)]) == want, (index, d[os.path.basename(p)], want)
        kept += len(want)
    assert stats["records"] == 14 and stats["kept"] == kept and json_path.endswith("semantic_filtering-model_confidence_based_filtering_top_10_classes-aug.json")
    print("SHARDED_OK", kept)
else:
    assert json_path is None
dist.destroy_process_group()
'''


def test_world_size_2_run_sharded_host_logic_gloo(tmp_path):
    """The end-to-end rank entry with the two GPU stages replaced by stand-ins: interleaved sharding, barrier, ONE gather of the
    fixed-size records, rank 0 rebuilding the file names from the rank-independent prompt draw and writing the reference-layout JSON."""
    w = tmp_path / "worker.py"
    w.write_text(_SHARDED_WORKER)
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
                        "--master-port", "29613", str(w), ROOT, str(tmp_path / "ds")], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "SHARDED_OK" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]


@pytest.mark.skipif(not ref_import.available(), reason="/root/reference not present (GPU box)")
def test_reference_consumer_reads_the_json_we_write(tmp_path):
    """The aug JSON is a contract with the reference's training side: its OWN AugWrapperDataset (fgvc/datasets/aug_wrapper_dataset.py:106-186,
    imported from /root/reference) loads the file written by our writer, drops the sources without kept augmentations at ratio 1 and serves
    the kept files."""
    import importlib.util
    import random

    from PIL import Image

    from saspa_aug_b200 import filtering
    from saspa_aug_b200.datasets import SyntheticUtils

    spec = importlib.util.spec_from_file_location("ref_aug_wrapper", os.path.join(ref_import.REFERENCE_ROOT, "fgvc", "datasets", "aug_wrapper_dataset.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)

    ds = SyntheticUtils(root=str(tmp_path / "ds"), n_images=6, size=(32, 32)).materialize()
    cfg = run_aug.AugConfig()
    out_dir = run_aug.output_folder(str(tmp_path / "ds"), cfg)
    os.makedirs(out_dir)
    decisions = {}
    for index, p in enumerate(ds.original_images_paths):
        stem = os.path.splitext(os.path.basename(p))[0]
        for i in range(2):
            path = os.path.join(out_dir, run_aug.aug_file_name(stem, f"an airplane, take {i}", i))
            Image.new("RGB", (8, 8), (index * 40, 200, i * 200)).save(path)
            decisions[(os.path.basename(p), path)] = (int(index % 3 != 0), int(i == 0 or index == 4))  # sources 0 and 3 keep nothing
    jp = filtering.create_json_of_image_name_to_augmented_images_paths("synthetic", out_dir, semantic_filtering=True, model_confidence_based_filtering=True,
                                                                       init_log=False, ds_utils=ds, decisions=decisions)

    class Train(mod.AugWrapperDataset):
        def __init__(self, **kw):
            self._image_files = list(ds.original_images_paths)
            self._labels = [ds.label_of(i) for i in range(len(self._image_files))]
            self.num_classes, self.dataset_name = ds.num_classes, "synthetic"
            super().__init__(root=str(tmp_path / "ds"), split="train", print_func=lambda *a: None, **kw)

    random.seed(0)
    t = Train(aug_json=jp, aug_sample_ratio=1)
    assert len(t) == 4 and sorted(os.path.basename(p) for p in t._image_files) == [os.path.basename(ds.original_images_paths[i]) for i in (1, 2, 4, 5)]
    served = set()
    for rep in range(8):
        for idx in range(len(t)):
            img, label = t[idx]
            assert img.size == (8, 8) and label == t._labels[idx]  # always an augmentation at ratio 1
            served.add(img.getpixel((0, 0)))
    kept = {p for (_, p), (a, b) in decisions.items() if a and b}
    assert served == {Image.open(p).getpixel((0, 0)) for p in kept} and len(kept) == 5
    assert t.times_used_aug_images == 32 and t.times_used_orig_images == 0
