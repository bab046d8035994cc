"""BASELINE config 5: synthetic sources sharded over the GPUs of one box (torchrun, one rank per GPU): generate (SD v1.5 ControlNet-canny
text2img, 20 UniPC steps, CFG) -> verify -> per-rank filter (WSDAN_CAL-R50 top-10 + CLIP semantic check) -> NCCL all-gather of the filter
records -> rank 0 writes the aug JSON; then the filter + gather + JSON again with the other CLIP tower (--clip both: ViT-L/14 first, RN50
second; generation is resumed from the files on disk).  Prints per-rank stage timings, files/s and the JSON statistics.

  python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 tools/run_config5.py --sources 10000 --out gpurun_out/r2_config5.txt

Random-init CLIP towers decide the semantic check almost independently of the image (the 7 prompts' text features dominate), so a seed
either keeps everything or nothing: rank 0 probes a few seeds on the first sources and broadcasts the first one whose semantic check
passes, so that the kept set is not empty and the JSON exercises both outcomes of the confidence filter (labels i % 100: about one
source in ten has its label among the random classifier's ten favourite classes)."""
import argparse
import json
import os
import shutil
import sys
import tempfile
import time
from concurrent.futures import ThreadPoolExecutor

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))  # repo root (python tools/<name>.py)
import numpy as np
import torch
import torch.distributed as dist

from saspa_aug_b200 import run_aug
from saspa_aug_b200.datasets import SyntheticUtils


def materialize_shard(ds, rank, world, threads=8):
    """Every rank writes its own sources (10k synthetic PNGs on one rank would take minutes)."""
    from PIL import Image

    from saspa_aug_b200.synthetic import synthetic_source

    ds.images_path.mkdir(parents=True, exist_ok=True)

    def one(i):
        p = ds.original_images_paths[i]
        if not os.path.exists(p):
            Image.fromarray(synthetic_source(ds.seed_base + i, *ds.size_of(i))).save(p, compress_level=1)

    with ThreadPoolExecutor(max_workers=threads) as ex:
        list(ex.map(one, run_aug.shard_indices(len(ds.original_images_paths), rank, world)))


def pick_clip_seed(ds, device, tries):
    """First CLIP seed (from ds.clip_seed) whose semantic check keeps the probe images; falls back to the last one tried."""
    from saspa_aug_b200.filter_nets import AugmentationFilter
    from saspa_aug_b200.filtering import SEMANTIC_NEGATIVE_PROMPTS
    from saspa_aug_b200.synthetic import synthetic_source

    imgs = torch.from_numpy(np.stack([synthetic_source(9000 + k, kind=("blobs", "noise")[k % 2]) for k in range(8)])).to(device)
    seed0 = ds.clip_seed
    for s in range(seed0, seed0 + tries):
        ds.clip_seed = s
        _, clip, tok = ds.load_filter_models(ds, device)
        flt = AugmentationFilter(None, clip, tok([ds.get_basic_prompt()] + SEMANTIC_NEGATIVE_PROMPTS))
        keep = float(flt(imgs, torch.zeros(8, dtype=torch.int32, device=device))["semantic"].float().mean())
        del flt, clip
        torch.cuda.empty_cache()
        if keep >= 0.5:
            return s, keep
    return ds.clip_seed, keep


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--sources", type=int, default=64)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--clip", default="both", choices=["both", "ViT-L/14", "RN50"])
    ap.add_argument("--micro-batch", type=int, default=32)
    ap.add_argument("--root", default=None)
    ap.add_argument("--out", default=None, help="append the report to this file (rank 0)")
    ap.add_argument("--keep-files", action="store_true")
    ap.add_argument("--seed-tries", type=int, default=1, help="CLIP seeds to probe from the per-tower default (tools/clip_seed_scan.py found them)")
    a = ap.parse_args()
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    root = a.root or os.path.join(tempfile.gettempdir(), "saspa_config5")
    torch.cuda.set_device(local)
    dev = f"cuda:{local}"
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device(dev))
    clips = ["ViT-L/14", "RN50"] if a.clip == "both" else [a.clip]
    ds = SyntheticUtils(root=root, n_images=a.sources, clip_model=clips[0])
    t0 = time.perf_counter()
    materialize_shard(ds, rank, world)
    if world > 1:
        dist.barrier()
    t_mat = time.perf_counter() - t0
    cfg = run_aug.AugConfig(NUM_INFERENCE_STEPS=a.steps, SAMPLER="unipcmultistep", MICRO_BATCH=a.micro_batch).apply_dataset_rules()
    prompts = [f"an airplane on a runway at dusk, variation {i}." for i in range(40)]
    lines = []

    def report(s):
        print(s, flush=True)
        lines.append(s)

    pipe = run_aug.init_pipeline(cfg.BASE_MODEL, cfg.CONTROLNET, cfg.SDEDIT, sampler=cfg.SAMPLER, device=dev)
    for ci, clip_name in enumerate(clips):
        ds.clip_model = clip_name
        # seeds whose random-init text tower lets the basic prompt win for a share of the images (profiles/r2_clip_seed_scan.txt)
        ds.clip_seed = {"RN50": 813, "ViT-L/14": 783}[clip_name]
        seed = torch.zeros(1, dtype=torch.int64, device=dev)
        if rank == 0:
            s, frac = pick_clip_seed(ds, dev, a.seed_tries)
            seed[0] = s
            report(f"CONFIG5 {clip_name}: CLIP seed {s} (semantic check keeps {frac:.2f} of the probe images)")
        if world > 1:
            dist.broadcast(seed, 0)
        ds.clip_seed = int(seed.item())
        run_aug._FILTER_CACHE.clear()
        json_path, stats = run_aug.run_sharded(cfg, ds, prompts, root, pipe=pipe, device=dev)
        allstats = [None] * world
        if world > 1:
            dist.all_gather_object(allstats, stats)
        else:
            allstats = [stats]
        if rank == 0:
            d = json.load(open(json_path))
            kept = sum(len(v) for v in d.values())
            n_files = sum(s["generated"] for s in allstats)
            gen = max(s["generate_s"] for s in allstats)
            flt = max(s["filter_s"] for s in allstats)
            report(f"CONFIG5 world {world} clip {clip_name}: {len(d)} JSON keys, {stats['records']} augmentations filtered, {kept} kept "
                   f"({sum(1 for v in d.values() if v)} sources with at least one) -> {os.path.basename(json_path)}")
            if ci == 0:
                report(f"  sources materialised in {t_mat:.1f} s; generate (incl. PNG encode + write): slowest rank {gen:.1f} s = {n_files / gen:.1f} files/s job-wide")
            else:
                report(f"  generation resumed from disk: slowest rank {gen:.1f} s to re-plan {n_files} existing files")
            report(f"  verify + filter: slowest rank {flt:.1f} s = {n_files / flt:.1f} images/s job-wide; gather {max(s['gather_s'] for s in allstats) * 1e3:.1f} ms; "
                   f"rank-0 JSON {stats['json_s']:.2f} s")
            for s in allstats:
                report(f"    rank {s['rank']}: generated {s['generated']}, generate {s['generate_s']:.1f} s, filter {s['filter_s']:.1f} s, gather {s['gather_s'] * 1e3:.1f} ms")
            shutil.copy(json_path, json_path.replace("-aug.json", f"-aug.{clip_name.replace('/', '_')}.json"))
    if rank == 0 and a.out:
        os.makedirs(os.path.dirname(os.path.abspath(a.out)), exist_ok=True)
        with open(a.out, "a") as f:
            f.write("\n".join(lines) + "\n")
    if world > 1:
        dist.barrier()
    if rank == 0 and not a.keep_files:
        shutil.rmtree(root, ignore_errors=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
