"""BASELINE config 5 at reduced size: synthetic sources sharded over the GPUs of one box (torchrun, one rank per GPU): generate (SD v1.5
ControlNet-canny text2img) -> per-rank filter (WSDAN_CAL-R50 top-10 + CLIP semantic check; --clip ViT-L/14 | RN50) -> NCCL all-gather of
the filter records -> rank 0 writes the aug JSON.  Prints per-rank timings and the JSON statistics.
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 tools/run_config5.py --sources 64 --steps 20"""
import argparse
import json
import os
import sys
import tempfile

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))  # repo root (python tools/<name>.py)
import torch
import torch.distributed as dist

from saspa_aug_b200 import run_aug
from saspa_aug_b200.datasets import SyntheticUtils


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--sources", type=int, default=64)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--clip", default="ViT-L/14")
    ap.add_argument("--root", default=None)
    a = ap.parse_args()
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    root = a.root or os.path.join(tempfile.gettempdir(), "saspa_config5")
    ds = SyntheticUtils(root=root, n_images=a.sources, clip_model=a.clip)
    if rank == 0:
        ds.materialize()
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
    if world > 1:
        dist.init_process_group("nccl")
        dist.barrier()
    cfg = run_aug.AugConfig(NUM_INFERENCE_STEPS=a.steps, SAMPLER="unipcmultistep", MICRO_BATCH=32).apply_dataset_rules()
    prompts = [f"an airplane on a runway at dusk, variation {i}." for i in range(40)]
    json_path, stats = run_aug.run_sharded(cfg, ds, prompts, root)
    print(json.dumps(stats), flush=True)
    if rank == 0:
        d = json.load(open(json_path))
        n = sum(len(v) for v in d.values())
        ips = world * stats["generated"] / stats["generate_s"]
        print(f"CONFIG5 world {world}: {len(d)} sources, {stats['records']} augmentations filtered, {n} kept -> {json_path}; "
              f"rank-0 generate {stats['generate_s']:.1f} s (~{ips:.1f} img/s job-wide incl. PNG writes), filter {stats['filter_s']:.1f} s", flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
