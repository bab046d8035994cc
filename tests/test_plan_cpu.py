"""Host-side planning of the sharded generation loop: mixed source aspect ratios, partially resumed folders, corrupt leftovers,
failure propagation through the one collective (run_aug.py:313, :406)."""
import os
import subprocess
import sys
from pathlib import Path

import numpy as np
import pytest
from PIL import Image

from saspa_aug_b200 import run_aug
from saspa_aug_b200.datasets import SyntheticUtils

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_micro_batches_are_uniform_for_mixed_aspect_ratios_and_partial_resume(tmp_path):
    """Real FGVC sources are not square: resize_image keeps the aspect ratio (300x400 -> 512x704, 400x300 -> 704x512).  Work is
    bucketed by the resized shape, so every micro-batch holds one latent shape whatever MICRO_BATCH is and whatever already exists."""
    sizes = [(300, 400), (400, 300), (512, 512), (300, 400), (600, 600)]
    ds = SyntheticUtils(root=str(tmp_path / "ds"), n_images=11, sizes=sizes).materialize()
    cfg = run_aug.AugConfig(MICRO_BATCH=4, NUM_PER_IMAGE=2)
    out_dir = run_aug.output_folder(str(tmp_path / "ds"), cfg)
    os.makedirs(out_dir)
    paths = ds.original_images_paths
    prompts = [f"an airplane, view {k}." for k in range(9)]
    names = run_aug.sample_prompts(prompts, paths, cfg, ds)
    # a killed run left: both outputs of source 0, the first of source 3, the second of source 6
    done = {(0, 0), (0, 1), (3, 0), (6, 1)}
    for index, i in done:
        Image.new("RGB", (8, 8)).save(os.path.join(out_dir, run_aug.aug_file_name(Path(paths[index]).stem, names[index][i], i)))

    def exists(index, i, prompt):
        return os.path.exists(os.path.join(out_dir, run_aug.aug_file_name(Path(paths[index]).stem, prompt, i)))

    draws = run_aug.replay_prompt_draws([p.strip() for p in prompts], paths, cfg, ds, skip=exists)
    hw = {k: run_aug.source_hw(paths[k], cfg.RESOLUTION) for k in range(len(paths))}
    assert hw[0] == (512, 704) and hw[1] == (704, 512) and hw[2] == (512, 512) and hw[4] == (512, 512)
    for world in (1, 2, 3):
        seen = set()
        for rank in range(world):
            mine = run_aug.shard_indices(len(paths), rank, world)
            existing, chunks = run_aug.plan_work(cfg, paths, draws, out_dir, mine, lambda k: hw[k])
            assert {(a, b) for a, b, _ in existing} == {d for d in done if d[0] in mine}
            for chunk in chunks:
                assert 1 <= len(chunk) <= cfg.MICRO_BATCH and len({hw[w.index] for w in chunk}) == 1  # uniform by construction
                assert all(w.index in mine and w.out_path.endswith(f"_{w.i}.png") for w in chunk)
                seen |= {(w.index, w.i) for w in chunk}
        assert seen == {(k, i) for k in range(len(paths)) for i in range(2)} - done


def test_verify_written_drops_and_deletes_truncated_files(tmp_path):
    good, bad, gone = tmp_path / "a.png", tmp_path / "b.png", tmp_path / "c.png"
    Image.fromarray(np.random.default_rng(0).integers(0, 255, (64, 64, 3), dtype=np.uint8)).save(good)
    Image.fromarray(np.random.default_rng(1).integers(0, 255, (64, 64, 3), dtype=np.uint8)).save(bad)
    bad.write_bytes(bad.read_bytes()[:200])
    out = run_aug.verify_written([(0, 0, str(good)), (0, 1, str(bad)), (1, 0, str(gone))])
    assert out == [(0, 0, str(good))] and not bad.exists() and good.exists()


_FAIL_WORKER = r'''
import os, sys
import numpy as np
sys.path.insert(0, sys.argv[1])
from saspa_aug_b200 import run_aug
from saspa_aug_b200.datasets import SyntheticUtils
import torch.distributed as dist
rank = int(os.environ["RANK"])
ds = SyntheticUtils(root=sys.argv[2], n_images=4, size=(32, 32))
if rank == 0:
    ds.materialize()
dist.init_process_group("gloo")
dist.barrier()

def gen(cfg, ds_utils, pipe, prompts, out_dir, rank=0, world=1):
    if rank == 1:
        raise RuntimeError("CUDA out of memory (simulated) on rank 1")
    return []

try:
    run_aug.run_sharded(run_aug.AugConfig(), ds, ["an airplane."], sys.argv[2], generate_fn=gen, filter_fn=lambda *a, **k: np.zeros((0, 4), np.int32), device="cpu")
    print(f"RANK{rank}_NO_ERROR")
except RuntimeError as e:
    print(f"RANK{rank}_RAISED:{e}")
dist.destroy_process_group()
'''


def test_a_failing_rank_does_not_strand_the_others_in_the_collective(tmp_path):
    """A stage that raises on ONE rank: every rank still reaches the all-gather, learns about the failure there and raises; nobody hangs,
    no JSON is written."""
    w = tmp_path / "worker.py"
    w.write_text(_FAIL_WORKER)
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
                        "--master-port", "29617", str(w), ROOT, str(tmp_path / "ds")], capture_output=True, text=True, timeout=240)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "RANK0_RAISED:the generate / filter stage failed on rank(s) [1]" in r.stdout and "RANK1_RAISED:" in r.stdout and "NO_ERROR" not in r.stdout
    assert not list((tmp_path / "ds").rglob("*aug.json"))
