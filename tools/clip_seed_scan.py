import sys, os, time
sys.path.insert(0, os.getcwd())
import numpy as np, torch
from saspa_aug_b200.datasets import SyntheticUtils
from saspa_aug_b200.filter_nets import AugmentationFilter
from saspa_aug_b200.filtering import SEMANTIC_NEGATIVE_PROMPTS
from saspa_aug_b200.synthetic import synthetic_source
dev = "cuda:0"
rng = np.random.default_rng(0)
imgs = np.stack([synthetic_source(9000 + k, kind=("blobs", "noise")[k % 2]) for k in range(8)] + [rng.integers(0, 256, (512, 512, 3), dtype=np.uint8) for _ in range(4)]
                + [np.clip(rng.normal(128, 60, (512, 512, 3)), 0, 255).astype(np.uint8) for _ in range(4)])
imgs = torch.from_numpy(imgs).to(dev)
for clip_model, seeds in (("RN50", [813, 909, 929, 797]), ("ViT-L/14", list(range(777, 777 + 30)))):
    ds = SyntheticUtils(n_images=1, clip_model=clip_model)
    for s in seeds:
        t0 = time.time()
        ds.clip_seed = s
        _, clip, tok = ds.load_filter_models(ds, dev)
        flt = AugmentationFilter(None, clip, tok([ds.get_basic_prompt()] + SEMANTIC_NEGATIVE_PROMPTS))
        out = flt(imgs, torch.zeros(imgs.shape[0], dtype=torch.int32, device=dev))
        keep = out["semantic"].float()
        print(f"{clip_model} seed {s}: keep probes {keep[:8].mean():.2f} uniform-noise {keep[8:12].mean():.2f} gaussian-noise {keep[12:].mean():.2f}  ({time.time()-t0:.1f}s)", flush=True)
        del flt, clip
        torch.cuda.empty_cache()
        if clip_model != "RN50" and keep.mean() >= 0.75:
            break
