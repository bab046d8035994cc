"""The optional LPIPS filter (all_utils/utils.py:269-270, :377-381, :576-590) on the B200 kernels against the fp32
restatement of the `lpips` package's algorithm (oracle/lpips_alex.py, parity unpinned: the package is absent offline).  Integer part
(PIL "L" conversion + bicubic resize) bit-exact; distances within the bf16-trunk tolerance; keep/drop decisions identical for bounds placed
in gaps of the oracle's distances."""
import json
import os

import numpy as np
import pytest
import torch

from oracle import lpips_alex as L
from saspa_aug_b200 import checkpoints as ck
from saspa_aug_b200 import filtering, ops, run_aug
from saspa_aug_b200.datasets import SyntheticUtils
from saspa_aug_b200.filter_nets import LPIPSAlex
from saspa_aug_b200.synthetic import synthetic_source

pytestmark = pytest.mark.gpu


def test_luma_and_resize_bit_exact_vs_pil(cuda_device):
    from PIL import Image

    imgs = np.stack([synthetic_source(s, 160, 224, kind=("blobs", "noise", "smooth")[s % 3]) for s in range(3)])
    got = ops.rgb_to_luma3(torch.from_numpy(imgs).cuda()).cpu().numpy()
    for a, g in zip(imgs, got):
        assert np.array_equal(g, np.asarray(Image.fromarray(a).convert("L").convert("RGB")))
    net = LPIPSAlex(ck.random_lpips_state_dict(31))
    pre = net.preprocess(torch.from_numpy(imgs).cuda()).cpu().numpy()
    for a, g in zip(imgs, pre):
        want = np.asarray(Image.fromarray(a).convert("L").convert("RGB").resize((256, 256)))
        assert np.array_equal(g, want)


def test_lpips_distance_matches_oracle(cuda_device):
    sd = ck.random_lpips_state_dict(31)
    assert [k for k, _ in ck.lpips_alex_shapes()] == [k for k, _ in L.state_dict_keys()]
    oracle = L.load(L.LPIPSAlex(), sd)
    net = LPIPSAlex(sd)
    a = np.stack([synthetic_source(10 + s, 192, 192, kind=("blobs", "smooth")[s % 2]) for s in range(6)])
    b = np.stack([synthetic_source(40 + s, 128, 160, kind=("blobs", "noise", "smooth")[s % 3]) for s in range(6)])
    b[0] = 0  # a black image against a textured one
    want = oracle(torch.stack([L.preprocess(x) for x in a]), torch.stack([L.preprocess(x) for x in b]))
    pa, pb = net.preprocess(torch.from_numpy(a).cuda()), net.preprocess(torch.from_numpy(b).cuda())
    got = net(pa, pb).cpu()
    print("lpips oracle", [round(float(v), 5) for v in want], "ours", [round(float(v), 5) for v in got])
    assert torch.allclose(got, want, rtol=3e-2, atol=2e-4)
    assert float(net(pa, pa).abs().max()) == 0.0  # identical inputs -> exactly zero


def test_lpips_filter_in_the_json_writer(cuda_device, tmp_path):
    """lpips_min <= d(source, augmentation) <= lpips_max applied after the confidence filter and before the CLIP filters, in the
    reference's order; expected JSON from the oracle's distances with the bounds midway in its widest gaps."""
    from PIL import Image

    ds = SyntheticUtils(root=str(tmp_path / "ds"), n_images=5, sizes=[(160, 160), (128, 192)]).materialize()
    out_dir = run_aug.output_folder(str(tmp_path / "ds"), run_aug.AugConfig())
    os.makedirs(out_dir)
    for index, p in enumerate(ds.original_images_paths):
        stem = os.path.splitext(os.path.basename(p))[0]
        src = np.asarray(Image.open(p))
        for i in range(3):
            if i == 0:  # a light perturbation of the source, a blend, an unrelated image: three distance regimes
                aug = np.clip(src.astype(int) + np.random.default_rng(index).integers(-6, 7, src.shape), 0, 255).astype(np.uint8)
            elif i == 1:
                other = synthetic_source(300 + index, *src.shape[:2])
                aug = ((src.astype(int) + other) // 2).astype(np.uint8)
            else:
                aug = synthetic_source(600 + index, 144, 144, kind="noise")
            Image.fromarray(aug).save(os.path.join(out_dir, run_aug.aug_file_name(stem, f"an airplane, take {i}", i)))
    oracle = L.load(L.LPIPSAlex(), ck.random_lpips_state_dict(ds.lpips_seed))
    names = os.listdir(out_dir)
    matched = filtering.match_augmentations(ds.original_images_paths, names, out_dir)
    order = [(p, a) for p in ds.original_images_paths for a in matched[os.path.basename(p)]]
    d = oracle(torch.stack([L.preprocess(np.asarray(Image.open(p).convert("RGB"))) for p, _ in order]),
               torch.stack([L.preprocess(np.asarray(Image.open(a).convert("RGB"))) for _, a in order]))
    srt = torch.sort(d).values
    gaps = srt[1:] - srt[:-1]
    lo_k = int(torch.argmax(gaps[: len(srt) // 2]))
    hi_k = int(torch.argmax(gaps[len(srt) // 2:])) + len(srt) // 2
    lo, hi = float((srt[lo_k] + srt[lo_k + 1]) / 2), float((srt[hi_k] + srt[hi_k + 1]) / 2)
    want = {os.path.basename(p): [] for p in ds.original_images_paths}
    for (p, a), v in zip(order, d):
        if lo <= float(v) <= hi:
            want[os.path.basename(p)].append(a)
    jp, det = filtering.create_json_of_image_name_to_augmented_images_paths("synthetic", out_dir, lpips_min=lo, lpips_max=hi, init_log=False, ds_utils=ds,
                                                                            return_details=True)
    assert os.path.basename(jp) == f"lpips_min_{lo}-lpips_max_{hi}-aug.json"
    print("lpips distances: oracle", [round(float(v), 4) for v in d], "ours", [round(float(v), 4) for v in det["lpips"]], "bounds", lo, hi)
    assert json.load(open(jp)) == want
    assert 0 < sum(len(v) for v in want.values()) < len(order)
    with pytest.raises(TypeError):
        filtering.create_json_of_image_name_to_augmented_images_paths("synthetic", out_dir, lpips_min=0.1, init_log=False, ds_utils=ds)
