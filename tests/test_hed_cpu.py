"""HED conditioning, CPU side: the state-dict layout the product reads is the restated network's, the host resize mirror follows
controlnet_aux's rule, the tail restatement behaves, and a missing checkpoint is reported (no silent random weights)."""
import numpy as np
import pytest
import torch

from oracle import hed as ohed
from saspa_aug_b200 import checkpoints as ck
from saspa_aug_b200 import hed, run_aug


def test_state_dict_layout_matches_the_restated_network():
    want = {k: tuple(v.shape) for k, v in ohed.ControlNetHED().state_dict().items()}
    assert dict(ck.hed_shapes()) == want
    sd = ck.random_hed_state_dict(3)
    assert {k: tuple(v.shape) for k, v in sd.items()} == want
    assert all(torch.equal(sd[k], ck.random_hed_state_dict(3)[k]) for k in sd)  # deterministic in the seed
    assert sum(v.numel() for v in sd.values()) == 14_716_168  # 13 VGG convolutions + 5 projections + norm


@pytest.mark.parametrize("hw,res", [((100, 150), 128), ((512, 512), 512), ((300, 200), 256), ((700, 1000), 512)])
def test_resize_image_mirror(hw, res):
    img = np.random.default_rng(hw[0]).integers(0, 256, size=(*hw, 3), dtype=np.uint8)
    got, want = hed.resize_image(img, res), ohed.resize_image(img, res)
    assert got.shape == want.shape and np.array_equal(got, want)
    assert min(got.shape[:2]) % 64 == 0 and got.shape[0] % 64 == 0 and got.shape[1] % 64 == 0


def test_tail_restatement_levels_and_safe_step():
    rng = np.random.default_rng(0)
    sides = [rng.normal(size=(64 >> k, 96 >> k)).astype(np.float32) for k in range(5)]
    out = ohed.fuse_sides(sides, 64, 96)
    assert out.dtype == np.uint8 and out.shape == (64, 96) and 60 < out.mean() < 200
    flat = ohed.fuse_sides([np.full_like(s, 0.0) for s in sides], 64, 96)
    assert np.all(flat == 127)  # sigmoid(0) * 255 = 127.5, truncated
    assert set(np.unique(ohed.fuse_sides(sides, 64, 96, safe=True))) <= {0, 127, 255}


def test_missing_checkpoint_and_flavour_table():
    with pytest.raises(FileNotFoundError):
        hed.HEDdetector.from_pretrained("lllyasviel/ControlNet", device="cpu")
    assert run_aug.CONTROLNET_DICT_SD["hed"] == "lllyasviel/sd-controlnet-hed" and "hed" not in run_aug.CONTROLNET_DICT_SD_XL
    cfg = run_aug.AugConfig(CONTROLNET="hed").apply_dataset_rules()
    assert "/controlnet/" in run_aug.output_folder("/data", cfg) and "/hed/" in run_aug.output_folder("/data", cfg)
