"""saspa_resize_area_u8 against the installed cv2 (the call the reference's utils.resize_image makes for k <= 1, all_utils/utils.py:58-79):
bit-exact on OpenCV's three INTER_AREA code paths, on the shapes resize_image produces, for 1 / 3 / 4 channels."""
import numpy as np
import pytest
import torch

cv2 = pytest.importorskip("cv2")

from saspa_aug_b200 import ops, run_aug

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("sh,sw,dh,dw,c", [
    (37, 53, 16, 24, 3), (100, 150, 64, 96, 3), (333, 500, 192, 256, 3), (700, 1000, 512, 704, 3), (1365, 2048, 512, 768, 3),   # float area
    (64, 96, 32, 48, 3), (96, 96, 32, 32, 3), (128, 192, 64, 64, 1), (90, 120, 30, 60, 4), (1024, 1536, 512, 768, 3), (512, 1024, 512, 512, 3),  # integer
    (75, 50, 64, 64, 3), (525, 700, 512, 704, 3), (700, 525, 704, 512, 3), (130, 60, 128, 64, 1), (40, 100, 64, 64, 4),        # one axis up
    (64, 64, 64, 64, 3)])
def test_resize_area_matches_cv2(cuda_device, sh, sw, dh, dw, c):
    src = np.random.default_rng(sh * 1000 + sw + c).integers(0, 256, (sh, sw, c), dtype=np.uint8)
    want = cv2.resize(src, (dw, dh), interpolation=cv2.INTER_AREA).reshape(dh, dw, c)
    got = ops.resize_area(torch.from_numpy(src).cuda(), dh, dw).cpu().numpy()
    assert got.shape == want.shape
    assert np.array_equal(got, want), (np.abs(got.astype(int) - want.astype(int)).max(), (got != want).mean())


@pytest.mark.parametrize("hw", [(695, 1024), (768, 1024), (525, 700), (1200, 1600), (512, 683), (600, 600), (512, 512), (3000, 2000)])
def test_resize_image_device_equals_the_host_mirror(cuda_device, hw):
    src = np.random.default_rng(hw[0]).integers(0, 256, (*hw, 3), dtype=np.uint8)
    want = run_aug.resize_image(src, 512)  # cv2 on the host (equal to the reference's function: tests/test_host_logic_cpu.py)
    got = run_aug.resize_image_device(torch.from_numpy(src).cuda(), 512).cpu().numpy()
    assert np.array_equal(got, want)


@pytest.mark.parametrize("sh,sw,dh,dw,c", [(30, 40, 64, 64, 3), (375, 500, 512, 704, 3), (300, 400, 512, 704, 3), (100, 100, 128, 192, 1), (64, 64, 128, 128, 4),
                                           (333, 250, 704, 512, 3), (97, 131, 256, 320, 3), (17, 23, 512, 704, 3)])
def test_resize_lanczos4_matches_cv2(cuda_device, sh, sw, dh, dw, c):
    src = np.random.default_rng(sh * 1000 + sw + c).integers(0, 256, (sh, sw, c), dtype=np.uint8)
    want = cv2.resize(src, (dw, dh), interpolation=cv2.INTER_LANCZOS4).reshape(dh, dw, c)
    got = ops.resize_lanczos4(torch.from_numpy(src).cuda(), dh, dw).cpu().numpy()
    assert np.array_equal(got, want), (np.abs(got.astype(int) - want.astype(int)).max(), (got != want).mean())


@pytest.mark.parametrize("hw", [(300, 400), (375, 500), (200, 333), (500, 500), (120, 90)])
def test_resize_image_device_upscales_like_the_host_mirror(cuda_device, hw):
    src = np.random.default_rng(hw[0]).integers(0, 256, (*hw, 3), dtype=np.uint8)
    want = run_aug.resize_image(src, 512)
    got = run_aug.resize_image_device(torch.from_numpy(src).cuda(), 512).cpu().numpy()
    assert np.array_equal(got, want)


def test_generate_with_device_resize_writes_the_same_files(cuda_device, tmp_path):
    """AugConfig.DEVICE_RESIZE: the loader threads resize on the GPU; sources larger than the resolution (non-square, so the x64 rounding
    makes one axis a slight up-scale for some of them) and one smaller source (LANCZOS4) give byte-identical "_source.png" and augmentation pixels."""
    import os

    from PIL import Image

    from saspa_aug_b200.datasets import SyntheticUtils

    ds = SyntheticUtils(root=str(tmp_path / "ds"), n_images=4, sizes=[(200, 300), (260, 256), (333, 250), (100, 90)]).materialize()
    prompts = [f"an airplane over a field {i}." for i in range(6)]
    outs = {}
    for flag in (False, True):
        cfg = run_aug.AugConfig(BASE_MODEL="tiny", RESOLUTION=128, NUM_INFERENCE_STEPS=2, MICRO_BATCH=4, USE_ARTISTIC_PROMPTS=True, DEVICE_RESIZE=flag).apply_dataset_rules()
        pipe = run_aug.init_pipeline("tiny", "canny", cfg.SDEDIT, sampler="ddim")
        out_dir = str(tmp_path / f"out_{int(flag)}")
        written = run_aug.generate(cfg, ds, pipe, prompts, out_dir)
        assert len(written) == 8
        outs[flag] = {os.path.basename(p): np.array(Image.open(p)) for p in sorted(os.path.join(out_dir, n) for n in os.listdir(out_dir))}
    assert outs[False].keys() == outs[True].keys() and any("_source" in k for k in outs[True])
    for k in outs[False]:
        assert np.array_equal(outs[False][k], outs[True][k]), k
