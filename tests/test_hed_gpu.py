"""HED conditioning (run_aug/run_aug.py:311-312, :438-439: controlnet_aux.HEDdetector; un-vendored => oracle/hed.py restates it, parity
unpinned): the detector's tail kernel against the numpy / OpenCV tail, the network's side outputs and the final control map against the
fp32 oracle with a tolerance calibrated on stock torch bf16, the drop-in call surface, and the sharded driver with CONTROLNET == "hed"."""
import os

import numpy as np
import pytest
import torch

from oracle import hed as ohed
from saspa_aug_b200 import checkpoints as ck
from saspa_aug_b200 import ops, run_aug
from saspa_aug_b200.hed import HEDdetector
from saspa_aug_b200.synthetic import synthetic_source

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("H,W,n,safe", [(128, 192, 3, False), (64, 64, 1, False), (192, 128, 2, True), (512, 512, 2, False)])
def test_hed_fuse_matches_the_numpy_opencv_tail(cuda_device, H, W, n, safe):
    """saspa_hed_fuse_u8 on random side outputs: cv2.resize(INTER_LINEAR) x 5 -> mean -> sigmoid (fp64) -> * 255 truncated.  The u8
    truncation turns a last-bit difference of the fp32 interpolation (OpenCV's SIMD build may contract to FMA) into a +-1 step: at most
    0.5 % of the pixels may differ, by one level."""
    g = torch.Generator().manual_seed(H * 7 + W)
    sides = [(torch.randn((n, H >> k, W >> k, 1), generator=g) * 1.5 + 0.2 * k).contiguous() for k in range(5)]
    got = ops.hed_fuse([s.cuda() for s in sides], H, W, safe=safe, out_channels=3).cpu().numpy()
    assert got.shape == (n, H, W, 3) and np.array_equal(got[..., 0], got[..., 1]) and np.array_equal(got[..., 0], got[..., 2])
    for i in range(n):
        want = ohed.fuse_sides([s[i, :, :, 0].numpy() for s in sides], H, W, safe=safe)
        d = np.abs(got[i, :, :, 0].astype(np.int32) - want.astype(np.int32))
        if safe:  # three output levels: a flip is a 127-level step, none may happen away from the exact thirds
            assert (d != 0).mean() <= 1e-4
        else:
            assert d.max() <= 1 and (d != 0).mean() <= 5e-3, (d.max(), (d != 0).mean())
    one = ops.hed_fuse([s.cuda() for s in sides], H, W, safe=safe, out_channels=1)
    assert one.shape == (n, H, W, 1) and torch.equal(one[..., 0].cpu(), torch.from_numpy(got[..., 0]))


def _nets(seed):
    sd = ck.random_hed_state_dict(seed)
    net = ohed.ControlNetHED()
    net.load_state_dict(sd)
    return sd, net.eval()


def test_hed_network_and_control_map_against_the_fp32_oracle(cuda_device):
    """Side outputs: max error <= max(2 x the error of stock torch bf16 on the same graph, 2 % of max|ref|).  Control map: mean |diff| in
    u8 levels <= max(1.5 x torch bf16's, 0.75) and 99 % of the pixels within 4 levels."""
    sd, net = _nets(11)
    imgs = np.stack([synthetic_source(40, 128, 192, kind="blobs"), synthetic_source(41, 128, 192, kind="noise")])
    det = HEDdetector.from_state_dict(sd, "cuda")
    got_sides = [s.cpu().numpy()[..., 0] for s in det.netNetwork(torch.from_numpy(imgs).cuda())]
    got_map = det.detect_batch(torch.from_numpy(imgs).cuda()).cpu().numpy()
    x = torch.from_numpy(imgs).float().permute(0, 3, 1, 2)
    with torch.no_grad():
        ref_sides = [s[:, 0].numpy() for s in net(x)]
        nb = ohed.ControlNetHED()
        nb.load_state_dict(sd)
        nb = nb.cuda().to(torch.bfloat16)
        bf_sides = [s[:, 0].float().cpu().numpy() for s in nb(x.cuda().to(torch.bfloat16))]
    for k in range(5):
        e, eb = np.abs(got_sides[k] - ref_sides[k]).max(), np.abs(bf_sides[k] - ref_sides[k]).max()
        print(f"HED side {k + 1} {ref_sides[k].shape}: max err {e:.4g} (torch bf16 {eb:.4g}), max|ref| {np.abs(ref_sides[k]).max():.3g}")
        assert e <= max(2.0 * eb, 2e-2 * np.abs(ref_sides[k]).max())
    for i in range(2):
        want = ohed.hed_detect(net, imgs[i], 128, 128)
        bf = ohed.fuse_sides([s[i] for s in bf_sides], 128, 192)
        assert want.shape == got_map[i].shape == (128, 192, 3)
        d = np.abs(got_map[i].astype(np.int32) - want.astype(np.int32))
        db = np.abs(bf.astype(np.int32) - want[..., 0].astype(np.int32))
        print(f"HED control map {i}: mean |diff| {d.mean():.3f} levels (torch bf16 {db.mean():.3f}), max {d.max()}, spread of the map {want.std():.1f}")
        assert want.std() > 10  # not a saturated map
        assert d.mean() <= max(1.5 * db.mean(), 0.75) and (d <= 4).mean() >= 0.99


def test_hed_detector_call_surface(cuda_device, tmp_path):
    """from_pretrained(directory) reads ControlNetHED.pth; __call__(PIL | ndarray) -> PIL RGB at the resized resolution, equal to the
    batched entry; grey / RGBA inputs go through HWC3; a missing checkpoint fails loudly."""
    from PIL import Image

    sd, net = _nets(12)
    torch.save(sd, tmp_path / "ControlNetHED.pth")
    det = HEDdetector.from_pretrained(str(tmp_path), device="cuda")
    img = synthetic_source(5, 100, 150, kind="blobs")  # -> 128 x 192 at detect_resolution 128
    out = det(Image.fromarray(img), detect_resolution=128, image_resolution=128)
    assert isinstance(out, Image.Image) and out.mode == "RGB" and out.size == (192, 128)
    resized = ohed.resize_image(img, 128)
    batch = det.detect_batch(torch.from_numpy(resized)[None].cuda())[0].cpu().numpy()
    assert np.array_equal(np.array(out), batch)
    want = ohed.hed_detect(net, img, 128, 128)
    assert np.abs(np.array(out).astype(np.int32) - want.astype(np.int32)).mean() <= 1.5
    arr = det(img[..., 0], detect_resolution=128, image_resolution=128, output_type="np")
    assert isinstance(arr, np.ndarray) and arr.shape == (128, 192, 3)
    safe = det(img, detect_resolution=128, image_resolution=128, safe=True, output_type="np")
    assert set(np.unique(safe)) <= {0, 127, 255}
    with pytest.raises(NotImplementedError):
        det(img, detect_resolution=128, image_resolution=128, scribble=True)
    with pytest.raises(FileNotFoundError):
        HEDdetector.from_pretrained(str(tmp_path / "nowhere"), device="cuda")


def test_driver_with_hed_controlnet(cuda_device, tmp_path):
    """CONTROLNET == "hed" through init_pipeline + generate: the output folder carries the flavour (run_aug.py:97 layout), the saved
    "_control.png" of each source is the detector's map, and the images differ from the canny-conditioned ones."""
    from PIL import Image

    from saspa_aug_b200.datasets import SyntheticUtils

    ds = SyntheticUtils(root=str(tmp_path / "ds"), n_images=3, size=(128, 128)).materialize()
    prompts = [f"an airplane over water {i}." for i in range(8)]
    outs = {}
    for flavour in ("hed", "canny"):
        cfg = run_aug.AugConfig(BASE_MODEL="tiny", CONTROLNET=flavour, RESOLUTION=128, NUM_INFERENCE_STEPS=3, MICRO_BATCH=4, USE_ARTISTIC_PROMPTS=True).apply_dataset_rules()
        pipe = run_aug.init_pipeline("tiny", flavour, cfg.SDEDIT, sampler="ddim")
        out_dir = run_aug.output_folder(str(tmp_path / "ds"), cfg)
        assert f"/{flavour}/" in out_dir
        written = run_aug.generate(cfg, ds, pipe, prompts, out_dir)
        assert len(written) == 6 and all(os.path.exists(p) for _, _, p in written)
        outs[flavour] = (pipe, out_dir, written)
    pipe, out_dir, written = outs["hed"]
    for p in ds.original_images_paths:
        src = np.array(Image.open(p).convert("RGB"))
        ctrl = np.array(Image.open(os.path.join(out_dir, os.path.basename(p)[:-4] + "_control.png")))
        assert np.array_equal(ctrl, pipe.hed_detector.detect_batch(torch.from_numpy(src)[None].cuda())[0].cpu().numpy())
    a = np.array(Image.open(written[0][2])).astype(np.int32)
    b = np.array(Image.open(outs["canny"][2][0][2])).astype(np.int32)
    assert a.shape == b.shape and np.abs(a - b).mean() > 0.5
    with pytest.raises(KeyError):
        run_aug.init_pipeline("tiny_xl", "hed", False)
