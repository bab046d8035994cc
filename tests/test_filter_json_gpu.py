"""The aug JSON against the REFERENCE's own writer: tests/golden/filter_json_golden.json holds the files
all_utils.utils.create_json_of_image_name_to_augmented_images_paths (utils.py:221-465, run unmodified by
tests/golden/make_filter_json_golden.py with fp32 nets on the CPU) wrote for the folder of tests/filter_fixture.py.  The B200 path
(bf16 trunks, batched, device-side pre-processing) must write the same bytes.

Hot-path configuration (semantic + model-confidence filtering, run_aug.py:721-733): byte-for-byte, no exemption.
Optional filters (disabled in run_aug.py): their thresholds sit inside clusters of near-identical random-net outputs, closer together
than the bf16 forward's own error; a decision may differ ONLY for a pair whose reference value lies within the stated tolerance of the
threshold (prob 4e-3, logit 4e-2), and the test prints every such pair."""
import json
import os
import random

import numpy as np
import pytest

from saspa_aug_b200 import filtering
from saspa_aug_b200.datasets import SyntheticUtils
from tests import filter_fixture as fx

pytestmark = pytest.mark.gpu
GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "filter_json_golden.json")))
PROB_TOL, LOGIT_TOL = 4e-3, 4e-2


@pytest.fixture()
def folder(tmp_path, monkeypatch):
    root, out_dir = fx.build(str(tmp_path / "fx"))
    real = os.listdir

    def listdir(p="."):
        # value order of the JSON = os.listdir order (utils.py:343), which is a property of the file system: replay the order the
        # reference saw when the golden file was made; files it never saw (the truncated one, before its deletion) come last
        if os.path.abspath(p) != os.path.abspath(out_dir):
            return real(p)
        have = set(real(p))
        return [f for f in GOLD["listdir_order"] if f in have] + sorted(have - set(GOLD["listdir_order"]))

    monkeypatch.setattr(os, "listdir", listdir)
    ds = SyntheticUtils(root=root, names=fx.source_names(), labels=fx.labels(), n_classes=fx.NUM_CLASSES, wsdan_seed=GOLD["wsdan_seed"], clip_seed=GOLD["clip_seed"])
    ds.clip_filtering_suffix = ", a type of a bird"  # utils.py:294-296 (the golden case ran under the dataset name "cub")
    ds.alia_threshold = GOLD["alia_threshold"]
    assert ds.get_basic_prompt() == "a photo of an airplane"
    return root, out_dir, ds


def _expected(case, root):
    body = json.loads(json.dumps(GOLD["cases"][case]["body"]).replace("{ROOT}", root.rstrip("/")))
    return body, json.dumps(body)


def test_hot_path_json_is_byte_identical_to_the_reference_writer(cuda_device, folder):
    root, out_dir, ds = folder
    case = GOLD["cases"]["sem+conf"]
    assert case["raw_equals_json_dumps"]
    jp, det = filtering.create_json_of_image_name_to_augmented_images_paths("synthetic", out_dir, init_log=False, ds_utils=ds, return_details=True, **case["kwargs"])
    assert os.path.basename(jp) == case["json_name"]
    want, raw = _expected("sem+conf", root)
    # the writer visited the same (source, file) pairs in the same order -- including the substring cross-matches of "syn_1" and without
    # the truncated file, which was deleted first
    assert [[fx.source_names().index(n), os.path.basename(p)] for n, p, _ in det["pairs"]] == GOLD["pairs"]
    assert not any("truncated" in f for f in os.listdir(out_dir))
    lo = np.array(GOLD["wsdan_logits"])
    sl = np.array(GOLD["semantic_logits"])
    print(f"min reference margins: semantic {np.abs(sl[:, 0] - sl[:, 1:].max(1)).min():.4f}")
    got = open(jp).read()
    assert json.loads(got) == want
    assert got == raw  # byte for byte
    kept = sum(len(v) for v in want.values())
    assert 0 < kept < len(GOLD["pairs"]) and len(want["syn_1.png"]) == 3 and want["syn_11.png"] == []


def _check_optional(case, root, jp, det, value_key, thr, tol, extra_keep=None):
    want, raw = _expected(case, root)
    got = json.load(open(jp))
    assert list(got) == list(want)
    ref_vals = np.array(GOLD[value_key])
    diffs = []
    for name in want:
        for p in set(want[name]) ^ set(got[name]):
            k = [i for i, (n, q, _) in enumerate(det["pairs"]) if n == name and q == p][0]
            diffs.append((name, os.path.basename(p), float(ref_vals[k]), abs(float(ref_vals[k]) - thr)))
    for d in diffs:
        print(f"{case}: decision differs for {d[:2]}: reference value {d[2]:.5f}, {d[3]:.5f} from the threshold {thr:.5f}")
    assert all(d[3] < tol for d in diffs), diffs
    if not diffs:
        assert open(jp).read() == raw
    return diffs


def test_optional_filters_against_the_reference_writer(cuda_device, folder):
    root, out_dir, ds = folder
    common = dict(init_log=False, ds_utils=ds, return_details=True)
    # (a) top-k + too-high confidence (utils.py:357-376)
    c = GOLD["cases"]["conf+too_high"]
    jp, det = filtering.create_json_of_image_name_to_augmented_images_paths("synthetic", out_dir, **c["kwargs"], **common)
    assert os.path.basename(jp) == c["json_name"]
    err = np.abs(det["label_conf"] - np.array(GOLD["label_conf"])).max()
    print(f"label confidence: max |B200 - reference| {err:.5f} (threshold gap {GOLD['label_conf_gap']:.5f})")
    assert err < PROB_TOL
    _check_optional("conf+too_high", root, jp, det, "label_conf", c["kwargs"]["filter_confidence_higher_than"], PROB_TOL)
    # (b) per-class CLIP confidence + semantic (utils.py:186-191, :272-303, :383-404)
    c = GOLD["cases"]["clip_per_class+sem"]
    jp, det = filtering.create_json_of_image_name_to_augmented_images_paths("synthetic", out_dir, **c["kwargs"], **common)
    assert os.path.basename(jp) == c["json_name"]
    err = np.abs(det["class_conf"] - np.array(GOLD["class_conf"])).max()
    print(f"class confidence: max |B200 - reference| {err:.5f} (threshold gap {GOLD['class_conf_gap']:.5f})")
    assert err < PROB_TOL
    _check_optional("clip_per_class+sem", root, jp, det, "class_conf", 1 / fx.NUM_CLASSES / c["kwargs"]["clip_filtering_discount"], PROB_TOL)
    # (c) ALIA confidence filter + semantic (utils.py:411-434): the same `random.random()` draws in the same order
    c = GOLD["cases"]["alia+sem"]
    random.seed(GOLD["alia_random_seed"])
    jp, det = filtering.create_json_of_image_name_to_augmented_images_paths("synthetic", out_dir, **c["kwargs"], **common)
    assert os.path.basename(jp) == c["json_name"]
    err = np.abs(det["max_logit"] - np.array(GOLD["max_logit"])).max()
    print(f"max logit: max |B200 - reference| {err:.5f} (threshold gap {GOLD['alia_gap']:.5f})")
    assert err < LOGIT_TOL
    thr = GOLD["alia_threshold"]
    fired_ref, fired = np.array(GOLD["max_logit"]) > thr, det["max_logit"] > thr
    near = np.abs(np.array(GOLD["max_logit"]) - thr) < LOGIT_TOL
    assert not (fired_ref != fired)[~near].any()
    if (fired_ref == fired).all():
        _check_optional("alia+sem", root, jp, det, "max_logit", thr, 0.0)  # same confidence tests => same draws => same bytes
    else:  # a flipped near-threshold test shifts every later random.random() draw: only the pairs before the first flip are comparable
        first = int(np.argmax(fired_ref != fired))
        print(f"alia+sem: confidence test flipped at pair {first} ({GOLD['max_logit'][first]:.5f} vs threshold {thr}); later draws are shifted")
