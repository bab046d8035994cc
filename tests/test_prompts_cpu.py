"""Prompt assembly + output file naming against a transcript of the reference's OWN lines (run_aug.py:380-429 exec'd
unchanged, tests/golden/make_prompt_golden.py): np.random / random draw order, artistic + camera suffixes, the short-circuited
random.random(), compcars-parts prefix, sub-class insertion per dataset, '/' -> '-' and the 40-character stem cut in the file name."""
import importlib.util
import json
import os
from pathlib import Path

import pytest

from oracle import ref_import
from saspa_aug_b200 import run_aug
from saspa_aug_b200.prompts import ARTISTIC_PROMPTS, IMAGE_VARIATIONS_PROMPTS

G = os.path.join(os.path.dirname(__file__), "golden")
spec = importlib.util.spec_from_file_location("make_prompt_golden", os.path.join(G, "make_prompt_golden.py"))
gen = importlib.util.module_from_spec(spec)
spec.loader.exec_module(gen)


class _Ds(gen.DsStub):
    def __init__(self, dataset):
        self.paths, _, self.classes = gen.fixture(dataset)
        self.meta_class = gen.META[dataset]

    def get_image_stem_to_class_str_dict(self):
        return {Path(p).stem: c for p, c in zip(self.paths, self.classes)}

    def get_image_path_to_class_str_dict(self):
        return dict(zip(self.paths, self.classes))


def product_transcript(case):
    cfg = run_aug.AugConfig(BASE_MODEL="sd_v1.5", **case)
    paths, prompts, _ = gen.fixture(case["DATASET"])
    draws = run_aug.replay_prompt_draws([p.strip()[:run_aug.MAX_PROMPT_LENGTH] for p in prompts], paths, cfg, _Ds(case["DATASET"]))
    return [[index, i, d.prompt, run_aug.aug_file_name(Path(paths[index]).stem, d.prompt, i)] for index, row in enumerate(draws) for i, d in enumerate(row)]


@pytest.mark.parametrize("name", sorted(gen.CASES))
def test_prompt_replay_matches_the_reference_transcript(name):
    gold = json.load(open(os.path.join(G, "prompt_transcript.json")))["cases"][name]
    assert gold["config"] == gen.CASES[name]
    got = product_transcript(gen.CASES[name])
    assert got == gold["transcript"]
    assert any(len(t[3]) > 60 for t in got)


@pytest.mark.skipif(not ref_import.available(), reason="/root/reference not present (GPU box)")
@pytest.mark.parametrize("name", sorted(gen.CASES))
def test_prompt_replay_matches_the_reference_lines_live(name):
    assert product_transcript(gen.CASES[name]) == gen.reference_transcript(gen.CASES[name])


@pytest.mark.skipif(not ref_import.available(), reason="/root/reference not present (GPU box)")
def test_suffix_vocabularies_are_the_reference_constants():
    import sys

    sys.path.insert(0, ref_import.REFERENCE_ROOT)
    import prompts_engineering as pe

    assert ARTISTIC_PROMPTS == pe.ARTISTIC_PROMPTS and IMAGE_VARIATIONS_PROMPTS == pe.IMAGE_VARIATIONS_PROMPTS


def test_blip_subject_draw_shares_the_random_stream_and_skips_existing_outputs():
    """run_aug.py:430-432,:446: the same-class subject image is drawn from the global `random` stream AFTER the exists-check, so skipped
    items draw nothing and shift every later draw."""
    import random

    case = dict(gen.CASES["cub_plain"], DATASET="cub")
    ds = _Ds("cub")
    ds.get_image_path_with_same_class = lambda p: [q for q, c in zip(ds.paths, ds.classes) if c == dict(zip(ds.paths, ds.classes))[p]]
    cfg = run_aug.AugConfig(BASE_MODEL="blip_diffusion", **case)
    _, prompts, _ = gen.fixture("cub")
    skipped = {(1, 0), (2, 1)}
    draws = run_aug.replay_prompt_draws([p.strip() for p in prompts], ds.paths, cfg, ds, skip=lambda index, i, prompt: (index, i) in skipped)
    rs = random.Random(case["SEED"])
    for index, row in enumerate(draws):
        for i, d in enumerate(row):
            if (index, i) in skipped:
                assert d.skipped and d.subject_path is None
            else:
                assert d.subject_path == rs.choice(ds.get_image_path_with_same_class(ds.paths[index]))
    assert "_style_img_from_diff_img" in run_aug.output_folder("/d", cfg)
