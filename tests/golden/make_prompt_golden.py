"""Golden transcript of the reference's prompt assembly + file naming: lines 380-429 of /root/reference/run_aug/run_aug.py are read from
the reference tree, dedented and exec'd UNCHANGED inside a two-line harness that supplies what ``__main__`` defines (the module constants,
``utils.set_seed(SEED)`` seeding, a dataset-utils stand-in) and records (index, i, prompt, output file name) after line 429.  The product's
``run_aug.replay_prompt_draws`` must reproduce every transcript (tests/test_prompts_cpu.py); the same test re-runs this harness live
when /root/reference is present.

Run in the build container:  python tests/golden/make_prompt_golden.py
"""
import json
import os
import random
import sys
import textwrap
from pathlib import Path

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
REF = "/root/reference/run_aug/run_aug.py"
FIRST, LAST = 380, 429  # prompts clean-up ... output_path (inclusive, 1-based)

CASES = {
    # name: constants of run_aug.py:513-556 that matter for the draw order
    "planes_default": dict(DATASET="planes", SEED=1, NUM_PER_IMAGE=2, USE_ARTISTIC_PROMPTS=True, ARTISTIC_PROMPTS_PROB=0.5, USE_CAMERA_VARIATIONS_PROMPTS=False,
                           CAMERA_VAIRATIONS_PROB=0.5, PROMPT_WITH_SUB_CLASS=True),
    "planes_art_p03_camera": dict(DATASET="planes", SEED=3, NUM_PER_IMAGE=3, USE_ARTISTIC_PROMPTS=True, ARTISTIC_PROMPTS_PROB=0.3,
                                  USE_CAMERA_VARIATIONS_PROMPTS=True, CAMERA_VAIRATIONS_PROB=0.5, PROMPT_WITH_SUB_CLASS=True),
    "cars_camera_only": dict(DATASET="cars", SEED=5, NUM_PER_IMAGE=2, USE_ARTISTIC_PROMPTS=False, ARTISTIC_PROMPTS_PROB=0.5, USE_CAMERA_VARIATIONS_PROMPTS=True,
                             CAMERA_VAIRATIONS_PROB=0.7, PROMPT_WITH_SUB_CLASS=True),
    "cub_plain": dict(DATASET="cub", SEED=1, NUM_PER_IMAGE=2, USE_ARTISTIC_PROMPTS=False, ARTISTIC_PROMPTS_PROB=0.5, USE_CAMERA_VARIATIONS_PROMPTS=False,
                      CAMERA_VAIRATIONS_PROB=0.5, PROMPT_WITH_SUB_CLASS=True),
    "dtd_art_odd": dict(DATASET="dtd", SEED=11, NUM_PER_IMAGE=3, USE_ARTISTIC_PROMPTS=True, ARTISTIC_PROMPTS_PROB=0.5, USE_CAMERA_VARIATIONS_PROMPTS=True,
                        CAMERA_VAIRATIONS_PROB=0.4, PROMPT_WITH_SUB_CLASS=True),
    "compcars_parts": dict(DATASET="compcars-parts", SEED=2, NUM_PER_IMAGE=2, USE_ARTISTIC_PROMPTS=True, ARTISTIC_PROMPTS_PROB=0.5,
                           USE_CAMERA_VARIATIONS_PROMPTS=False, CAMERA_VAIRATIONS_PROB=0.5, PROMPT_WITH_SUB_CLASS=True),
    "planes_no_subclass": dict(DATASET="planes", SEED=1, NUM_PER_IMAGE=2, USE_ARTISTIC_PROMPTS=True, ARTISTIC_PROMPTS_PROB=0.5, USE_CAMERA_VARIATIONS_PROMPTS=False,
                               CAMERA_VAIRATIONS_PROB=0.5, PROMPT_WITH_SUB_CLASS=False),
}
META = {"planes": "airplane", "cars": "car", "cub": "bird", "dtd": "texture", "compcars-parts": "car"}
N_SOURCES = 9


def fixture(dataset):
    """(paths, prompts, class string per source) -- synthetic, deterministic."""
    meta = META[dataset]
    parts = ["headlight", "taillight", "fog light"]
    paths = [f"/data/{dataset}/{parts[k % 3] if dataset == 'compcars-parts' else 'images'}/img_{k:04d}{'_with_a_very_long_stem_' * 2 if k == 4 else ''}.jpg" for k in range(N_SOURCES)]
    prompts = [f"A {meta} seen from angle {k}/7 in soft light." for k in range(12)] + [f"Two {meta}s, one {meta} in front"]
    classes = [f"Maker{k % 4} Model-{k % 3}" for k in range(N_SOURCES)]
    return paths, prompts, classes


class DsStub:
    def get_basic_prompt(self, part=None):
        return f"a photo of the {part} of a car"


def reference_transcript(case: dict):
    import types

    for name in ("matplotlib", "matplotlib.pyplot"):
        sys.modules.setdefault(name, types.ModuleType(name))
    sys.path.insert(0, "/root/reference")
    from prompts_engineering import ARTISTIC_PROMPTS, IMAGE_VARIATIONS_PROMPTS

    lines = open(REF).read().split("\n")[FIRST - 1: LAST]
    body = textwrap.dedent("\n".join(lines))
    # harness: the reference's own for-loop header (run_aug.py:357-358) around its own lines, plus ONE appended line that records the result
    harness = ("for index, source_image_path in enumerate(original_images_paths):\n"
               "    image_stem = Path(source_image_path).stem\n"
               + textwrap.indent(body, "    ") + "\n"
               "        transcript.append((index, i, prompt, output_path.name))\n")
    paths, prompts, classes = fixture(case["DATASET"])
    stem_keyed = case["DATASET"] in ["planes", "cars", "planes_biased"]  # run_aug.py:698
    g = dict(case)
    g.update(np=np, random=random, Path=Path, ARTISTIC_PROMPTS=ARTISTIC_PROMPTS, IMAGE_VARIATIONS_PROMPTS=IMAGE_VARIATIONS_PROMPTS, MAX_FILENAME_LENGTH=40,
             output_folder="/out/images", ds_utils=DsStub(), original_images_paths=paths, transcript=[],
             prompts=[p.strip()[:150] for p in prompts],  # run_aug.py:345
             image_classes_dict={(Path(p).stem if stem_keyed else p): c for p, c in zip(paths, classes)})
    random.seed(case["SEED"])  # utils.set_seed (all_utils/utils.py:32-36)
    np.random.seed(case["SEED"])
    exec(compile(harness, "<run_aug.py:380-429>", "exec"), g)
    return [list(t) for t in g["transcript"]]


def main():
    out = {"reference_lines": [FIRST, LAST], "n_sources": N_SOURCES, "cases": {}}
    for name, case in CASES.items():
        out["cases"][name] = {"config": case, "transcript": reference_transcript(case)}
        print(name, out["cases"][name]["transcript"][0], "...", len(out["cases"][name]["transcript"]))
    json.dump(out, open(os.path.join(ROOT, "tests/golden/prompt_transcript.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
