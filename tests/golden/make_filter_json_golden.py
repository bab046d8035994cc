"""Golden aug-JSON bodies produced by the REFERENCE's own writer, all_utils.utils.create_json_of_image_name_to_augmented_images_paths
(/root/reference/all_utils/utils.py:221-465), run unmodified on the deterministic folder of tests/filter_fixture.py.

Shims -- none of them touches the writer's logic:
  * dataset_utils.DS_UTILS_DICT gets a synthetic BaseUtils subclass (sources / labels / classes of the fixture) whose load_baseline_model
    returns the reference's own fgvc.models.cal.WSDAN_CAL (resnet50) with the seeded random state dict and BaseUtils.get_transform();
  * the absent openai-clip package is replaced by a module whose load() returns the oracle CLIP RN50 restatement (fp32, seeded weights) with
    openai-clip's _transform, and whose tokenize() is the repo's SyntheticTokenizer (no vocabulary exists offline);
  * utils.device = cpu and Tensor.cuda = identity ('cuda' is hard-coded at utils.py:253,301,310).
Cases: the hot-path pair (semantic + model-confidence), + filter_confidence_higher_than, per-class CLIP filtering (registered as "cub": the
reference dispatches its class prompts on the dataset name, :277-299), ALIA confidence filtering with seeded random.random() draws.
The CLIP seed is searched (first seed in range whose semantic decisions are mixed AND whose closest decision margin is wide enough for a
bf16 forward not to flip it); all margins are stored so a flipped near-tie is diagnosable.

Run in the build container:  python tests/golden/make_filter_json_golden.py
"""
import json
import logging
import os
import random
import sys
import tempfile
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import clip_rn50, ref_import  # noqa: E402
from saspa_aug_b200 import checkpoints as ck  # noqa: E402
from saspa_aug_b200.pipelines import SyntheticTokenizer  # noqa: E402
from tests import filter_fixture as fx  # noqa: E402


def install_shims(ru, du, cal, fixture_root, clip_seed, state):
    import torchvision.transforms as T

    tok = SyntheticTokenizer()
    clip_mod = sys.modules["clip.clip"]

    def load(name, device=None, jit=False):
        assert name == "RN50"
        m = clip_rn50.CLIP().eval()
        m.load_state_dict(ck.random_filter_state_dict(ck.clip_rn50_shapes(), clip_seed))
        pre = T.Compose([T.Resize(224, interpolation=T.InterpolationMode.BICUBIC), T.CenterCrop(224), lambda im: im.convert("RGB"), T.ToTensor(),
                         T.Normalize((0.48145466, 0.4578275, 0.40821073), (0.26862954, 0.26130258, 0.27577711))])  # openai-clip clip.py _transform
        return m, pre

    clip_mod.load = load
    clip_mod.tokenize = lambda texts: tok([texts] if isinstance(texts, str) else list(texts))
    ru.device = torch.device("cpu")
    du.device = torch.device("cpu")

    class SyntheticRefUtils(du.BaseUtils):
        def __init__(self, split="train", root_path=fixture_root, print_func=print):
            super().__init__(split, root_path, print_func=print_func)
            self.name = "synthetic"
            self.meta_class = "airplane"
            self.images_path = self.root_path / "images"
            self.original_images_paths = [str(self.images_path / n) for n in fx.source_names()]

        def get_classes(self):
            return [f"class_{i}" for i in range(fx.NUM_CLASSES)]

        def get_image_path_to_class_id_dict(self, split="train"):
            return dict(zip(self.original_images_paths, fx.labels()))

        def get_image_path_to_class_str_dict(self):
            return {p: f"class_{c}" for p, c in zip(self.original_images_paths, fx.labels())}

        def get_basic_prompt(self):
            return "a photo of an airplane"

        def load_baseline_model(self, resize=(224, 224)):
            m = cal.WSDAN_CAL(num_classes=fx.NUM_CLASSES, net="resnet50", print_func=lambda *a: None)
            m.load_state_dict(ck.random_filter_state_dict(ck.wsdan_shapes(fx.NUM_CLASSES, "resnet50"), fx.WSDAN_SEED))
            m.eval()
            real = m.forward

            def fwd(x):  # record the logits the writer sees (margins for the golden file); arithmetic untouched
                with torch.no_grad():
                    out = real(x)
                state["logits"].append(out[0][0].clone())
                return out

            m.forward = fwd
            return m, self.get_transform(resize=resize)

        def get_baseline_conf_threshold(self):
            return {str(c): state["alia_threshold"] for c in range(fx.NUM_CLASSES)}

    du.DS_UTILS_DICT["synthetic"] = SyntheticRefUtils
    du.DS_UTILS_DICT["cub"] = SyntheticRefUtils  # clip_filtering dispatches its prompts on the dataset NAME (utils.py:294-296)
    return SyntheticRefUtils


def run_case(ru, dataset, out_dir, **kw):
    jp = ru.create_json_of_image_name_to_augmented_images_paths(dataset, out_dir, init_log=False, **kw)
    return os.path.basename(jp), json.load(open(jp)), open(jp).read()


def select():
    """One-off selection of the fixture's augmentation images and source labels from the REFERENCE nets' fp32 outputs (see the comment in
    tests/filter_fixture.py): prints the two constants to paste there."""
    from PIL import Image

    ru = ref_import.import_reference_utils()
    cal = ref_import.import_reference_cal()
    import all_utils.dataset_utils as du

    torch.set_num_threads(os.cpu_count())
    state = {"logits": [], "alia_threshold": 0.0}
    Utils = install_shims(ru, du, cal, "/tmp/unused", fx.CLIP_SEED, state)
    clip_mod = sys.modules["clip.clip"]
    model, pre = clip_mod.load("RN50")
    sel = ru.CLIP_selector(model, pre, pre, clip_mod.tokenize(["a photo of an airplane", "a photo of an object", "a photo of a scene",
                                                               "a photo of geometric shapes", "a photo", "an image", "a black photo"]))
    n_cand = 168
    imgs = [Image.fromarray(fx.candidate_image(c)) for c in range(n_cand)]
    with torch.no_grad():
        lg = torch.cat([sel(pre(im).unsqueeze(0)) for im in imgs])
    m = lg[:, 0] - lg[:, 1:].max(dim=1).values
    T = float(os.environ.get("SEM_MARGIN", "0.05"))
    print("semantic margins: keep", int((m > T).sum()), "drop", int((m < -T).sum()), "ambiguous", int((m.abs() <= T).sum()))
    print("quantiles of l0 - max other:", [round(float(q), 4) for q in torch.quantile(m, torch.tensor([0.0, 0.1, 0.25, 0.5, 0.75, 0.9, 1.0]))])
    for kind in range(len(fx.KINDS)):
        idx = [c for c in range(n_cand) if (c // 3) % len(fx.KINDS) == kind]
        print("  kind", fx.KINDS[kind], "mean margin", round(float(m[idx].mean()), 4), "std", round(float(m[idx].std()), 4))
    clf, tf = Utils().load_baseline_model()
    need = fx.N_SOURCES * fx.NUM_PER_IMAGE
    keep = [c for c in range(n_cand) if m[c] > T]
    drop = [c for c in range(n_cand) if m[c] < -T]
    assert len(keep) >= 6 and len(keep) + len(drop) >= need, (len(keep), len(drop))
    chosen = []
    while len(chosen) < need:  # alternate keep / drop while both last
        src = keep if (len(chosen) % 2 == 0 and keep) or not drop else drop
        chosen.append(src.pop(0))
    state["logits"] = []
    for c in chosen:
        clf(tf(imgs[c]).unsqueeze(0))
    lo = torch.stack(state["logits"])  # [need, C]
    srt = lo.sort(dim=1, descending=True).values
    bound = (srt[:, 9] + srt[:, 10]) / 2
    labels = []
    for i in range(fx.N_SOURCES):
        own = [i * fx.NUM_PER_IMAGE + j for j in range(fx.NUM_PER_IMAGE)]
        if i == 1:  # "syn_1" is a substring of syn_10 / syn_11: their augmentations are scored against syn_1's label too
            own += [k * fx.NUM_PER_IMAGE + j for k in (10, 11) for j in range(fx.NUM_PER_IMAGE)]
        best, best_score = None, -1.0
        want = ("mixed", "in", "out")[i % 3]  # spread the outcomes over the sources
        for lab in range(fx.NUM_CLASSES):
            d = lo[own, lab] - bound[own]
            score = d.abs().min().item()
            pattern = "in" if (d > 0).all() else "out" if (d < 0).all() else "mixed"
            if pattern == want and score > best_score:
                best, best_score = lab, score
        if best is None or best_score < 0.1:  # the wanted pattern is not available with a safe margin: take the safest label
            best = max(range(fx.NUM_CLASSES), key=lambda lab: (lo[own, lab] - bound[own]).abs().min().item())
            best_score = (lo[own, best] - bound[own]).abs().min().item()
        labels.append(best)
        print(f"source {i}: label {best}, min classifier margin {best_score:.3f}, in_topk {[bool(x) for x in (lo[own, best] > bound[own])]}")
    print("AUG_CANDIDATES =", chosen)
    print("LABELS =", labels)
    print("min semantic margin", m[chosen].abs().min().item())


def main():
    if "--select" in sys.argv:
        return select()
    logging.disable(logging.CRITICAL)
    ru = ref_import.import_reference_utils()
    cal = ref_import.import_reference_cal()
    import all_utils.dataset_utils as du

    torch.Tensor.cuda = lambda self, *a, **k: self
    torch.set_num_threads(os.cpu_count())
    tmp = tempfile.mkdtemp(prefix="saspa_filter_golden_")
    root, out_dir = fx.build(tmp)
    state = {"logits": [], "alia_threshold": 0.0}
    golden = {"fixture": "tests/filter_fixture.py", "wsdan_seed": fx.WSDAN_SEED, "num_classes": fx.NUM_CLASSES, "cases": {}}

    golden["clip_seed"] = fx.CLIP_SEED
    install_shims(ru, du, cal, root, fx.CLIP_SEED, state)

    # ---- case 1: the hot-path pair (run_aug.py:721-733) ----
    state["logits"] = []
    name, body, raw = run_case(ru, "synthetic", out_dir, semantic_filtering=1, model_confidence_based_filtering=1)
    listdir_order = [f for f in os.listdir(out_dir)]
    golden["listdir_order"] = listdir_order
    golden["cases"]["sem+conf"] = {"json_name": name, "kwargs": {"semantic_filtering": 1, "model_confidence_based_filtering": 1}, "body": fx.relativize(body, root),
                                   "raw_equals_json_dumps": raw == json.dumps(body)}
    lo = torch.stack(state["logits"])
    golden["n_classifier_calls"] = int(lo.shape[0])
    s = lo.sort(dim=1, descending=True).values
    golden["topk_margin_min"] = float((s[:, 9] - s[:, 10]).min())
    # thresholds for the optional filters: midway in the widest central gap of the reference's own values
    def mid(v):
        """Threshold in the WIDEST gap of the values that still leaves at least two of them on either side."""
        srt = torch.sort(v).values
        gaps = srt[1:] - srt[:-1]
        k = int(torch.argmax(gaps[1:-1])) + 1
        return float((srt[k] + srt[k + 1]) / 2)

    # (source index, file) pairs in the writer's visiting order (utils.py:343-355), to attribute the recorded net outputs
    excl = ["_source.", "_style.", "_target.", "_control.", "_original.", "_subject.", "subject_"]
    files = [f for f in listdir_order if not any(x in f for x in excl)]
    pairs = [(i, f) for i, n in enumerate(fx.source_names()) for f in files if os.path.splitext(n)[0][:40] in f]
    assert len(pairs) == lo.shape[0], (len(pairs), lo.shape)
    lab = torch.tensor([fx.labels()[i] for i, _ in pairs])
    golden["pairs"] = [[i, f] for i, f in pairs]
    golden["wsdan_logits"] = [[round(float(v), 4) for v in row] for row in lo]

    # ---- case 2: + filter_confidence_higher_than (utils.py:368-376): threshold midway in a gap of the label confidences ----
    conf = torch.softmax(lo, 1)[torch.arange(len(pairs)), lab]
    state["logits"] = []
    passed = torch.tensor([int(lab[k]) in lo[k].topk(10)[1].tolist() for k in range(len(pairs))])  # the test only runs on top-k survivors
    conf_thr = round(mid(conf[passed]), 4)
    golden["label_conf_gap"] = float((conf[passed] - conf_thr).abs().min())
    name, body, _ = run_case(ru, "synthetic", out_dir, model_confidence_based_filtering=True, filter_confidence_higher_than=conf_thr)
    golden["cases"]["conf+too_high"] = {"json_name": name, "kwargs": {"model_confidence_based_filtering": True, "filter_confidence_higher_than": conf_thr},
                                        "body": fx.relativize(body, root)}
    # ---- case 3: per-class CLIP filtering + semantic (utils.py:272-303, :383-404), dataset name "cub".  First pass records the class
    # confidences the reference computes; the discount then puts the threshold 1/C/discount midway in a gap of those values ----
    rec = []
    real_fwd = ru.CLIP_selector.forward

    def fwd(self, image):
        out = real_fwd(self, image)
        rec.append(out[0].clone())
        return out

    ru.CLIP_selector.forward = fwd
    run_case(ru, "cub", out_dir, clip_filtering="per_class", clip_filtering_discount=1)
    ru.CLIP_selector.forward = real_fwd
    cl = torch.stack(rec)  # one call per pair (no earlier filter in this configuration)
    assert cl.shape == (len(pairs), fx.NUM_CLASSES)
    rec.clear()
    ru.CLIP_selector.forward = fwd
    run_case(ru, "synthetic", out_dir, semantic_filtering=True)  # semantic logits of EVERY pair, for the stored margins
    ru.CLIP_selector.forward = real_fwd
    sl = torch.stack(rec)
    assert sl.shape == (len(pairs), 7)
    golden["semantic_logits"] = [[round(float(v), 4) for v in row] for row in sl]
    golden["class_conf"] = [round(float(v), 6) for v in torch.softmax(cl, 1)[torch.arange(len(pairs)), lab]]
    golden["label_conf"] = [round(float(v), 6) for v in conf]
    golden["max_logit"] = [round(float(v), 5) for v in lo.max(1).values]
    cconf = torch.softmax(cl, 1)[torch.arange(len(pairs)), lab]  # classnames = class_0.. => class index = label
    discount = round(1.0 / (fx.NUM_CLASSES * mid(cconf)), 4)
    golden["class_conf_gap"] = float((cconf - 1 / fx.NUM_CLASSES / discount).abs().min())
    name, body, _ = run_case(ru, "cub", out_dir, clip_filtering="per_class", clip_filtering_discount=discount, semantic_filtering=True)
    golden["cases"]["clip_per_class+sem"] = {"json_name": name, "dataset": "cub", "kwargs": {"clip_filtering": "per_class", "clip_filtering_discount": discount,
                                             "semantic_filtering": True}, "body": fx.relativize(body, root)}
    # ---- case 4: ALIA confidence filtering (utils.py:411-434), seeded global random ----
    state["alia_threshold"] = round(mid(lo.max(1).values), 4)
    golden["alia_threshold"] = state["alia_threshold"]
    golden["alia_gap"] = float((lo.max(1).values - state["alia_threshold"]).abs().min())
    golden["alia_random_seed"] = 123
    random.seed(123)
    name, body, _ = run_case(ru, "synthetic", out_dir, alia_conf_filtering=True, semantic_filtering=True)
    golden["cases"]["alia+sem"] = {"json_name": name, "kwargs": {"alia_conf_filtering": True, "semantic_filtering": True}, "body": fx.relativize(body, root)}

    for k, c in golden["cases"].items():
        kept = sum(len(v) for v in c["body"].values())
        print(f"case {k}: {c['json_name']}: kept {kept} of {sum(1 for f in listdir_order if '_prompt_' in f)} candidate files "
              f"(syn_1 row holds {len(c['body']['syn_1.png'])})")
    json.dump(golden, open(os.path.join(ROOT, "tests/golden/filter_json_golden.json"), "w"), indent=1)
    print("wrote tests/golden/filter_json_golden.json")


if __name__ == "__main__":
    main()
