"""GPU parity for the filter-side kernels: bit-exact PIL resize (integer), heads vs torch fp32."""
import hashlib
import json
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import clib
from saspa_aug_b200 import ops
from saspa_aug_b200.synthetic import synthetic_source

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(__file__), "golden")


def test_resize_pil_bit_exact_vs_oracle_and_golden(cuda_device):
    gold = json.load(open(os.path.join(G, "resize_golden.json")))
    for rec in gold["cases"][:6]:
        img = synthetic_source(rec["seed"], rec["h"], rec["w"], rec["kind"])
        t = torch.from_numpy(img[None]).cuda()
        a = ops.resize_pil(t, 256, 256, "bilinear")[0].cpu().numpy()
        assert hashlib.sha256(a.tobytes()).hexdigest() == rec["bilinear_256"]["sha256"]
        sh = rec["bicubic_224"]["shape"]
        b = ops.resize_pil(t, sh[0], sh[1], "bicubic")[0].cpu().numpy()
        assert hashlib.sha256(b.tobytes()).hexdigest() == rec["bicubic_224"]["sha256"]
    imgs = np.stack([synthetic_source(s, 300, 200, "noise") for s in range(3)])
    got = ops.resize_pil(torch.from_numpy(imgs).cuda(), 224, 150, "bicubic").cpu().numpy()
    for i in range(3):
        assert np.array_equal(got[i], clib.pil_resize(imgs[i], 224, 150, "bicubic"))
    up = ops.resize_pil(torch.from_numpy(imgs).cuda(), 600, 640, "bilinear").cpu().numpy()
    assert np.array_equal(up[1], clib.pil_resize(imgs[1], 600, 640, "bilinear"))


def test_crop_normalize(cuda_device):
    img = torch.from_numpy(np.stack([synthetic_source(s, 256, 256, "noise") for s in range(2)])).cuda()
    mean, std = (0.485, 0.456, 0.406), (0.229, 0.224, 0.225)
    got = ops.crop_normalize(img, 16, 16, 224, 224, mean, std, out_c=8)
    x = img[:, 16:240, 16:240].float() / 255.0
    ref = ((x - torch.tensor(mean, device="cuda")) / torch.tensor(std, device="cuda")).to(torch.bfloat16)
    assert torch.equal(got[..., :3], ref) and got[..., 3:].abs().max() == 0


def test_bap_fc_topk(cuda_device):
    n, hw, c, m, classes = 5, 196, 2048, 32, 100
    g = torch.Generator().manual_seed(0)
    feat = torch.randn((n, hw, c), generator=g).relu().to(torch.bfloat16).cuda()
    att = torch.randn((n, hw, m), generator=g).relu().to(torch.bfloat16).cuda()
    fm = ops.bap_head(feat, att)
    p = torch.einsum("nhm,nhc->nmc", att.float(), feat.float()) / hw
    p = torch.sign(p) * torch.sqrt(p.abs() + 1e-6)
    ref = F.normalize(p.reshape(n, -1), dim=-1) * 100.0
    assert (fm - ref).abs().max().item() < 1e-3
    w = torch.randn((classes, m * c), generator=g).cuda() * 0.01
    b = torch.randn(classes, generator=g).cuda()
    logits = ops.fc_f32(fm, w, b)
    refl = fm @ w.t() + b
    assert (logits - refl).abs().max().item() < 1e-2
    labels = torch.tensor([3, 50, 99, 0, 7], dtype=torch.int32, device="cuda")
    keep, margin = ops.topk_contains(logits, labels, 10)
    top = logits.topk(10)[1]
    refk = torch.tensor([int(labels[i].item() in top[i].tolist()) for i in range(n)], dtype=torch.uint8, device="cuda")
    assert torch.equal(keep, refk)
    # ties: equal logits -> lower index wins, as torch.topk
    tie = torch.zeros((2, 20), device="cuda")
    k2, _ = ops.topk_contains(tie, torch.tensor([9, 10], dtype=torch.int32, device="cuda"), 10)
    assert k2.tolist() == [1, 0]


def test_clip_score_argmax(cuda_device):
    g = torch.Generator().manual_seed(1)
    img, txt = torch.randn((33, 1024), generator=g).cuda(), torch.randn((7, 1024), generator=g).cuda()
    logits, arg = ops.clip_score_argmax(img, txt, 100.0)
    ref = 100.0 * F.normalize(img, dim=-1) @ F.normalize(txt, dim=-1).t()
    assert (logits - ref).abs().max().item() < 1e-3
    assert torch.equal(arg.long(), ref.argmax(-1))


def _filter_fixture():
    import json

    from saspa_aug_b200 import checkpoints as ck
    from saspa_aug_b200.filter_nets import CLIPRN50, AugmentationFilter, WSDANClassifier
    from saspa_aug_b200.synthetic import synthetic_token_ids

    meta = json.load(open(os.path.join(G, "filter_golden.json")))
    gold = np.load(os.path.join(G, "filter_golden.npz"))
    wsd = ck.random_filter_state_dict(ck.wsdan_shapes(meta["classes"], "resnet50"), meta["wsdan_seed"])
    csd = ck.random_filter_state_dict(ck.clip_rn50_shapes(), meta["clip_seed"])
    ids = torch.cat([synthetic_token_ids(s) for s in meta["prompt_id_seeds"]])
    flt = AugmentationFilter(WSDANClassifier(wsd, meta["classes"], "resnet50"), CLIPRN50(csd), ids, conf_top_k=10)
    return meta, gold, flt, wsd, csd, ids


def test_filter_nets_match_reference_golden(cuda_device):
    """WSDAN_CAL logits and CLIP_selector logits of the REFERENCE code (tests/golden/make_filter_golden.py) on the
    same seeds; bf16 trunk vs the reference's fp32: tolerance 5e-2 on logits with std ~0.4 / CLIP logits O(1)."""
    meta, gold, flt, *_ = _filter_fixture()
    imgs = torch.from_numpy(np.stack([synthetic_source(s) for s in meta["image_seeds"]])).cuda()
    r = ops.resize_pil(imgs, 256, 256, "bilinear")
    x = ops.crop_normalize(r, 16, 16, 224, 224, (0.485, 0.456, 0.406), (0.229, 0.224, 0.225), out_c=8)
    logits = flt.classifier(x).cpu().numpy()
    err = np.abs(logits - gold["wsdan_logits"]).max()
    print("wsdan logits max err", err, "std", gold["wsdan_logits"].std())
    assert err < 5e-2, err
    out = flt(imgs, torch.zeros(len(meta["image_seeds"]), dtype=torch.int32, device="cuda"))
    cerr = np.abs(out["clip_logits"].cpu().numpy() - gold["clip_logits"]).max()
    print("clip logits max err", cerr)
    assert cerr < 5e-2, cerr
    assert (out["semantic"].cpu().numpy() == gold["semantic_keep"]).all()


def test_filter_decisions_identical_to_oracle(cuda_device):
    """Keep/drop decisions on 48 synthetic images vs the fp32 oracle nets (WSDAN restatement pinned to the reference;
    CLIP restatement).  Decisions must be identical; margins are printed so a near-tie flip is diagnosable."""
    from oracle import clip_rn50, wsdan
    from tests.test_filter_oracle_cpu import preprocess_baseline, preprocess_clip
    import torch.nn.functional as F

    meta, gold, flt, wsd, csd, ids = _filter_fixture()
    n = 48
    src = np.stack([synthetic_source(700 + i, kind=("blobs", "noise", "smooth")[i % 3]) for i in range(n)])
    labels = torch.arange(n, dtype=torch.int32) % meta["classes"]
    om = wsdan.WSDANOracle(meta["classes"], "resnet50").eval()
    om.load_state_dict(wsd, strict=False)
    oc = clip_rn50.CLIP().eval()
    oc.load_state_dict(csd)
    with torch.no_grad():
        lo = torch.cat([om(torch.stack([preprocess_baseline(s) for s in src[i:i + 8]])) for i in range(0, n, 8)])
        fi = torch.cat([oc.encode_image(torch.stack([preprocess_clip(s) for s in src[i:i + 8]])) for i in range(0, n, 8)])
        ft = oc.encode_text(ids)
        cl = oc.logit_scale.exp() * F.normalize(fi, dim=-1) @ F.normalize(ft, dim=-1).t()
    keep_topk = wsdan.in_topk(lo, labels.tolist(), 10)
    keep_sem = (cl.argmax(-1) == 0).to(torch.uint8)
    out = flt(torch.from_numpy(src).cuda(), labels.cuda())
    srt = lo.sort(dim=-1, descending=True)[0]
    gap = (srt[:, 9] - srt[:, 10]).min().item()
    top2 = cl.topk(2, dim=-1)[0]
    print(f"min 10th-11th logit gap {gap:.4g}; min CLIP top1-top2 gap {(top2[:, 0] - top2[:, 1]).min().item():.4g}; kept {int(keep_topk.sum())}/{n} topk, {int(keep_sem.sum())}/{n} semantic")
    assert torch.equal(out["in_topk"].cpu(), keep_topk)
    assert torch.equal(out["semantic"].cpu(), keep_sem)
    assert torch.equal(out["keep"].cpu(), keep_topk & keep_sem)


@pytest.mark.parametrize("name", ["tiny", "vit_l14"])
def test_clip_vit_matches_oracle(cuda_device, name):
    """CLIP with a VisionTransformer tower (BASELINE config 5: ViT-L/14) vs the fp32 oracle restatement (pinned to transformers.CLIPModel
    in tests/test_filter_oracle_cpu.py): image / text features within max(2 x stock-torch-bf16's own error, 1e-2 x max|ref|)."""
    import copy

    from oracle import clip_rn50
    from saspa_aug_b200 import checkpoints as ck
    from saspa_aug_b200.filter_nets import CLIPViT
    from saspa_aug_b200.synthetic import synthetic_token_ids
    from tests.test_models_gpu import _bound

    torch.set_num_threads(max(torch.get_num_threads(), 8))
    kw = ck.clip_vit_tiny_kwargs() if name == "tiny" else {}
    sd = ck.random_filter_state_dict(ck.clip_vit_shapes(**kw), 9)
    oc = clip_rn50.clip_vit(**kw).eval()
    oc.load_state_dict(sd)
    R = kw.get("res", 224)
    vocab = kw.get("vocab", 49408)
    x = torch.randn((3, 3, R, R), generator=torch.Generator().manual_seed(0))
    ids = torch.cat([synthetic_token_ids(s, vocab=vocab) for s in range(7)])
    with torch.no_grad():
        fi, ft = oc.encode_image(x), oc.encode_text(ids)
        ob = copy.deepcopy(oc).to("cuda", torch.bfloat16)
        fib, ftb = ob.encode_image(x.cuda().bfloat16()).float().cpu(), ob.encode_text(ids.cuda()).float().cpu()
    m = CLIPViT(sd)
    assert m.resolution == R
    gi = m.encode_image(x.permute(0, 2, 3, 1).contiguous().cuda().bfloat16()).cpu()
    gt = m.encode_text(ids.cuda()).cpu()
    e1, b1 = _bound(gi, fi, fib)
    e2, b2 = _bound(gt, ft, ftb)
    print(f"clip {name}: image feat err ours {e1:.4g} / torch-bf16 {b1:.4g}; text feat err ours {e2:.4g} / torch-bf16 {b2:.4g}")


def test_filter_semantic_decisions_identical_vit_l14(cuda_device):
    """Semantic keep/drop (argmax over [basic prompt + 6 negatives] == 0) with the ViT-L/14 CLIP on 24 synthetic images: identical to the
    fp32 oracle; the smallest top1-top2 logit gap is printed."""
    from oracle import clip_rn50
    from saspa_aug_b200 import checkpoints as ck
    from saspa_aug_b200.filter_nets import AugmentationFilter, CLIPViT
    from saspa_aug_b200.synthetic import synthetic_token_ids
    from tests.test_filter_oracle_cpu import preprocess_clip

    torch.set_num_threads(max(torch.get_num_threads(), 8))
    sd = ck.random_filter_state_dict(ck.clip_vit_shapes(), 777)
    oc = clip_rn50.clip_vit().eval()
    oc.load_state_dict(sd)
    n = 24
    src = np.stack([synthetic_source(900 + i, kind=("blobs", "noise", "smooth")[i % 3]) for i in range(n)])
    ids = torch.cat([synthetic_token_ids(s) for s in range(40, 47)])
    with torch.no_grad():
        fi = torch.cat([oc.encode_image(torch.stack([preprocess_clip(s) for s in src[i:i + 8]])) for i in range(0, n, 8)])
        cl = oc.logit_scale.exp() * F.normalize(fi, dim=-1) @ F.normalize(oc.encode_text(ids), dim=-1).t()
    flt = AugmentationFilter(None, CLIPViT(sd), ids)
    out = flt(torch.from_numpy(src).cuda(), torch.zeros(n, dtype=torch.int32, device="cuda"))
    top2 = cl.topk(2, dim=-1)[0]
    err = (out["clip_logits"].cpu() - cl).abs().max().item()
    print(f"ViT-L/14 CLIP logits max err {err:.4g}; min top1-top2 gap {(top2[:, 0] - top2[:, 1]).min().item():.4g}; kept {int((cl.argmax(-1) == 0).sum())}/{n}")
    assert torch.equal(out["semantic"].cpu(), (cl.argmax(-1) == 0).to(torch.uint8))


@pytest.mark.parametrize("name,n", [("tiny", 24), ("vit_l14", 12)])
def test_safety_checker_matches_oracle(cuda_device, name, n):
    """StableDiffusionSafetyChecker (SD v1.5's default-loaded checker): cosines vs the fp32 oracle (transformers CLIPVisionModel tower)
    and IDENTICAL flag decisions / blacked-out images; a score within 2e-3 of zero (the rule rounds to 3 decimals) would make the
    comparison ill-posed and is reported instead of silently passing."""
    from oracle import safety_checker as osc
    from saspa_aug_b200 import checkpoints as ck
    from saspa_aug_b200.filter_nets import SafetyChecker

    torch.set_num_threads(max(torch.get_num_threads(), 8))
    kw = ck.safety_checker_tiny_kwargs() if name == "tiny" else {}
    sd = ck.random_safety_checker_state_dict(ck.safety_checker_shapes(**kw), 3)
    o = osc.SafetyCheckerOracle(**kw).eval()
    o.load_state_dict(sd, strict=False)
    imgs = np.stack([synthetic_source(100 + i, 96, 128, kind=("blobs", "noise", "smooth")[i % 3]) for i in range(n)])
    x = osc.clip_image_processor(imgs, o.res)
    sp, co = o.cosines(x)
    want_img, want_flags, res = o(x, imgs)
    m = SafetyChecker(sd)
    cos = m.cosines(torch.from_numpy(imgs).cuda()).cpu()
    ref = torch.cat([sp, co], 1)
    err = (cos - ref).abs().max().item()
    margins = [abs(v) for r in res for v in list(r["special_scores"].values()) + list(r["concept_scores"].values())]
    print(f"safety checker {name}: cosine max err {err:.4g}; smallest |score| {min(margins):.4g}; flagged {sum(want_flags)}/{n}")
    assert err < 1e-2, err
    got_img, got_flags = m(torch.from_numpy(imgs).cuda())
    near = [i for i, r in enumerate(res) if min(abs(v) for v in list(r["special_scores"].values()) + list(r["concept_scores"].values())) < 2 * err + 1e-3]
    for i in range(n):
        if i in near:
            continue  # within the numerical error of the rounding boundary: decision not comparable
        assert got_flags[i] == want_flags[i], (i, res[i])
        assert np.array_equal(got_img[i].cpu().numpy(), want_img[i])
    assert len(near) <= n // 3, near


def test_pipeline_safety_checker_hook(cuda_device):
    """run_safety_checker in the pipeline call: flagged images come back black and nsfw_content_detected is reported."""
    from saspa_aug_b200 import checkpoints as ck
    from saspa_aug_b200.filter_nets import SafetyChecker
    from saspa_aug_b200.pipelines import SaspaControlNetPipeline
    from saspa_aug_b200.synthetic import synthetic_token_ids

    pipe = SaspaControlNetPipeline.random_init("tiny", seed=7, sampler="ddim", img2img=False)
    ctrl = np.zeros((4, 128, 128, 3), np.uint8)
    ctrl[:, 30:90, 64] = 255
    kw = dict(prompt_ids=synthetic_token_ids(1, batch=4, vocab=1000), negative_prompt_ids=synthetic_token_ids(2, batch=4, vocab=1000), image=ctrl,
              num_inference_steps=2, guidance_scale=7.5, output_type="np")
    plain = pipe(generator=torch.Generator().manual_seed(1), **kw)
    assert plain.nsfw_content_detected is None
    sd = ck.random_safety_checker_state_dict(ck.safety_checker_shapes(**ck.safety_checker_tiny_kwargs()), 3)
    sd["concept_embeds_weights"] = sd["concept_embeds_weights"] - 10.0  # every image trips a concept
    pipe.safety_checker = SafetyChecker(sd)
    out = pipe(generator=torch.Generator().manual_seed(1), **kw)
    assert out.nsfw_content_detected == [True] * 4 and all((im == 0).all() for im in out.images)
    sd["concept_embeds_weights"] = sd["concept_embeds_weights"] + 20.0  # none does
    pipe.safety_checker = SafetyChecker(sd)
    out = pipe(generator=torch.Generator().manual_seed(1), **kw)
    assert out.nsfw_content_detected == [False] * 4 and all(np.array_equal(a, b) for a, b in zip(out.images, plain.images))


def test_softmax_at_kernel(cuda_device):
    g = torch.Generator().manual_seed(0)
    for n, c in [(1, 7), (5, 100), (33, 196), (4, 1000)]:
        logits = (torch.randn((n, c), generator=g) * 3).cuda()
        logits[0, min(3, c - 1)] = logits[0].max() + 1.0
        logits[-1, c - 1] = logits[-1, 0] = logits[-1].max() + 2.0  # tie: the first index wins (torch.argmax)
        idx = torch.randint(0, c, (n,), generator=g).to(torch.int32).cuda()
        prob, mx, arg = ops.softmax_at(logits, idx)
        ref = torch.softmax(logits, dim=1)
        assert torch.allclose(prob, ref[torch.arange(n), idx.long()], rtol=1e-5, atol=1e-7)
        assert torch.equal(mx, logits.max(dim=1).values) and torch.equal(arg.long(), logits.argmax(dim=1))


def test_optional_filters_match_reference_rule(cuda_device, tmp_path):
    """The filters run_aug.py leaves disabled: filter_confidence_higher_than, clip_filtering (per-class CLIP softmax
    threshold 1/C/discount) and alia_conf_filtering (max logit over a per-class threshold, 20 % spared through `random.random()`), applied in
    the reference's order.  Expected JSONs come from the fp32 oracle nets + a literal restatement of all_utils/utils.py:357-434; thresholds sit
    midway between oracle values so the comparison is well-posed."""
    import random

    from PIL import Image

    from oracle import clip_rn50, wsdan
    from saspa_aug_b200 import checkpoints as ck
    from saspa_aug_b200 import filtering, run_aug
    from saspa_aug_b200.datasets import SyntheticUtils
    from saspa_aug_b200.pipelines import SyntheticTokenizer
    from tests.test_filter_oracle_cpu import preprocess_baseline, preprocess_clip

    torch.set_num_threads(max(torch.get_num_threads(), 8))
    ds = SyntheticUtils(root=str(tmp_path / "ds"), n_images=4, size=(160, 160)).materialize()
    out_dir = run_aug.output_folder(str(tmp_path / "ds"), run_aug.AugConfig())
    os.makedirs(out_dir)
    for index, p in enumerate(ds.original_images_paths):
        stem = os.path.splitext(os.path.basename(p))[0]
        for i in range(3):
            Image.fromarray(synthetic_source(500 + 3 * index + i, 160, 160, kind=("blobs", "noise", "smooth")[i])).save(
                os.path.join(out_dir, run_aug.aug_file_name(stem, f"an airplane, take {i}", i)))
    # oracle logits for every augmentation, in the writer's own (dataset, listdir) order
    names = [f for f in os.listdir(out_dir)]
    matched = filtering.match_augmentations(ds.original_images_paths, names, out_dir)
    order = [(p, a) for p in ds.original_images_paths for a in matched[os.path.basename(p)]]
    imgs = [np.asarray(Image.open(a).convert("RGB")) for _, a in order]
    labels = [ds.get_image_path_to_class_id_dict()[p] for p, _ in order]
    om = wsdan.WSDANOracle(ds.num_classes, "resnet50").eval()
    om.load_state_dict(ck.random_filter_state_dict(ck.wsdan_shapes(ds.num_classes, "resnet50"), ds.wsdan_seed), strict=False)
    oc = clip_rn50.CLIP().eval()
    oc.load_state_dict(ck.random_filter_state_dict(ck.clip_rn50_shapes(), ds.clip_seed))
    tok = SyntheticTokenizer()
    with torch.no_grad():
        lo = om(torch.stack([preprocess_baseline(a) for a in imgs]))
        fi = F.normalize(oc.encode_image(torch.stack([preprocess_clip(a) for a in imgs])), dim=-1)
        cls_logits = oc.logit_scale.exp() * fi @ F.normalize(oc.encode_text(tok(ds.get_clip_filtering_prompts())), dim=-1).t()
        sem_logits = oc.logit_scale.exp() * fi @ F.normalize(oc.encode_text(tok([ds.get_basic_prompt()] + filtering.SEMANTIC_NEGATIVE_PROMPTS)), dim=-1).t()
    conf = torch.softmax(lo, 1)[torch.arange(len(order)), torch.tensor(labels)]
    cconf = torch.softmax(cls_logits, 1)[torch.arange(len(order)), torch.tensor(labels)]
    in_topk = wsdan.in_topk(lo, labels, 10)

    def mid(v):
        s = torch.sort(v).values
        gaps = s[1:] - s[:-1]
        k = int(torch.argmax(gaps[len(s) // 4: 3 * len(s) // 4])) + len(s) // 4  # the widest gap in the middle half
        return float((s[k] + s[k + 1]) / 2)

    def expect(keep_fn):
        d = {os.path.basename(p): [] for p in ds.original_images_paths}
        for k, (p, a) in enumerate(order):
            if keep_fn(k):
                d[os.path.basename(p)].append(a)
        return d

    common = dict(init_log=False, ds_utils=ds)
    # (a) top-k + too-high confidence (utils.py:357-376)
    thr = mid(conf)
    jp = filtering.create_json_of_image_name_to_augmented_images_paths("synthetic", out_dir, model_confidence_based_filtering=True,
                                                                       filter_confidence_higher_than=thr, **common)
    assert "filter_confidence_higher_than" in os.path.basename(jp)
    assert json.load(open(jp)) == expect(lambda k: bool(in_topk[k]) and not float(conf[k]) > thr)
    # (b) per-class CLIP confidence + semantic (utils.py:186-191, :377-404); threshold 1/C/discount placed in a gap of the oracle values
    cthr = mid(cconf)
    discount = 1.0 / (ds.num_classes * cthr)
    jp = filtering.create_json_of_image_name_to_augmented_images_paths("synthetic", out_dir, clip_filtering="per_class", clip_filtering_discount=discount,
                                                                       semantic_filtering=True, **common)
    assert os.path.basename(jp).startswith("clip_filtering_per_class_discount_")
    assert json.load(open(jp)) == expect(lambda k: float(cconf[k]) >= 1 / ds.num_classes / discount and int(sem_logits[k].argmax()) == 0)
    # (c) ALIA confidence filter (utils.py:411-434): same global-`random` draws in the same order
    ds.alia_threshold = mid(lo.max(dim=1).values)
    random.seed(123)
    draws = []

    def alia_keep(k):
        if float(lo[k].max()) > ds.alia_threshold:
            draws.append(random.random())
            return not draws[-1] > 0.2
        return True

    want = expect(alia_keep)
    random.seed(123)
    jp = filtering.create_json_of_image_name_to_augmented_images_paths("synthetic", out_dir, alia_conf_filtering=True, **common)
    assert os.path.basename(jp) == "alia_conf_filtering-aug.json" and json.load(open(jp)) == want and len(draws) > 0
