"""Golden vectors for the filter path, produced by the REFERENCE's own code imported from /root/reference:
  * fgvc.models.cal.WSDAN_CAL (resnet50, 100 classes, eval) logits on preprocessed synthetic images;
  * all_utils.utils.CLIP_selector / get_semantic_filtering arithmetic (the reference's wrappers) run on the oracle
    CLIP RN50 restatement (openai-clip itself is not installable) with the reference's 7 semantic prompts as ids;
  * all_utils.utils.get_aug_json_path file names.
Inputs are regenerated from seeds by the tests; only outputs are committed (tests/golden/filter_golden.npz, .json).
Run in the build container:  python tests/golden/make_filter_golden.py
"""
import json
import os
import sys

import numpy as np
import torch
from PIL import Image

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import clip_rn50, ref_import  # noqa: E402
from saspa_aug_b200 import checkpoints as ck  # noqa: E402
from saspa_aug_b200.synthetic import synthetic_source, synthetic_token_ids  # noqa: E402

N_IMG, CLASSES = 6, 100


def main():
    ru = ref_import.import_reference_utils()
    cal = ref_import.import_reference_cal()
    import all_utils.dataset_utils as du

    torch.manual_seed(0)
    # --- classifier: reference WSDAN_CAL with the deterministic random state dict ---
    sd = ck.random_filter_state_dict(ck.wsdan_shapes(CLASSES, "resnet50"), 4242)
    model = cal.WSDAN_CAL(CLASSES, net="resnet50", print_func=lambda *a: None)
    model.load_state_dict(sd)
    model.eval()
    tf = du.BaseUtils.get_transform(None)  # Resize(256,256) -> CenterCrop(224) -> ToTensor -> Normalize (dataset_utils.py:78-85)
    imgs = [Image.fromarray(synthetic_source(500 + i)) for i in range(N_IMG)]
    with torch.no_grad():
        x = torch.stack([tf(im) for im in imgs])
        logits = model(x)[0]
    # --- CLIP semantic filter through the reference's own wrappers ---
    csd = ck.random_filter_state_dict(ck.clip_rn50_shapes(), 777)
    clip_model = clip_rn50.CLIP().eval()
    clip_model.load_state_dict(csd)
    prompt_ids = torch.cat([synthetic_token_ids(9000 + j) for j in range(7)])  # 1 basic prompt + 6 negatives (utils.py:306-310)
    import torchvision.transforms as T
    preprocess = T.Compose([T.Resize(224, interpolation=T.InterpolationMode.BICUBIC), T.CenterCrop(224), lambda im: im.convert("RGB"), T.ToTensor(),
                            T.Normalize((0.48145466, 0.4578275, 0.40821073), (0.26862954, 0.26130258, 0.27577711))])  # openai-clip _transform
    selector = ru.CLIP_selector(clip_model, preprocess, preprocess, prompt_ids)
    ru.device = torch.device("cpu")
    with torch.no_grad():
        clip_logits = torch.cat([selector(preprocess(im).unsqueeze(0)) for im in imgs])
        sem = torch.stack([ru.get_semantic_filtering(im, selector, preprocess, cls_idx=0) for im in imgs]).flatten()
    np.savez_compressed(os.path.join(ROOT, "tests/golden/filter_golden.npz"), wsdan_logits=logits.numpy(), clip_logits=clip_logits.numpy(),
                        semantic_keep=sem.numpy().astype(np.uint8), wsdan_input_sample=x[0, :, ::32, ::32].numpy())
    names = {
        "sem+conf": ru.get_aug_json_path("/x/y/images", semantic_filtering=1, model_confidence_based_filtering=1),
        "none": ru.get_aug_json_path("/x/y/images"),
        "conf_top5": ru.get_aug_json_path("/x/y/images", model_confidence_based_filtering=True, conf_top_k=5),
        "all": ru.get_aug_json_path("/x/y/images", lpips_min=0.1, lpips_max=0.7, clip_filtering="per_class", clip_filtering_discount=2, semantic_filtering=True,
                                    alia_conf_filtering=True),
    }
    json.dump({"n_img": N_IMG, "classes": CLASSES, "wsdan_seed": 4242, "clip_seed": 777, "image_seeds": [500 + i for i in range(N_IMG)],
               "prompt_id_seeds": [9000 + j for j in range(7)], "json_names": names}, open(os.path.join(ROOT, "tests/golden/filter_golden.json"), "w"), indent=1)
    print("wsdan logits", logits.shape, float(logits.std()), "clip logits", clip_logits[0].tolist(), "semantic keep", sem.tolist())


if __name__ == "__main__":
    main()
