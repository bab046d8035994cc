"""Post-generation filter + aug-JSON writer: host-side mirror of the reference's
``all_utils.utils.create_json_of_image_name_to_augmented_images_paths`` (all_utils/utils.py:221-465) and
``get_aug_json_path`` (:194-218).  Same names, argument meaning, file naming, JSON layout and error behaviour;
the two enabled filters of run_aug.py (semantic_filtering + model_confidence_based_filtering, run_aug.py:551-556,
:721-733) run batched on the B200 kernels (saspa_aug_b200.filter_nets) instead of batch-1 torch calls.

JSON contract (consumed by fgvc/datasets/aug_wrapper_dataset.py:106-186): one object, key per source image in
dataset order = ``Path(src).name``, value = list (possibly empty) of full path strings of kept augmentations in
``os.listdir`` order, written with ``json.dump`` defaults.
"""
from __future__ import annotations

import json
import logging
import os
from pathlib import Path
from typing import Callable, Dict, List, Optional, Sequence, Tuple

import numpy as np

MAX_FILE_NAME_LENGTH = 40  # all_utils/utils.py:341; run_aug.py MAX_FILENAME_LENGTH
SUBSTRINGS_TO_EXCLUDE = ["_source.", "_style.", "_target.", "_control.", "_original.", "_subject.", "subject_"]  # utils.py:246
SEMANTIC_NEGATIVE_PROMPTS = ["a photo of an object", "a photo of a scene", "a photo of geometric shapes", "a photo", "an image", "a black photo"]  # utils.py:306


def get_aug_json_path(augmented_image_folder_path, lpips_min=None, lpips_max=None, clip_filtering=False, clip_filtering_discount=1,
                      semantic_filtering=False, model_confidence_based_filtering=False, conf_top_k: int = 10,
                      filter_confidence_higher_than: int = None, alia_conf_filtering=False):
    """all_utils/utils.py:194-218."""
    name = ""
    if lpips_min:
        name += f"lpips_min_{lpips_min}-"
    if lpips_max:
        name += f"lpips_max_{lpips_max}-"
    if clip_filtering:
        name += f"clip_filtering_{clip_filtering}_discount_{clip_filtering_discount}-"
    if semantic_filtering:
        name += "semantic_filtering-"
    if model_confidence_based_filtering:
        name += f"model_confidence_based_filtering_top_{conf_top_k}_classes-"
        if filter_confidence_higher_than:
            name += f"filter_confidence_higher_than_{filter_confidence_higher_than}-"
    if alia_conf_filtering:
        name += "alia_conf_filtering-"
    name += "aug.json"
    return str(Path(augmented_image_folder_path).parent / name)


def check_folder_of_images_with_pil(folder, max_delete=20, substrings_to_exclude=()):
    """all_utils/utils.py:681-703: PIL-verify every non-excluded file of the folder, delete the corrupt ones, stop checking once
    ``max_delete`` files were deleted (the reference breaks out of its loop; it does not raise)."""
    from PIL import Image

    deleted = 0
    for name in [f for f in os.listdir(folder) if not any(s in f for s in substrings_to_exclude)]:
        p = Path(folder) / name
        try:
            with Image.open(p) as im:
                im.verify()
        except Exception:
            logging.info(f"image {p} is corrupted, deleting")
            os.remove(p)
            deleted += 1
            if deleted >= max_delete:
                logging.info(f"reached max_delete = {max_delete}, breaking")
                break
    return deleted


def load_filter_models(ds_utils, device):
    """(classifier, clip, tokenizer) for a dataset-utils object.  Two ways in, tried in this order:
      * ``ds_utils.load_filter_models(ds_utils, device)`` -- a dataset that builds the B200 nets itself (the synthetic one does);
      * the REFERENCE's interface (all_utils/dataset_utils.py:78-115): a class that only offers ``load_baseline_model()`` -> a torch
        ``WSDAN_CAL`` (or its ``state_dict``) + transform registers unchanged -- the weights are re-laid out into the B200 classifier
        (checkpoint_io.wsdan_from_reference_model); CLIP RN50 comes from ``ds_utils.load_clip(device)`` when offered, else from the
        openai-clip checkpoint named by $SASPA_CLIP_RN50 (a TorchScript / state-dict file, as clip.load reads it)."""
    if hasattr(ds_utils, "load_filter_models"):
        return ds_utils.load_filter_models(ds_utils, device)
    from . import checkpoint_io

    return checkpoint_io.filter_models_from_reference_interface(ds_utils, device)


def match_augmentations(original_images_paths: Sequence[str], all_file_names: Sequence[str], folder: str) -> Dict[str, List[str]]:
    """Matching rule of all_utils/utils.py:343-355: an augmentation belongs to a source iff stem[:40] is a SUBSTRING of its
    file name (exclusion substrings already removed).  Keys keep dataset order, values keep listdir order.

    The reference tests every (source, file) pair -- 2 x 10^8 substring tests for 10 000 sources x 20 000 files (14 s on rank 0 of the
    config-5 run).  Same relation, computed per file instead: the stems of a dataset have a handful of distinct lengths, so every
    window of each such length of a file name is looked up in a hash table of the stems of that length."""
    out: Dict[str, List[str]] = {}
    by_len: Dict[int, Dict[str, List[str]]] = {}
    for image_path in original_images_paths:
        name = Path(image_path).name
        out[name] = []
        stem = Path(name).stem[:MAX_FILE_NAME_LENGTH]
        by_len.setdefault(len(stem), {}).setdefault(stem, []).append(name)
    for f in all_file_names:
        full = str(Path(folder) / f)
        hit = set()
        for L, table in by_len.items():
            if L == 0:  # the empty string is a substring of everything
                hit.update(table[""])
                continue
            for i in range(len(f) - L + 1):
                names = table.get(f[i : i + L])
                if names:
                    hit.update(names)
        for name in hit:
            out[name].append(full)
    return out


def get_dict_of_value_counts(d: Dict[str, List[str]]) -> Dict[int, int]:
    """utils.py:468-482: histogram {n_kept: n_sources}."""
    counts: Dict[int, int] = {}
    for v in d.values():
        counts[len(v)] = counts.get(len(v), 0) + 1
    return dict(sorted(counts.items()))


def _load_batches(paths: Sequence[str], batch: int, threads: int = 8):
    """Decode PNGs on the host (PIL, as the reference does at utils.py:360,404) in equal-sized batches, streaming: sizes come from the
    image HEADERS, files are bucketed by (H, W) and decoded ``batch`` at a time by a small thread pool while the previous batch is on
    the GPU -- host memory holds two batches, not the folder.  Yields (indices into ``paths``, u8 [n, H, W, 3])."""
    from concurrent.futures import ThreadPoolExecutor

    from PIL import Image

    by_shape: Dict[Tuple[int, int], List[int]] = {}
    for i, p in enumerate(paths):
        with Image.open(p) as im:
            w, h = im.size
        by_shape.setdefault((h, w), []).append(i)
    chunks = [idx[j : j + batch] for idx in by_shape.values() for j in range(0, len(idx), batch)]

    def decode(i):
        return np.asarray(Image.open(paths[i]).convert("RGB"))

    with ThreadPoolExecutor(max_workers=threads) as ex:
        nxt = [ex.submit(decode, i) for i in chunks[0]] if chunks else []
        for c, idx in enumerate(chunks):
            cur = nxt
            nxt = [ex.submit(decode, i) for i in chunks[c + 1]] if c + 1 < len(chunks) else []
            yield idx, np.stack([f.result() for f in cur])


def create_json_of_image_name_to_augmented_images_paths(dataset, augmented_image_folder_path, lpips_min=None, lpips_max=None,
                                                        resize: Tuple = (256, 256), clip_filtering=False, clip_filtering_discount=1,
                                                        semantic_filtering=False, model_confidence_based_filtering=False, conf_top_k: int = 10,
                                                        filter_confidence_higher_than: int = None, init_log=True, alia_conf_filtering=False, *,
                                                        ds_utils=None, filter_models: Optional[Callable] = None, device="cuda", batch_size: int = 64,
                                                        return_details: bool = False, decisions: Optional[Dict[Tuple[str, str], Tuple[int, int]]] = None,
                                                        missing_decisions: Optional[Callable] = None, assume_verified: bool = False):
    """Drop-in for all_utils/utils.py:221-465 (same positional signature).  Keyword-only extras:
      ds_utils       dataset-utils object (default: saspa_aug_b200.datasets.DS_UTILS_DICT[dataset]())
      filter_models  callable(ds_utils, device) -> (WSDANClassifier | None, CLIPRN50 | None, tokenizer) (default: ds_utils.load_filter_models)
      decisions      {(source file name, augmentation path): (in_topk, semantic)} computed elsewhere (the sharded driver gathers every
                     rank's filter records and lets rank 0 write the JSON through this same function: same matching, ordering and
                     file name).  Keyed by the PAIR: the reference matches by substring (utils.py:352-354), so one file can belong to
                     several sources and is scored against each source's own label.
      assume_verified  the files were already PIL-verified (each rank of the sharded driver verifies its own shard in parallel before
                     filtering): skip the serial re-check of the whole folder
      missing_decisions  callable([(source name, path)]) -> {pair: (in_topk, semantic)} for matched pairs without a gathered record
                     (substring cross-matches, files of earlier runs); without it such pairs are logged and left out.
    """
    import torch

    from .filter_nets import AugmentationFilter

    assert not (clip_filtering and model_confidence_based_filtering), "can't use both clip_filtering and model_confidence_based_filtering"
    use_lpips = bool(lpips_min or lpips_max)
    if use_lpips and (lpips_min is None or lpips_max is None):
        # the reference evaluates `lpips_min <= d <= lpips_max` (utils.py:379): with one bound missing Python raises TypeError there
        raise TypeError("lpips_min and lpips_max must both be given ('<=' not supported between a number and None, all_utils/utils.py:379)")
    if use_lpips and decisions is not None:
        raise NotImplementedError("gathered decisions carry the two hot-path filters only; run the LPIPS filter in one process")
    extra = bool(clip_filtering or alia_conf_filtering or filter_confidence_higher_than)
    if extra and decisions is not None:
        raise NotImplementedError("gathered decisions carry the two hot-path filters only; run the optional filters in one process")
    if not augmented_image_folder_path.endswith("/images"):
        augmented_image_folder_path = str(Path(augmented_image_folder_path) / "images")
    json_path = get_aug_json_path(augmented_image_folder_path, lpips_min, lpips_max, clip_filtering, clip_filtering_discount, semantic_filtering,
                                  model_confidence_based_filtering, conf_top_k, filter_confidence_higher_than, alia_conf_filtering)
    if init_log:
        logging.info(f"log file: {json_path.replace('.json', '.log')}")
    logging.info(f"json_path = {json_path}")
    if not assume_verified:
        check_folder_of_images_with_pil(augmented_image_folder_path, max_delete=50, substrings_to_exclude=SUBSTRINGS_TO_EXCLUDE)
    if ds_utils is None:
        from .datasets import DS_UTILS_DICT

        ds_utils = DS_UTILS_DICT[dataset](print_func=logging.info)
    original_images_paths_list = ds_utils.original_images_paths
    if len(list(Path(augmented_image_folder_path).glob("*.*"))) < 10:
        logging.info(f"augmented_image_folder_path = {augmented_image_folder_path} doesn't exist or has less than 10 images")
        augmented_image_folder_path = str(Path(augmented_image_folder_path) / "images")
        if len(list(Path(augmented_image_folder_path).glob("*.*"))) < 10:
            raise FileNotFoundError(f"augmented_image_folder_path = {augmented_image_folder_path} doesn't exist or has less than 10 images")

    classifier = clip = tokenizer = None
    if (semantic_filtering or model_confidence_based_filtering or clip_filtering or alia_conf_filtering) and decisions is None:
        classifier, clip, tokenizer = (filter_models or load_filter_models)(ds_utils, device)
    prompt_ids = None
    if semantic_filtering and decisions is None:
        prompts = [ds_utils.get_basic_prompt()] + SEMANTIC_NEGATIVE_PROMPTS
        logging.info(f"using semantic filtering with prompts = {prompts}")
        prompt_ids = tokenizer(prompts)
    if model_confidence_based_filtering:
        image_path_to_class_id = ds_utils.get_image_path_to_class_id_dict()
        conf_top_k = min(conf_top_k, ds_utils.num_classes)
        logging.info(f"using model_confidence_based_filtering with conf_top_k = {conf_top_k}")
    class_prompt_ids = None
    clip_threshold = None
    if clip_filtering:  # all_utils/utils.py:269-303
        class_prompts = ds_utils.get_clip_filtering_prompts()
        clip_threshold = 1 / len(class_prompts) / clip_filtering_discount
        logging.info(f"using CLIP filtering, of type {clip_filtering}; total number of classes = {len(class_prompts)}; threhold = {clip_threshold}")
        class_prompt_ids = tokenizer(class_prompts)
    class_id_to_conf_threshold = None
    if alia_conf_filtering:  # :323-328
        logging.info("using alia_conf_filtering")
        image_path_to_class_id = ds_utils.get_image_path_to_class_id_dict()
        class_id_to_conf_threshold = ds_utils.get_baseline_conf_threshold()
    lpips_net = None
    if use_lpips:  # utils.py:269-270
        lpips_net = ds_utils.load_lpips(device) if hasattr(ds_utils, "load_lpips") else None
        if lpips_net is None:
            raise FileNotFoundError("the LPIPS filter needs the lpips AlexNet weights: give the dataset class a load_lpips(device) method "
                                    "returning filter_nets.LPIPSAlex (no such checkpoint exists offline)")
    flt = None
    if decisions is None:
        need_classifier = model_confidence_based_filtering or alia_conf_filtering
        need_clip = semantic_filtering or clip_filtering
        flt = AugmentationFilter(classifier if need_classifier else None, clip if need_clip else None, prompt_ids, conf_top_k, micro_batch=batch_size,
                                 class_prompt_ids=class_prompt_ids)

    all_file_names = [f for f in os.listdir(augmented_image_folder_path) if not any(s in f for s in SUBSTRINGS_TO_EXCLUDE)]
    matched = match_augmentations(original_images_paths_list, all_file_names, augmented_image_folder_path)
    # flatten (source, aug) pairs, run the nets batched, scatter the decisions back
    pairs: List[Tuple[str, str, int]] = []
    for image_path in original_images_paths_list:
        name = Path(image_path).name
        label = int(image_path_to_class_id[image_path]) if (model_confidence_based_filtering or alia_conf_filtering) else 0
        for p in matched[name]:
            pairs.append((name, p, label))
    pair_class_idx = None
    if clip_filtering:
        cls_of = {Path(p).name: ds_utils.get_class_index_for_image(p) for p in original_images_paths_list}
        pair_class_idx = [cls_of[name] for name, _, _ in pairs]
    in_topk = np.ones(len(pairs), np.uint8)
    sem = np.ones(len(pairs), np.uint8)
    label_conf = np.zeros(len(pairs), np.float32)
    max_logit = np.zeros(len(pairs), np.float32)
    argmax = np.zeros(len(pairs), np.int32)
    class_conf = np.ones(len(pairs), np.float32)
    unscored = np.zeros(len(pairs), bool)
    if decisions is not None:
        miss = [(name, path) for name, path, _ in pairs if (name, path) not in decisions]
        more = missing_decisions(miss) if (miss and missing_decisions is not None) else {}
        for k, (name, path, _) in enumerate(pairs):
            d = decisions.get((name, path), more.get((name, path)))
            if d is None:  # a file nobody scored (left by an earlier run) and no way to score it here: logged and left out, not fatal
                logging.info(f"no filter record for {path} (source {name}); left out of the JSON")
                unscored[k] = True
                continue
            in_topk[k], sem[k] = d
    elif (semantic_filtering or model_confidence_based_filtering or extra) and pairs and flt is not None:
        dev = torch.device(device)
        for idx, imgs in _load_batches([p[1] for p in pairs], batch_size):
            labels = torch.tensor([pairs[i][2] for i in idx], dtype=torch.int32, device=dev)
            cidx = torch.tensor([pair_class_idx[i] for i in idx], dtype=torch.int32, device=dev) if pair_class_idx is not None else None
            out = flt(torch.from_numpy(imgs).to(dev), labels, cidx)
            in_topk[idx] = out["in_topk"].cpu().numpy()
            sem[idx] = out["semantic"].cpu().numpy()
            label_conf[idx] = out["label_conf"].cpu().numpy()
            max_logit[idx] = out["max_logit"].cpu().numpy()
            argmax[idx] = out["argmax"].cpu().numpy()
            class_conf[idx] = out["class_conf"].cpu().numpy()
    lpips_d = np.zeros(len(pairs), np.float32)
    if use_lpips and pairs:  # calc_lpips_distance(source path, augmentation path, ...) (utils.py:377-381, :576-590), batched
        from PIL import Image

        dev = torch.device(device)
        src_of = {Path(p).name: p for p in original_images_paths_list}
        cache: Dict[str, "torch.Tensor"] = {}

        def prep(path):  # L -> RGB -> resize(256, 256) on the device, one image at a time (sources differ in size), cached per file
            if path not in cache:
                a = torch.from_numpy(np.array(Image.open(path).convert("RGB"))[None]).to(dev)
                cache[path] = lpips_net.preprocess(a, resize)[0]
            return cache[path]

        for i0 in range(0, len(pairs), batch_size):
            chunk = pairs[i0 : i0 + batch_size]
            a = torch.stack([prep(src_of[name]) for name, _, _ in chunk])
            b = torch.stack([prep(path) for _, path, _ in chunk])
            lpips_d[i0 : i0 + len(chunk)] = lpips_net(a, b).cpu().numpy()
            for _, path, _ in chunk:
                cache.pop(path, None)
    # Per source image, in dataset order, the reference applies (utils.py:357-434): top-k / too-high-confidence -> [LPIPS] -> CLIP class
    # confidence -> semantic -> ALIA confidence, each on the survivors of the previous one.  The ALIA filter spares a random 20 %
    # (`random.random() > 0.2`, global `random` state, drawn only when the confidence test fired): the draws happen here in the same order.
    import random

    result: Dict[str, List[str]] = {Path(p).name: [] for p in original_images_paths_list}
    n_topk = n_sem = n_high = n_clip = n_alia_ok = n_alia_wrong = n_lpips = 0
    for k, (name, path, label) in enumerate(pairs):
        if unscored[k]:
            continue
        if model_confidence_based_filtering:
            if not in_topk[k]:
                n_topk += 1
                continue
            if filter_confidence_higher_than and float(label_conf[k]) > filter_confidence_higher_than:
                n_high += 1
                continue
        if use_lpips and not (lpips_min <= float(lpips_d[k]) <= lpips_max):
            n_lpips += 1
            continue
        if clip_filtering and not float(class_conf[k]) >= clip_threshold:
            n_clip += 1
            continue
        if semantic_filtering and not sem[k]:
            n_sem += 1
            continue
        if alia_conf_filtering and float(max_logit[k]) > class_id_to_conf_threshold[str(label)] and random.random() > 0.2:
            if int(argmax[k]) == label:
                n_alia_ok += 1
            else:
                n_alia_wrong += 1
            continue
        result[name].append(path)
    Path(json_path).parent.mkdir(parents=True, exist_ok=True)
    with open(json_path, "w") as f:
        json.dump(result, f)
    logging.info(f"Finished creating json of image name to augmented images paths in: \n{json_path}")
    if semantic_filtering:
        logging.info(f"For filter = semantic_filtering, filtered {n_sem} images")
    if model_confidence_based_filtering:
        logging.info(f"For filter = not_in_top_{conf_top_k}, filtered {n_topk} images")
        if filter_confidence_higher_than:
            logging.info(f"For filter = confidence higher than {filter_confidence_higher_than}, filtered {n_high} images")
    if use_lpips:
        logging.info(f"For filter = lpips_min / lpips_max, filtered {n_lpips} images")
    if clip_filtering:
        logging.info(f"For filter = clip_filtering, filtered {n_clip} images")
    if alia_conf_filtering:
        logging.info(f"For filter = ALIA (conf higher than threshold), filtered {n_alia_ok} correct + {n_alia_wrong} wrong predictions")
    logging.info(f"dict_num_augmentations_per_image = {get_dict_of_value_counts(result)}")
    if return_details:
        return json_path, {"pairs": pairs, "in_topk": in_topk, "semantic": sem, "label_conf": label_conf, "max_logit": max_logit, "argmax": argmax,
                           "class_conf": class_conf, "lpips": lpips_d}
    return json_path
