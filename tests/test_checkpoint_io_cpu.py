"""Checkpoint ingestion (reference: from_pretrained of HF repos, run_aug.py:184-211; the baseline classifier's .pth rules,
all_utils/dataset_utils.py:87-115) and the CLIP BPE tokenizer -- host logic, no GPU."""
import collections
import json
import os

import pytest
import torch

from saspa_aug_b200 import checkpoint_io as cio
from saspa_aug_b200 import checkpoints as ck
from saspa_aug_b200.tokenizer import CLIPBPETokenizer, base_vocab, bytes_to_unicode


def test_safetensors_container_round_trip_and_against_the_library(tmp_path):
    g = torch.Generator().manual_seed(0)
    sd = {"a.weight": torch.randn((5, 3), generator=g), "b": torch.randn((7,), generator=g).to(torch.bfloat16), "c.h": torch.randn((2, 2, 2), generator=g).half(),
          "ids": torch.arange(6).reshape(2, 3), "empty": torch.zeros((0, 4))}
    p = tmp_path / "m.safetensors"
    cio.write_safetensors(p, sd, metadata={"format": "pt"})
    got = cio.read_safetensors(p)
    assert set(got) == set(sd) and all(got[k].dtype == sd[k].dtype and torch.equal(got[k], sd[k]) for k in sd)
    st = pytest.importorskip("safetensors.torch")
    lib = st.load_file(str(p))  # the library reads what we write ...
    assert all(torch.equal(lib[k], sd[k]) for k in sd)
    p2 = tmp_path / "lib.safetensors"
    st.save_file({k: v for k, v in sd.items()}, str(p2))  # ... and we read what the library writes
    got2 = cio.read_safetensors(p2)
    assert all(torch.equal(got2[k], sd[k]) for k in sd)
    bad = tmp_path / "bad.safetensors"
    bad.write_bytes(p.read_bytes()[:-9])
    with pytest.raises(ValueError):
        cio.read_safetensors(bad)


def _write_component(folder, sd, config, name="diffusion_pytorch_model.safetensors"):
    os.makedirs(folder, exist_ok=True)
    cio.write_safetensors(os.path.join(folder, name), sd)
    json.dump(config, open(os.path.join(folder, "config.json"), "w"))


def test_diffusers_directory_layout_is_read_like_from_pretrained(tmp_path):
    """A tiny pipeline written in the HF directory layout (unet/ vae/ text_encoder/ tokenizer/ scheduler/ + a ControlNet directory, fp16
    variant next to fp32 .bin) comes back as the state dicts / configs / tokenizer the pipeline constructors take."""
    u, v, t = ck.UNetConfig.tiny(), ck.VAEConfig.tiny(), ck.CLIPTextConfig.tiny()
    sds = {"unet": ck.random_state_dict(ck.unet_shapes(u), 1), "controlnet": ck.random_state_dict(ck.controlnet_shapes(u), 2),
           "vae": ck.random_state_dict(ck.vae_shapes(v), 3), "text": ck.random_state_dict(ck.clip_text_shapes(t), 4)}
    base, cn = tmp_path / "models--runwayml--stable-diffusion-v1-5" / "snapshots" / "abc", tmp_path / "cn"
    unet_json = dict(in_channels=4, out_channels=4, block_out_channels=list(u.block_out_channels), down_block_types=list(u.down_block_types),
                     up_block_types=list(u.up_block_types), layers_per_block=2, attention_head_dim=4, num_attention_heads=None,
                     cross_attention_dim=u.cross_attention_dim, norm_num_groups=32, norm_eps=1e-5, use_linear_projection=False, flip_sin_to_cos=True, freq_shift=0)
    _write_component(base / "unet", {k: w.half() for k, w in sds["unet"].items()}, unet_json, "diffusion_pytorch_model.fp16.safetensors")
    torch.save({k: w * 0 for k, w in sds["unet"].items()}, base / "unet" / "diffusion_pytorch_model.bin")  # must lose against the fp16 safetensors
    _write_component(cn, sds["controlnet"], dict(unet_json, conditioning_embedding_out_channels=list(u.conditioning_embedding_out_channels)))
    _write_component(base / "vae", sds["vae"], dict(in_channels=3, out_channels=3, latent_channels=4, block_out_channels=list(v.block_out_channels),
                                                    layers_per_block=2, norm_num_groups=32, scaling_factor=0.18215))
    _write_component(base / "text_encoder", sds["text"], dict(architectures=["CLIPTextModel"], vocab_size=t.vocab_size, hidden_size=t.hidden_size,
                                                              intermediate_size=t.intermediate_size, num_hidden_layers=t.num_hidden_layers,
                                                              num_attention_heads=t.num_attention_heads, max_position_embeddings=77, hidden_act="quick_gelu",
                                                              layer_norm_eps=1e-5, projection_dim=768), "model.safetensors")
    merges = [("a", "n</w>"), ("a", "i"), ("ai", "r")]
    os.makedirs(base / "tokenizer")
    json.dump(base_vocab(merges), open(base / "tokenizer" / "vocab.json", "w"))
    open(base / "tokenizer" / "merges.txt", "w").write("#version: 0.2\n" + "\n".join(" ".join(m) for m in merges) + "\n")
    json.dump({"pad_token": {"content": "<|endoftext|>"}}, open(base / "tokenizer" / "special_tokens_map.json", "w"))
    os.makedirs(base / "scheduler")
    json.dump({"_class_name": "PNDMScheduler", "beta_start": 0.00085, "beta_end": 0.012, "beta_schedule": "scaled_linear", "steps_offset": 1,
               "num_train_timesteps": 1000, "skip_prk_steps": True, "trained_betas": None}, open(base / "scheduler" / "scheduler_config.json", "w"))

    os.environ["HF_HUB_CACHE"] = str(tmp_path)
    try:
        assert cio.resolve_model_dir("runwayml/stable-diffusion-v1-5") == base
        assert cio.resolve_model_dir("lllyasviel/control_v11p_sd15_canny") is None
    finally:
        del os.environ["HF_HUB_CACHE"]
    out = cio.load_pipeline_dir(base, cn)
    assert out["configs"]["unet"] == u and out["configs"]["vae"] == v and out["configs"]["text"] == t
    assert out["unet"]["conv_in.weight"].dtype == torch.float16 and torch.equal(out["unet"]["conv_in.weight"], sds["unet"]["conv_in.weight"].half())
    for part in ("controlnet", "vae", "text"):
        assert set(out[part]) == set(sds[part]) and all(torch.equal(out[part][k], sds[part][k]) for k in sds[part])
    assert out["scheduler"] == {"num_train_timesteps": 1000, "beta_start": 0.00085, "beta_end": 0.012, "beta_schedule": "scaled_linear", "steps_offset": 1,
                                "skip_prk_steps": True}
    tok = out["tokenizer"]
    assert isinstance(tok, CLIPBPETokenizer) and tok.pad_id == tok.eos_id and out["tokenizer_2"] is None
    assert tok.encode("An  AIR") == [tok.encoder["an</w>"], tok.encoder["ai"], tok.encoder["r</w>"]]


def test_wsdan_checkpoint_rules(tmp_path):
    """dataset_utils.py:87-115: exactly one .pth, checkpoint['state_dict'], `_orig_mod.` keys of torch.compile'd runs, ResNet-101 tried
    first with ResNet-50 as the fallback; a state dict that fits neither is an error."""
    for net in ("resnet50", "resnet101"):
        sd = ck.random_filter_state_dict(ck.wsdan_shapes(7, net), 3)
        sd["features.1.num_batches_tracked"] = torch.tensor(5)  # BatchNorm bookkeeping a real checkpoint carries
        assert cio.wsdan_net_of(sd, 7) == net
        assert cio.wsdan_net_of(cio._strip_compile_prefix({"_orig_mod." + k: v for k, v in sd.items()}), 7) == net
    with pytest.raises(RuntimeError):
        cio.wsdan_net_of(ck.random_filter_state_dict(ck.wsdan_shapes(7, "resnet50"), 3), 8)
    from saspa_aug_b200.datasets import BaseUtils

    class Planes(BaseUtils):  # a class written against the reference's BaseUtils registers unchanged
        def __init__(self, split="train", root_path="data/x", print_func=print):
            super().__init__(split, root_path, print_func=print_func)
            self.name = "planes"

        def get_classes(self):
            return ["a", "b", "c"]

    os.environ["SASPA_CHECKPOINTS"] = str(tmp_path)
    try:
        ds = Planes()
        assert ds.num_classes == 3 and ds.checkpoint_folder() == tmp_path / "planes"
        with pytest.raises(AssertionError, match="Found 0 checkpoints"):
            ds.load_baseline_model()
        tf = ds.get_transform()
        assert tf["resize"] == (256, 256) and tf["center_crop"] == (224, 224)
    finally:
        del os.environ["SASPA_CHECKPOINTS"]


def _train_bpe(words, n):
    """A few rounds of real BPE training on a toy corpus: gives a merges table with the structure of CLIP's."""
    import regex

    be = bytes_to_unicode()
    pat = regex.compile(r"""'s|'t|'re|'ve|'m|'ll|'d|[\\p{L}]+|[\\p{N}]|[^\\s\\p{L}\\p{N}]+""".replace("\\\\", "\\"), regex.I)
    vocab = collections.Counter()
    for w in words:
        for piece in pat.findall(w):
            sym = [be[b] for b in piece.encode()]
            sym[-1] += "</w>"
            vocab[tuple(sym)] += 1
    merges = []
    for _ in range(n):
        pairs = collections.Counter()
        for w, c in vocab.items():
            for i in range(len(w) - 1):
                pairs[(w[i], w[i + 1])] += c
        if not pairs:
            break
        best = max(sorted(pairs), key=lambda p: pairs[p])
        merges.append(best)
        new = collections.Counter()
        for w, c in vocab.items():
            out, i = [], 0
            while i < len(w):
                if i < len(w) - 1 and (w[i], w[i + 1]) == best:
                    out.append(w[i] + w[i + 1])
                    i += 2
                else:
                    out.append(w[i])
                    i += 1
            new[tuple(out)] += c
        vocab = new
    return merges


CORPUS = ("an airplane flying over a snowy mountain range at sunset, a painting of van gogh. a commercial airplane's wing; 747-400 jets don't fly low! "
          "über café over-exposure, under-exposure, saturated, duplicate, out of frame, lowres").lower().split()
TEXTS = ["An airplane flying over a snowy mountain range at sunset, a painting of van gogh", "a 747-400 doesn't fly LOW!!  über café", "", "sunset " * 60,
         "range's jets, (wing) &amp; café-au-lait #42", "over-exposure, under-exposure, saturated, duplicate, out of frame"]


def test_bpe_tokenizer_matches_transformers_clip_tokenizer_on_a_synthetic_merges_table():
    tr = pytest.importorskip("transformers")
    merges = _train_bpe(CORPUS, 150)
    assert len(merges) > 100
    mine = CLIPBPETokenizer.from_merges(merges)
    hf = tr.CLIPTokenizer(vocab=base_vocab(merges), merges=[tuple(m) for m in merges])
    for t in TEXTS:
        want = hf(t.replace("&amp;", "&"), padding="max_length", max_length=77, truncation=True)["input_ids"]  # HF does not html-unescape; openai-clip does
        got = mine([t])[0].tolist()
        assert got == want, t
        assert mine([t], max_length=61)[0].tolist() == hf(t.replace("&amp;", "&"), padding="max_length", max_length=61, truncation=True)["input_ids"]
    assert mine.decode(mine.encode("a snowy mountain range")) == "a snowy mountain range"


def test_bpe_tokenizer_openai_conventions(tmp_path):
    """clip.tokenize: ids DERIVED from the merges file, zero padding, RuntimeError when the text does not fit, truncate keeps EOS last."""
    import gzip

    merges = _train_bpe(CORPUS, 60)
    p = tmp_path / "bpe_simple_vocab_16e6.txt.gz"
    with gzip.open(p, "wb") as f:
        f.write(("#version: 0.2\n" + "\n".join(" ".join(m) for m in merges) + "\n").encode())
    tok = CLIPBPETokenizer.from_openai_bpe(p)
    assert tok.encoder == base_vocab(merges) and tok.bos_id == 512 + len(merges) and tok.eos_id == tok.bos_id + 1
    ids = tok.tokenize(["a photo of an aircraft", "a photo"])
    n1 = len(tok.encode("a photo"))
    assert ids.shape == (2, 77) and ids.dtype == torch.int64 and ids[1, 0] == tok.bos_id
    assert int(ids[1].argmax()) == n1 + 1 and ids[1, n1 + 1] == tok.eos_id and ids[1, n1 + 2:].abs().sum() == 0  # EOS = the highest id (utils.py:134 relies on it)
    with pytest.raises(RuntimeError):
        tok.tokenize("sunset " * 90)
    cut = tok.tokenize("sunset " * 90, truncate=True)
    assert cut[0, -1] == tok.eos_id and cut[0, 0] == tok.bos_id
    hf_style = CLIPBPETokenizer.from_merges(merges, pad_id=0)(["a photo"])  # SDXL tokenizer_2 pads with "!" (id 0)
    assert hf_style[0, n1 + 2:].sum() == 0 and hf_style[0, n1 + 1] == tok.eos_id
