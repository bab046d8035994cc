"""Filter networks on the sm_100a kernels: the baseline classifier WSDAN_CAL (ResNet-50/101 trunk + bilinear
attention pooling, reference fgvc/models/cal.py:131-213, resnet.py:61-178) and CLIP RN50 (openai-clip
ModifiedResNet + AttentionPool2d + text transformer, loaded at all_utils/utils.py:253 and wrapped by
TextEncoder / CLIP_selector at :113-166).  Eval-mode only: BatchNorm is folded into the preceding conv at load
(w' = w*g/sqrt(var+eps), b' = beta - mean*g/sqrt(var+eps)); ReLU and the residual add run in the GEMM epilogue.
Batched: the reference runs both nets at batch 1 with autograd on (all_utils/utils.py:360-361, :171-172) and
re-runs the CLIP text tower per image (:152-158); here images are batched and the text features cached."""
from __future__ import annotations

from typing import Dict, List, Optional

import torch

from . import nn as snn
from . import ops
from .checkpoints import RESNET_LAYERS
from .ops import ACT_QUICKGELU, ACT_RELU, BF16

SD = Dict[str, torch.Tensor]

IMAGENET_MEAN, IMAGENET_STD = (0.485, 0.456, 0.406), (0.229, 0.224, 0.225)           # all_utils/dataset_utils.py:83-84
CLIP_MEAN, CLIP_STD = (0.48145466, 0.4578275, 0.40821073), (0.26862954, 0.26130258, 0.27577711)  # openai-clip _transform


def _fold_bn(sd: SD, conv: str, bn: str, eps: float):
    w = sd[conv + ".weight"].float()
    g, b = sd[bn + ".weight"].float(), sd[bn + ".bias"].float()
    m, v = sd[bn + ".running_mean"].float(), sd[bn + ".running_var"].float()
    s = g / torch.sqrt(v + eps)
    return w * s[:, None, None, None], b - m * s


class ConvBN(snn.Conv):
    def __init__(self, sd: SD, conv: str, bn: str, dev, stride=1, padding=None, eps=1e-5):
        w, b = _fold_bn(sd, conv, bn, eps)
        if w.shape[1] == 3:  # RGB stems read the 8-channel (zero-padded) normalised image: pad Cin 3 -> 8 with zero weights
            w = torch.nn.functional.pad(w, (0, 0, 0, 0, 0, 5))
        super().__init__(sd, conv, dev, stride=stride, padding=padding, weight=w, bias=b)


class _Bottleneck:
    """ResNet bottleneck.  torchvision-style (stride on the 3x3) for WSDAN; CLIP-style (avg-pool anti-aliasing
    after the 3x3, avg-pool + 1x1 downsample) when ``clip`` is set."""

    def __init__(self, sd: SD, p: str, dev, stride: int, clip: bool):
        self.clip, self.stride = clip, stride
        self.c1 = ConvBN(sd, p + ".conv1", p + ".bn1", dev)
        self.c2 = ConvBN(sd, p + ".conv2", p + ".bn2", dev, stride=1 if clip else stride)
        self.c3 = ConvBN(sd, p + ".conv3", p + ".bn3", dev)
        self.down = None
        if p + ".downsample.0.weight" in sd:
            self.down = ConvBN(sd, p + ".downsample.0", p + ".downsample.1", dev, stride=1 if clip else stride, padding=0)

    def __call__(self, x):
        o = self.c1(x, act=ACT_RELU)
        o = self.c2(o, act=ACT_RELU)
        if self.clip and self.stride > 1:
            o = ops.pool2d(o, self.stride, self.stride, 0, False)
        idn = x
        if self.down is not None:
            xi = ops.pool2d(x, self.stride, self.stride, 0, False) if (self.clip and self.stride > 1) else x
            idn = self.down(xi)
        return self.c3(o, act=ACT_RELU, residual=idn, beta=1.0, act_after_residual=True)


class WSDANClassifier:
    """Eval forward of WSDAN_CAL -> logits p = fc(feature_matrix * 100) (the only output the filter reads,
    all_utils/utils.py:361)."""

    def __init__(self, sd: SD, num_classes: int, net: str = "resnet50", device="cuda"):
        dev = torch.device(device)
        sd = {k.replace("_orig_mod.", ""): v for k, v in sd.items()}  # torch.compile'd checkpoints (dataset_utils.py:101-102)
        self.dev, self.num_classes = dev, num_classes
        self.stem = ConvBN(sd, "features.0", "features.1", dev, stride=2, padding=3)
        self.blocks: List[_Bottleneck] = []
        for li, nb in enumerate(RESNET_LAYERS[net]):
            stride = 1 if li in (0, 3) else 2  # layer4 stride 1 -> total stride 16 (resnet.py:118-119)
            for b in range(nb):
                self.blocks.append(_Bottleneck(sd, f"features.{4 + li}.{b}", dev, stride if b == 0 else 1, clip=False))
        self.att = ConvBN(sd, "attentions.conv", "attentions.bn", dev, eps=1e-3)  # BasicConv2d (inception.py:374-384)
        self.fc_w = sd["fc.weight"].detach().to(dev, torch.float32).contiguous()

    def features(self, x: torch.Tensor) -> torch.Tensor:
        x = self.stem(x, act=ACT_RELU)
        x = ops.pool2d(x, 3, 2, 1, True)
        for b in self.blocks:
            x = b(x)
        return x

    def __call__(self, x: torch.Tensor) -> torch.Tensor:
        """x bf16 NHWC [n,224,224,8] (ImageNet-normalised, channels 3..7 zero) -> logits fp32 [n, classes]."""
        f = self.features(x)
        a = self.att(f, act=ACT_RELU)
        n, h, w, c = f.shape
        fm = ops.bap_head(f.view(n, h * w, c), a.view(n, h * w, a.shape[3]))
        return ops.fc_f32(fm, self.fc_w, None)


class _CLIPTextTower:
    """openai-clip text side shared by the RN50 and ViT models: token + positional embedding -> causal pre-LN transformer (QuickGELU)
    -> ln_final -> features at the EOT token (argmax id, all_utils/utils.py:134) @ text_projection."""

    def _init_text(self, sd: SD, dev, text_heads: int):
        self.text_heads = text_heads
        self.tok = snn._bf(sd["token_embedding.weight"], dev)
        self.tpos = snn._bf(sd["positional_embedding"], dev)
        self.layers = []
        i = 0
        while f"transformer.resblocks.{i}.ln_1.weight" in sd:
            q = f"transformer.resblocks.{i}."
            self.layers.append({"ln1": snn.Norm(sd, q + "ln_1", dev), "ln2": snn.Norm(sd, q + "ln_2", dev),
                                "qkv": snn.Linear(sd, "", dev, weight=sd[q + "attn.in_proj_weight"], bias=sd[q + "attn.in_proj_bias"]),
                                "o": snn.Linear(sd, q + "attn.out_proj", dev), "fc1": snn.Linear(sd, q + "mlp.c_fc", dev),
                                "fc2": snn.Linear(sd, q + "mlp.c_proj", dev)})
            i += 1
        self.ln_final = snn.Norm(sd, "ln_final", dev)
        self.text_proj_t = snn._bf(sd["text_projection"].t().contiguous(), dev)  # [embed, width] K-major
        self.logit_scale = float(sd["logit_scale"].exp())

    def encode_text(self, ids: torch.Tensor) -> torch.Tensor:
        """ids int64 [p, 77] -> fp32 [p, embed]."""
        p, t = ids.shape
        c = self.tok.shape[1]
        h = ops.add(self.tok.index_select(0, ids.reshape(-1)), self.tpos[:t].repeat(p, 1))
        for L in self.layers:
            y = ops.layernorm(h, 1e-5, L["ln1"].g, L["ln1"].b)
            qkv = L["qkv"](y).view(p, t, 3 * c)
            a = ops.attention(qkv[..., :c], qkv[..., c : 2 * c], qkv[..., 2 * c :], self.text_heads, causal=True)
            L["o"](a.view(p * t, c), out=h, residual=h, beta=1.0)
            y = ops.layernorm(h, 1e-5, L["ln2"].g, L["ln2"].b)
            L["fc2"](L["fc1"](y, act=ACT_QUICKGELU), out=h, residual=h, beta=1.0)
        x = ops.layernorm(h, 1e-5, self.ln_final.g, self.ln_final.b).view(p, t, c)
        eot = x[torch.arange(p, device=x.device), ids.argmax(dim=-1)].contiguous()  # index plumbing
        return ops.gemm(eot, self.text_proj_t, out_fp32=True)


class CLIPRN50(_CLIPTextTower):
    def __init__(self, sd: SD, device="cuda", heads: int = 32, text_heads: int = 8):
        dev = torch.device(device)
        self.dev, self.heads, self.text_heads = dev, heads, text_heads
        v = "visual."
        self.stem = [ConvBN(sd, v + "conv1", v + "bn1", dev, stride=2, padding=1), ConvBN(sd, v + "conv2", v + "bn2", dev), ConvBN(sd, v + "conv3", v + "bn3", dev)]
        self.blocks: List[_Bottleneck] = []
        li = 1
        while f"{v}layer{li}.0.conv1.weight" in sd:
            b = 0
            while f"{v}layer{li}.{b}.conv1.weight" in sd:
                self.blocks.append(_Bottleneck(sd, f"{v}layer{li}.{b}", dev, (1 if li == 1 else 2) if b == 0 else 1, clip=True))
                b += 1
            li += 1
        a = v + "attnpool."
        self.pos = snn._bf(sd[a + "positional_embedding"], dev)
        self.q = snn.Linear(sd, a + "q_proj", dev)
        self.kv = snn.Linear(sd, "", dev, weight=torch.cat([sd[a + "k_proj.weight"], sd[a + "v_proj.weight"]], 0),
                             bias=torch.cat([sd[a + "k_proj.bias"], sd[a + "v_proj.bias"]], 0))
        self.c = snn.Linear(sd, a + "c_proj", dev)
        self._init_text(sd, dev, text_heads)

    def encode_image(self, x: torch.Tensor) -> torch.Tensor:
        """x bf16 NHWC [n,224,224,8] (CLIP-normalised) -> fp32 [n, 1024]."""
        for s in self.stem:
            x = s(x, act=ACT_RELU)
        x = ops.pool2d(x, 2, 2, 0, False)
        for b in self.blocks:
            x = b(x)
        n, h, w, c = x.shape
        t = h * w
        mean = ops.pool2d(x, h, h, 0, False)  # [n,1,1,c] global mean token
        tok = torch.empty((n, t + 1, c), dtype=BF16, device=x.device)
        tok[:, 0].copy_(mean.view(n, c))  # token placement only
        tok[:, 1:].copy_(x.view(n, t, c))
        tok2 = ops.add(tok.view(n * (t + 1), c), self.pos.repeat(n, 1))
        tok3 = tok2.view(n, t + 1, c)
        q = self.q(tok3[:, 0])  # strided rows: first token of every image
        kv = self.kv(tok2).view(n, t + 1, 2 * c)
        a = ops.attention(q.view(n, 1, c), kv[..., :c], kv[..., c:], self.heads)
        return self.c(a.view(n, c), out_fp32=True)


class CLIPViT(_CLIPTextTower):
    """openai-clip VisionTransformer CLIP (ViT-L/14 for BASELINE config 5; ``clip.load`` at all_utils/utils.py:253 takes the model
    name) on the same kernels: snn.ViTEncoder for the tower, then ln_post on the class token and ``@ proj``."""

    input_channels = 3

    def __init__(self, sd: SD, device="cuda", text_heads: Optional[int] = None):
        dev = torch.device(device)
        self.dev = dev
        v = "visual."
        width = sd[v + "conv1.weight"].shape[0]
        layers = []
        i = 0
        while f"{v}transformer.resblocks.{i}.ln_1.weight" in sd:
            q = f"{v}transformer.resblocks.{i}."
            layers.append(snn.ViTEncoder._layer(sd, dev, sd[q + "attn.in_proj_weight"], sd[q + "attn.in_proj_bias"], q + "attn.out_proj", q + "ln_1", q + "ln_2",
                                                q + "mlp.c_fc", q + "mlp.c_proj"))
            i += 1
        self.tower = snn.ViTEncoder(dev, patch_w=sd[v + "conv1.weight"], class_emb=sd[v + "class_embedding"], pos_emb=sd[v + "positional_embedding"],
                                    pre_ln=snn.Norm(sd, v + "ln_pre", dev), layers=layers, post_ln=None, heads=width // 64, act=ACT_QUICKGELU, eps=1e-5)
        self.ln_post = snn.Norm(sd, v + "ln_post", dev)
        self.proj_t = snn._bf(sd[v + "proj"].t().contiguous(), dev)  # [embed, width] K-major
        self.resolution = self.tower.patch * int(round((sd[v + "positional_embedding"].shape[0] - 1) ** 0.5))
        self._init_text(sd, dev, text_heads or max(1, sd["ln_final.weight"].shape[0] // 64))

    def encode_image(self, x: torch.Tensor) -> torch.Tensor:
        """x bf16 NHWC [n,R,R,3] (CLIP-normalised) -> fp32 [n, embed]."""
        h = self.tower(x, apply_post_ln=False)
        cls = ops.layernorm(h[:, 0].contiguous(), 1e-5, self.ln_post.g, self.ln_post.b)
        return ops.gemm(cls, self.proj_t, out_fp32=True)


class SafetyChecker:
    """diffusers StableDiffusionSafetyChecker (pipelines/stable_diffusion/safety_checker.py), loaded by default with the SD v1.5
    pipelines the reference builds (run_aug.py:200-211): CLIP ViT-L/14 vision tower -> pooled (post-LayerNorm class token) ->
    visual_projection -> cosine against 3 special-care and 17 concept embeddings; an image is flagged when any
    round(cos - threshold + adjustment, 3) > 0, adjustment = 0.01 once a special-care concept fired; flagged images are blacked out.
    Feature extractor = CLIPImageProcessor: shortest side -> 224 (PIL bicubic, bit-exact kernel), center crop, CLIP mean/std."""

    def __init__(self, sd: SD, device="cuda"):
        dev = torch.device(device)
        self.dev = dev
        v = "vision_model.vision_model."
        width = sd[v + "embeddings.class_embedding"].shape[0]
        layers = []
        i = 0
        while f"{v}encoder.layers.{i}.layer_norm1.weight" in sd:
            q = f"{v}encoder.layers.{i}.self_attn."
            wqkv = torch.cat([sd[q + "q_proj.weight"], sd[q + "k_proj.weight"], sd[q + "v_proj.weight"]], 0)
            bqkv = torch.cat([sd[q + "q_proj.bias"], sd[q + "k_proj.bias"], sd[q + "v_proj.bias"]], 0)
            p = f"{v}encoder.layers.{i}."
            layers.append(snn.ViTEncoder._layer(sd, dev, wqkv, bqkv, q + "out_proj", p + "layer_norm1", p + "layer_norm2", p + "mlp.fc1", p + "mlp.fc2"))
            i += 1
        self.tower = snn.ViTEncoder(dev, patch_w=sd[v + "embeddings.patch_embedding.weight"], class_emb=sd[v + "embeddings.class_embedding"],
                                    pos_emb=sd[v + "embeddings.position_embedding.weight"], pre_ln=snn.Norm(sd, v + "pre_layrnorm", dev), layers=layers,
                                    post_ln=None, heads=max(1, width // 64), act=ACT_QUICKGELU, eps=1e-5)
        self.post_ln = snn.Norm(sd, v + "post_layernorm", dev)
        self.proj = snn._bf(sd["visual_projection.weight"], dev)
        self.resolution = self.tower.patch * int(round((sd[v + "embeddings.position_embedding.weight"].shape[0] - 1) ** 0.5))
        self.n_special = sd["special_care_embeds"].shape[0]
        self.embeds = torch.cat([sd["special_care_embeds"], sd["concept_embeds"]], 0).float().contiguous().to(dev)
        self.thresholds = torch.cat([sd["special_care_embeds_weights"], sd["concept_embeds_weights"]], 0).float().tolist()

    def cosines(self, images_u8: torch.Tensor) -> torch.Tensor:
        """u8 [n,H,W,3] (device) -> fp32 [n, 3 + 17] cosine of the projected image embedding with (special-care | concept) embeddings."""
        n, H, W, _ = images_u8.shape
        R = self.resolution
        oh, ow = (R, int(R * W / H)) if H <= W else (int(R * H / W), R)
        r = images_u8 if (oh, ow) == (H, W) else ops.resize_pil(images_u8.contiguous(), oh, ow, "bicubic")
        x = ops.crop_normalize(r, int(round((oh - R) / 2.0)), int(round((ow - R) / 2.0)), R, R, CLIP_MEAN, CLIP_STD, out_c=3)
        h = self.tower(x, apply_post_ln=False)
        pooled = ops.layernorm(h[:, 0].contiguous(), 1e-5, self.post_ln.g, self.post_ln.b)
        emb = ops.gemm(pooled, self.proj, out_fp32=True)
        cos, _ = ops.clip_score_argmax(emb, self.embeds, 1.0)
        return cos

    def decide(self, cos_rows) -> List[bool]:
        """The reference rule on host floats (exactly safety_checker.py's loop, including the round(., 3))."""
        out = []
        for row in cos_rows:
            adjustment = 0.0
            for c in range(self.n_special):
                if round(row[c] - self.thresholds[c] + adjustment, 3) > 0:
                    adjustment = 0.01
            bad = any(round(row[c] - self.thresholds[c] + adjustment, 3) > 0 for c in range(self.n_special, len(self.thresholds)))
            out.append(bad)
        return out

    @torch.no_grad()
    def __call__(self, images_u8: torch.Tensor):
        """-> (images with flagged ones zeroed, has_nsfw_concept list)."""
        cos = self.cosines(images_u8).cpu().tolist()
        flags = self.decide(cos)
        if any(flags):
            images_u8 = images_u8.clone()
            images_u8[torch.tensor(flags, device=images_u8.device)] = 0  # placement: black image (np.zeros in the reference)
        return images_u8, flags


class LPIPSAlex:
    """``lpips.LPIPS(net='alex')`` (all_utils/utils.py:269-270) on the sm_100a kernels: ScalingLayer folded into the u8 -> bf16
    normalisation, torchvision-AlexNet ``features`` (11x11 stride 4 and 5x5 stems as im2col + tcgen05 GEMM, 3x3 as implicit GEMM, ReLU in
    the epilogue, MaxPool(3, 2)), and the per-layer normalise / squared difference / ``lin`` / spatial mean in one kernel per layer.
    State dict: the lpips package's keys (``net.slice{k}.{idx}.weight|bias``, ``lin{k}.model.1.weight``)."""

    SHIFT, SCALE = (-0.030, -0.088, -0.188), (0.458, 0.448, 0.450)
    IDX = (0, 3, 6, 8, 10)
    GEOM = ((4, 2), (1, 2), (1, 1), (1, 1), (1, 1))  # (stride, padding) of the five convolutions

    def __init__(self, sd: SD, device="cuda"):
        dev = torch.device(device)
        self.dev = dev
        self.convs, self.lins = [], []
        for k, (idx, (stride, pad)) in enumerate(zip(self.IDX, self.GEOM)):
            w, b = sd[f"net.slice{k + 1}.{idx}.weight"], sd[f"net.slice{k + 1}.{idx}.bias"]
            if w.shape[1] == 3:  # the normalised image is stored with 8 channels (zero padded): vectorised im2col
                w = torch.nn.functional.pad(w, (0, 0, 0, 0, 0, 5))
            self.convs.append(snn.Conv({}, "", dev, stride=stride, padding=pad, weight=w, bias=b))
            self.lins.append(sd[f"lin{k}.model.1.weight"].reshape(-1).float().contiguous().to(dev))
        # x in [-1, 1] = 2 u / 255 - 1, then (x - shift) / scale  ==  (u / 255 - (1 + shift) / 2) / (scale / 2)
        self.mean = tuple((1.0 + s) / 2.0 for s in self.SHIFT)
        self.std = tuple(s / 2.0 for s in self.SCALE)

    def preprocess(self, img_u8: torch.Tensor, resize=(256, 256)) -> torch.Tensor:
        """u8 [n,H,W,3] (device) -> u8 [n,256,256,3]: PIL "L" -> "RGB" -> resize (PIL's default filter, bicubic), calc_lpips_distance :577-582."""
        g = ops.rgb_to_luma3(img_u8.contiguous())
        if resize and tuple(g.shape[1:3]) != (resize[1], resize[0]):
            g = ops.resize_pil(g, resize[1], resize[0], "bicubic")
        return g

    def features(self, x_u8: torch.Tensor):
        n, H, W, _ = x_u8.shape
        h = ops.crop_normalize(x_u8, 0, 0, H, W, self.mean, self.std, out_c=8)
        out = []
        for k, conv in enumerate(self.convs):
            if k in (1, 2):
                h = ops.pool2d(h, 3, 2, 0, True)
            h = conv(h, act=ACT_RELU)
            out.append(h)
        return out

    @torch.no_grad()
    def __call__(self, a_u8: torch.Tensor, b_u8: torch.Tensor) -> torch.Tensor:
        """Pre-processed u8 batches [n,S,S,3] -> LPIPS distances fp32 [n]."""
        acc = torch.zeros((a_u8.shape[0],), dtype=torch.float32, device=a_u8.device)
        for fa, fb, w in zip(self.features(a_u8), self.features(b_u8), self.lins):
            ops.lpips_layer_accum(fa.contiguous(), fb.contiguous(), w, acc)
        return acc


class AugmentationFilter:
    """The per-image decisions of all_utils/utils.py:357-434 on batches of u8 images resident on the device.
    Enabled on the reference's hot path (run_aug.py:551-556):
      in_topk   label in top-k(WSDAN logits)                                          (:357-365)
      semantic  argmax CLIP(img, [basic prompt, 6 negatives]) == 0                     (:169-177, :401-404)
    Optional filters (disabled in run_aug.py, used by run_aug_real_guidance.py) need only extra heads:
      label_conf  softmax(WSDAN logits)[label]      -> filter_confidence_higher_than     (:370-376)
      max_logit / argmax of the WSDAN logits        -> alia_conf_filtering               (:411-434)
      class_conf  softmax(CLIP logits over the class prompts)[class]  -> clip_filtering  (:186-191, :377-395)"""

    def __init__(self, classifier: Optional[WSDANClassifier], clip, prompt_ids: Optional[torch.Tensor], conf_top_k: int = 10,
                 micro_batch: int = 64, class_prompt_ids: Optional[torch.Tensor] = None):
        self.classifier, self.clip, self.micro_batch = classifier, clip, micro_batch
        self.conf_top_k = min(conf_top_k, classifier.num_classes) if classifier is not None else conf_top_k
        self.text_features = clip.encode_text(prompt_ids.to(clip.dev)) if (clip is not None and prompt_ids is not None) else None
        self.class_text_features = None
        if clip is not None and class_prompt_ids is not None:  # the text tower runs once, in chunks (the reference re-runs it per image)
            ids = class_prompt_ids.to(clip.dev)
            self.class_text_features = torch.cat([clip.encode_text(ids[i : i + 64]) for i in range(0, ids.shape[0], 64)], 0).contiguous()

    @torch.no_grad()
    def __call__(self, images_u8: torch.Tensor, labels: torch.Tensor, class_idx: Optional[torch.Tensor] = None):
        """images u8 [n,H,W,3] (device), labels int32 [n] (classifier label), class_idx int32 [n] (index into the class prompts)
        -> dict(keep, in_topk, semantic, topk_margin, clip_logits, label_conf, max_logit, argmax, class_conf)."""
        n = images_u8.shape[0]
        dev = images_u8.device
        in_topk = torch.ones(n, dtype=torch.uint8, device=dev)
        sem = torch.ones(n, dtype=torch.uint8, device=dev)
        margin = torch.zeros(n, dtype=torch.float32, device=dev)
        label_conf = torch.zeros(n, dtype=torch.float32, device=dev)
        max_logit = torch.zeros(n, dtype=torch.float32, device=dev)
        argmax = torch.zeros(n, dtype=torch.int32, device=dev)
        class_conf = torch.ones(n, dtype=torch.float32, device=dev)
        clip_logits = None
        H, W = images_u8.shape[1:3]
        for i0 in range(0, n, self.micro_batch):
            i1 = min(i0 + self.micro_batch, n)
            img = images_u8[i0:i1].contiguous()
            if self.classifier is not None:
                r = ops.resize_pil(img, 256, 256, "bilinear")                       # Resize((256,256)) on PIL
                x = ops.crop_normalize(r, 16, 16, 224, 224, IMAGENET_MEAN, IMAGENET_STD, out_c=8)  # CenterCrop(224), ToTensor, Normalize
                logits = self.classifier(x)
                lab = labels[i0:i1].contiguous()
                k, m = ops.topk_contains(logits, lab, self.conf_top_k)
                in_topk[i0:i1] = k
                margin[i0:i1] = m
                label_conf[i0:i1], max_logit[i0:i1], argmax[i0:i1] = ops.softmax_at(logits, lab)
            if self.clip is not None and (self.text_features is not None or self.class_text_features is not None):
                # Resize(R): shorter side -> R keeping aspect (bicubic), then CenterCrop(R)
                R = getattr(self.clip, "resolution", 224)
                if H <= W:
                    oh, ow = R, int(R * W / H)
                else:
                    oh, ow = int(R * H / W), R
                r = ops.resize_pil(img, oh, ow, "bicubic")
                cy, cx = int(round((oh - R) / 2.0)), int(round((ow - R) / 2.0))
                x = ops.crop_normalize(r, cy, cx, R, R, CLIP_MEAN, CLIP_STD, out_c=getattr(self.clip, "input_channels", 8))
                feats = self.clip.encode_image(x)
                if self.text_features is not None:
                    lg, arg = ops.clip_score_argmax(feats, self.text_features, self.clip.logit_scale)
                    sem[i0:i1] = (arg == 0).to(torch.uint8)
                    clip_logits = lg if clip_logits is None else torch.cat([clip_logits, lg], 0)
                if self.class_text_features is not None and class_idx is not None:
                    cl, _ = ops.clip_score_argmax(feats, self.class_text_features, self.clip.logit_scale)
                    class_conf[i0:i1], _, _ = ops.softmax_at(cl, class_idx[i0:i1].contiguous())
        return {"keep": in_topk & sem, "in_topk": in_topk, "semantic": sem, "topk_margin": margin, "clip_logits": clip_logits,
                "label_conf": label_conf, "max_logit": max_logit, "argmax": argmax, "class_conf": class_conf}
