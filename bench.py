#!/usr/bin/env python
"""bench.py -- headline benchmark of the SaSPA augmentation-generation hot path on B200.

Metric (BASELINE.json): augmented images/sec (512x512, 20 steps).  One "step" = one pass of the hot path over
BASELINE config 2: 64 synthetic 512x512 sources x 2 prompts = 128 augmentations, each = Canny(120/200) ->
CLIP text encode -> 20-step UniPC ControlNet-canny + SD v1.5 text2img denoise with CFG 7.5 (cond-scale 0.75)
-> VAE decode -> u8 image.  Random-init weights of the real architectures, synthetic inputs (no network).

  python bench.py --gpus N --steps K --warmup W            our arm (N>1 under torchrun, one rank per GPU)
  python bench.py --impl reference ...                      the reference's CPU path (oracle port; diffusers is not
                                                            installable offline) on the host cores, bounded sample
Prints ONE JSON line (rank 0).  See DESIGN.md "Measurement" for definitions.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "augmented images/sec (512x512, 20 steps)"
UNIT = "images/s"
GFLOP_PER_IMAGE_STEP = 2135.0  # SURVEY.md 8d / BASELINE.md 3: UNet + ControlNet, CFG pair, hoisted work removed
GFLOP_VAE_DECODE = 2515.0
GFLOP_ONE_TIME = 23.3


def workload(args):
    return {
        "workload": f"BASELINE config 2: {args.sources} synthetic 512x512 sources x {args.prompts} prompts, Canny 120/200 + SD v1.5 ControlNet-canny "
                    f"text2img, {args.num_inference_steps} UniPC steps, CFG 7.5, cond-scale 0.75, VAE decode to u8; random-init weights",
        "images_per_step": args.sources * args.prompts,
        "micro_batch": args.micro_batch,
        "num_inference_steps": args.num_inference_steps,
        "resolution": 512,
        "l2": "per-step working set (activations of one micro-batch >> 126 MB L2; 2.4 GB of weights) exceeds L2",
    }


# ----------------------------------------------------------------------------------------------------
# clocks sampler
# ----------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index: int):
        self.index, self.samples, self._stop = index, [], threading.Event()
        self.t = threading.Thread(target=self._run, daemon=True)

    def _run(self):
        while not self._stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i", str(self.index)],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.samples.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            self._stop.wait(0.2)

    def __enter__(self):
        self.t.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        self.t.join(timeout=6)

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        sm = sorted(float(s[0]) for s in self.samples if s[0].replace(".", "").isdigit())
        mx = max(float(s[1]) for s in self.samples if s[1].replace(".", "").isdigit())
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for s in self.samples for n, v in zip(names, s[3:7]) if v.lower().startswith("active")})
        pw = [float(s[2]) for s in self.samples if s[2].replace(".", "").isdigit()]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": reasons, "power_w_max": max(pw) if pw else None,
                "samples": len(self.samples)}


# ----------------------------------------------------------------------------------------------------
# our arm
# ----------------------------------------------------------------------------------------------------
def make_inputs(args, rank: int):
    """Deterministic synthetic inputs of one step for this rank (pinned host memory)."""
    import numpy as np
    import torch

    from saspa_aug_b200.synthetic import synthetic_source, synthetic_token_ids

    base = rank * args.sources
    src = np.stack([synthetic_source(base + s) for s in range(args.sources)])
    n_img = args.sources * args.prompts
    ids = torch.cat([synthetic_token_ids(2 * (base + s) + p) for s in range(args.sources) for p in range(args.prompts)])
    g = torch.Generator().manual_seed(1000 + rank)
    noise = torch.randn((n_img, 4, 64, 64), generator=g, dtype=torch.float32)
    neg_ids = synthetic_token_ids(999_999)
    return {"src": torch.from_numpy(src).pin_memory(), "ids": ids.pin_memory(), "noise": noise.pin_memory(), "neg_ids": neg_ids}


def run_step(pipe, dev_in, args, out_host=None):
    """One pass of the hot path over one step's batch with inputs resident on the device.  Returns u8 images (device)."""
    import torch

    from saspa_aug_b200 import ops

    src, ids, noise = dev_in["src"], dev_in["ids"], dev_in["noise"]
    # Canny once per source (the reference recomputes it per prompt, run_aug/run_aug.py:436-437; identical result)
    _, ctrl = ops.canny(src, 120, 200, out_channels=1, want_ctrl=True)
    neg = dev_in["neg"]
    outs = []
    P = args.prompts
    mb = args.micro_batch
    n_img = ids.shape[0]
    for i0 in range(0, n_img, mb):
        i1 = min(i0 + mb, n_img)
        text = pipe.encode_prompt_ids(ids[i0:i1])
        idx = torch.arange(i0, i1, device=src.device) // P
        c = ctrl.index_select(0, idx)  # placement: both prompts of a source share its control image
        img = pipe.generate_batch(text, neg.expand(i1 - i0, -1, -1).contiguous(), None, None, noise=noise[i0:i1], num_inference_steps=args.num_inference_steps,
                                  guidance_scale=7.5, controlnet_conditioning_scale=0.75, control_bf16=c)
        outs.append(img)
    return torch.cat(outs, 0)


def ours(args):
    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the hot path has no CPU fallback (use --impl reference for the CPU baseline)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    from saspa_aug_b200 import ops
    from saspa_aug_b200.pipelines import SaspaControlNetPipeline

    t0 = time.time()
    pipe = SaspaControlNetPipeline.random_init("sd15", seed=1234, sampler="unipc", device=dev, img2img=False)
    pipe.vae_micro_batch = args.vae_micro_batch
    build_s = time.time() - t0
    host = make_inputs(args, rank)
    dev_in = {k: v.to(dev) for k, v in host.items()}
    dev_in["neg"] = pipe.encode_prompt_ids(dev_in["neg_ids"])
    n_img = args.sources * args.prompts

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident throughput ("value") ----
    for _ in range(args.warmup):
        run_step(pipe, dev_in, args)
    barrier()
    launches0 = ops.LAUNCHES
    with ClockSampler(local) as clocks:
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(args.steps):
            run_step(pipe, dev_in, args)
        e.record()
        barrier()
    ms = s.elapsed_time(e)
    launches = ops.LAUNCHES - launches0
    tmax = torch.tensor([ms], device=dev)
    if world > 1:
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    ms_total = float(tmax.item())
    value = world * n_img * args.steps / (ms_total / 1e3)

    # ---- end to end through the public API with HOST buffers ("e2e") ----
    out_host = torch.empty((n_img, 512, 512, 3), dtype=torch.uint8).pin_memory()

    def e2e_step():
        d = {"src": host["src"].to(dev, non_blocking=True), "ids": host["ids"].to(dev, non_blocking=True), "noise": host["noise"].to(dev, non_blocking=True),
             "neg": dev_in["neg"]}
        img = run_step(pipe, d, args)
        out_host.copy_(img, non_blocking=True)

    e2e_step()
    barrier()
    k2 = max(1, min(args.steps, 2))
    t1 = time.perf_counter()
    for _ in range(k2):
        e2e_step()
    barrier()
    e2e_s = time.perf_counter() - t1
    te = torch.tensor([e2e_s], device=dev)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_value = world * n_img * k2 / float(te.item())
    h2d = host["src"].numel() + host["ids"].numel() * 8 + host["noise"].numel() * 4
    d2h = out_host.numel()

    # ---- roofline leg: CUDA events around every tcgen05 launch of one micro-batch pass ----
    roof = None
    unet_step_ms = None
    if rank == 0:
        roof, unet_step_ms = roofline_leg(pipe, dev_in, args)

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu = cpu_baseline_leg(args)

    if rank == 0:
        line = {
            "metric": METRIC, "value": round(value, 4), "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": round(ms_total / args.steps, 3), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "bf16", "data": "synthetic", "config": workload(args),
            "e2e": {"value": round(e2e_value, 4), "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h)},
            "gpu_launches": int(launches), "clocks": clocks.summary(), "roofline": roof, "cpu_baseline": cpu,
            "unet_step_ms": unet_step_ms, "model_build_s": round(build_s, 1),
        }
        emit(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def roofline_leg(pipe, dev_in, args):
    """Times every tcgen05 launch (GEMM + implicit conv) of one micro-batch generation with CUDA events on the
    launching stream; achieved = sum(algorithmic FLOPs) / sum(durations)."""
    import torch

    from saspa_aug_b200 import ops

    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("bf16_tflops_sustained", 1400.0))
    which = "measured sustained bf16 (MEASURED_PEAKS.json)" if peaks else "fallback 1.4 PFLOP/s sustained (B200_PROFILING.md)"
    mb = args.micro_batch
    ids = dev_in["ids"][:mb]
    _, ctrl = ops.canny(dev_in["src"][: max(1, mb // args.prompts)], 120, 200, want_ctrl=True)
    idx = torch.arange(0, mb, device=ids.device) // args.prompts
    c = ctrl.index_select(0, idx)
    text = pipe.encode_prompt_ids(ids)
    neg = dev_in["neg"].expand(mb, -1, -1).contiguous()
    steps = 3
    # un-instrumented timing of the whole denoise loop of one micro-batch (all num_inference_steps, so the once-per-image hoisted work --
    # ControlNet conditioning embedding, text K/V projections -- is amortised exactly as in the workload) -> UNet step ms
    loop_steps = args.num_inference_steps
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    pipe.generate_batch(text, neg, None, None, noise=dev_in["noise"][:mb], num_inference_steps=loop_steps, guidance_scale=7.5,
                        controlnet_conditioning_scale=0.75, control_bf16=c, decode=False)
    e.record()
    torch.cuda.synchronize()
    loop_ms = s.elapsed_time(e)
    ops.PROFILE = []
    pipe.generate_batch(text, neg, None, None, noise=dev_in["noise"][:mb], num_inference_steps=steps, guidance_scale=7.5,
                        controlnet_conditioning_scale=0.75, control_bf16=c, decode=True)
    torch.cuda.synchronize()
    prof, ops.PROFILE = ops.PROFILE, None
    agg = {}
    for kind, flops, a, b, _shape in prof:
        d = agg.setdefault(kind, [0.0, 0.0, 0])
        d[0] += flops
        d[1] += a.elapsed_time(b)
        d[2] += 1
    tc_flops = sum(agg[k][0] for k in ("gemm", "conv") if k in agg)
    tc_ms = sum(agg[k][1] for k in ("gemm", "conv") if k in agg)
    tc_n = sum(agg[k][2] for k in ("gemm", "conv") if k in agg)
    achieved = tc_flops / (tc_ms * 1e-3) / 1e12 if tc_ms > 0 else 0.0
    per_kind = {k: {"launches": v[2], "ms": round(v[1], 3), "tflops": round(v[0] / (v[1] * 1e-3) / 1e12, 1) if v[1] > 0 else None} for k, v in agg.items()}
    unet_step_ms = loop_ms / loop_steps
    # DRAM traffic per launch of the same kernel from the committed ncu capture of one step (never measured under bench.py)
    traffic, traffic_src = None, None
    try:
        import glob

        cands = sorted(glob.glob(os.path.join(ROOT, "profiles", "r*_launch_summary_*_dram.json")))
        if cands:
            js = json.load(open(cands[-1]))
            ks = [k for k in js["kernels"] if "gemm_tc_kernel" in k["kernel"] and k.get("dram_bytes_per_launch")]
            n = sum(k["launches"] for k in ks)
            if n:
                traffic = round(sum(k["dram_bytes_per_launch"] * k["launches"] for k in ks) / n)
                traffic_src = os.path.relpath(cands[-1], ROOT)
    except Exception:
        pass
    roof = {
        "bound": "tensor", "kernel": "gemm_tc_kernel<BN> (tcgen05 GEMM + implicit-GEMM conv), all launches of one micro-batch generation",
        "achieved": round(achieved, 1), "peak": peak, "unit": "TFLOP/s", "frac": round(achieved / peak, 4), "traffic": traffic, "traffic_unit": "DRAM bytes per launch (ncu dram__bytes_read.sum + dram__bytes_write.sum, mean over the step's gemm_tc_kernel launches)",
        "traffic_source": traffic_src, "algorithmic_flop_per_launch": round(tc_flops / max(tc_n, 1)),
        "peak_source": which, "launches_timed": tc_n, "avg_launch_ms": round(tc_ms / max(tc_n, 1), 4), "per_kind": per_kind,
        "step_tensor_frac": round(GFLOP_PER_IMAGE_STEP * mb / (unet_step_ms * 1e-3) / 1e3 / peak, 4),
        "note": "step_tensor_frac = 2135 GFLOP x micro_batch / UNet-step time / peak (whole denoise step incl. attention + memory-bound glue)",
    }
    return roof, {"micro_batch": mb, "ms": round(unet_step_ms, 3), "ms_per_image": round(unet_step_ms / mb, 4)}


# ----------------------------------------------------------------------------------------------------
# CPU baseline (oracle port of the reference's diffusers path) -- also the `--impl reference` arm
# ----------------------------------------------------------------------------------------------------
class CpuPath:
    """The reference's CPU implementation of the path, as far as it can exist offline: Canny (oracle C port of cv2.Canny, pinned to the
    reference's generate_canny) + the fp32 torch restatement of the diffusers SD v1.5 ControlNet graph (UNet + ControlNet CFG step, UniPC,
    VAE decode) on all host threads.  One `sample()` = a bounded slice of one image of the workload: Canny + `denoise_steps` of the 20
    steps + VAE decode, extrapolated to the full image (text encode omitted, < 1 %)."""

    def __init__(self, args, threads: int):
        import torch

        from oracle.diffusers_restated import models as om
        from saspa_aug_b200 import checkpoints as ck

        torch.set_num_threads(threads)
        self.args, self.threads = args, threads

        def ocfg(c):
            return om.UNetConfig(**{k: getattr(c, k) for k in om.UNetConfig.__dataclass_fields__})

        ucfg = ck.UNetConfig.sd15()
        with torch.no_grad():
            with torch.device("meta"):  # skip torch's default init (1.3 B parameters); values do not matter for timing
                self.unet, self.cn = om.UNet2DConditionModel(ocfg(ucfg)), om.ControlNetModel(ocfg(ucfg))
                self.vae = om.AutoencoderKL(om.VAEConfig.sd15())
            self.g = torch.Generator().manual_seed(0)
            for m in (self.unet, self.cn, self.vae):
                m.to_empty(device="cpu").eval()
                for p in m.parameters():
                    p.normal_(0.0, 0.02, generator=self.g)

    def sample(self, denoise_steps: int = 1):
        import numpy as np
        import torch

        from oracle import clib
        from oracle.diffusers_restated import schedulers as osched
        from saspa_aug_b200.synthetic import synthetic_source

        args, g = self.args, self.g
        with torch.no_grad():
            src = synthetic_source(0)
            t0 = time.perf_counter()
            edge = clib.canny(src, 120, 200)
            t_canny = time.perf_counter() - t0
            cond = torch.from_numpy(np.repeat(edge[None, ..., None], 3, 3).astype(np.float32) / 255.0).permute(0, 3, 1, 2)
            cond = torch.cat([cond] * 2)
            text = torch.randn((2, 77, 768), generator=g)
            lat = torch.randn((1, 4, 64, 64), generator=g)
            sched = osched.UniPCMultistepScheduler()
            sched.set_timesteps(args.num_inference_steps)
            t0 = time.perf_counter()
            for t in sched.timesteps[:denoise_steps]:
                x2 = torch.cat([lat] * 2)
                d, m = self.cn(x2, t, text, cond, 0.75)
                eps = self.unet(x2, t, text, d, m)
                eu, ec = eps.chunk(2)
                lat = sched.step(eu + 7.5 * (ec - eu), t, lat)
            t_step = (time.perf_counter() - t0) / denoise_steps
            t0 = time.perf_counter()
            self.vae.decode(lat / 0.18215)
            t_dec = time.perf_counter() - t0
        per_image = t_canny + args.num_inference_steps * t_step + t_dec
        return {"images_per_s": 1.0 / per_image, "t_canny_s": t_canny, "t_step_s": t_step, "t_decode_s": t_dec, "per_image_s": per_image}


def cpu_baseline_leg(args):
    threads = os.cpu_count() or 1
    r = CpuPath(args, threads).sample(1)
    return {"value": round(r["images_per_s"], 6), "unit": UNIT, "cores": threads, "kind": "port",
            "sample": f"1 source at 512x512: Canny + 1 of {args.num_inference_steps} UniPC CFG steps (UNet+ControlNet fp32) + VAE decode, torch CPU fp32 oracle; "
                      f"per-image time = canny {r['t_canny_s']:.3f}s + {args.num_inference_steps} x {r['t_step_s']:.2f}s + decode {r['t_decode_s']:.2f}s; text encode omitted (<1%)"}


def reference_arm(args):
    """`--impl reference`: W warm-up + exactly K timed steps, each step one bounded sample (CpuPath.sample) of the same workload on all
    host threads; rank 0 only (the other ranks of a torchrun launch exit 0 without work)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    path = CpuPath(args, threads)
    for _ in range(args.warmup):
        path.sample(1)
    vals = [path.sample(1) for _ in range(max(1, args.steps))]
    r = vals[-1]
    v = len(vals) / sum(x["per_image_s"] for x in vals)
    per_step_ms = (args.sources * args.prompts) / v * 1e3
    line = {
        "impl": "reference", "metric": METRIC, "value": round(v, 6), "unit": UNIT, "n_gpus": int(os.environ.get("WORLD_SIZE", "1")), "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": round(per_step_ms, 1), "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic", "config": workload(args),
        "cpu_baseline": {"value": round(v, 6), "unit": UNIT, "cores": threads, "kind": "port",
                         "sample": f"{len(vals)} timed steps after {args.warmup} warm-up, each a bounded sample of one image (Canny + 1 of {args.num_inference_steps} CFG "
                                   f"steps + VAE decode, fp32 torch CPU; last: step {r['t_step_s']:.2f}s, decode {r['t_decode_s']:.2f}s), extrapolated to "
                                   f"{args.num_inference_steps}-step images; the reference's diffusers stack cannot be installed offline, so this is the oracle port"},
        "e2e": {"value": round(v, 6), "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(json.dumps(line))


def _claim_stdout() -> int:
    """stdout must carry exactly ONE JSON line: library chatter written straight to fd 1 (e.g. the "NCCL version ..." banner at
    communicator creation) is sent to stderr for the rest of the run; the saved descriptor is used for the result line."""
    sys.stdout.flush()
    saved = os.dup(1)
    os.dup2(2, 1)
    return saved


_RESULT_FD = None


def emit(line: str) -> None:
    if _RESULT_FD is None:
        print(line, flush=True)
    else:
        os.write(_RESULT_FD, (line + "\n").encode())


def main():
    global _RESULT_FD
    _RESULT_FD = _claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--sources", type=int, default=64)
    ap.add_argument("--prompts", type=int, default=2)
    ap.add_argument("--micro-batch", type=int, default=32)
    ap.add_argument("--vae-micro-batch", type=int, default=8)
    ap.add_argument("--num-inference-steps", type=int, default=20)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        reference_arm(args)
    else:
        ours(args)


if __name__ == "__main__":
    main()
