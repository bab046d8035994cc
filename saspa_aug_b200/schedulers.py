"""Schedulers of the denoising loop, reduced to per-step coefficient tables for ONE fused kernel.

The reference selects DDIM (default), UniPC or (BLIP-Diffusion) PNDM/PLMS via
``Scheduler.from_config(pipe.scheduler.config)`` (run_aug/run_aug.py:217-228).  Every one of their
``step()`` updates is linear in {current sample x, CFG-combined eps, a few history tensors}.  Here the
host evaluates the scalar algebra in float64 once per step on *symbolic* linear expressions and hands the
resulting rows to ``saspa_cfg_sched_step`` (CFG combine + update + history bookkeeping in one launch);
diffusers launches 10-30 elementwise kernels per step for the same arithmetic.

Slots of the linear basis (inputs of the kernel):  0 = x,  1 = eps (after CFG),  2.. = history buffers.
Formulas follow diffusers 0.32.2 schedulers/scheduling_{ddim,unipc_multistep,pndm}.py with
SD v1.5's scheduler config: scaled_linear betas 0.00085..0.012, 1000 steps, steps_offset 1, "leading".
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Optional, Sequence

import numpy as np


def alphas_cumprod(num_train: int = 1000, beta_start: float = 0.00085, beta_end: float = 0.012) -> np.ndarray:
    # diffusers builds betas in float32 (torch.linspace(...)**2) and cumprods in float32
    betas = np.linspace(np.float32(beta_start) ** 0.5, np.float32(beta_end) ** 0.5, num_train, dtype=np.float32) ** 2
    return np.cumprod((1.0 - betas).astype(np.float32), dtype=np.float32).astype(np.float64)


@dataclass
class StepPlan:
    """One launch of saspa_cfg_sched_step: out[j] = sum_i coef[j][i] * in[i]."""
    inputs: List[Optional[str]]   # buffer names per slot ("x", None for eps, history names)
    outputs: List[str]            # buffer names written
    coef: List[List[float]]
    # optional clamped term (DDIM clip_sample): out[j] += clip_post[j] * clamp(sum_i clip_pre[i] * in[i], +-clip_range)
    clip_pre: Optional[List[float]] = None
    clip_post: Optional[List[float]] = None
    clip_range: float = 0.0


class _Lin:
    """Linear expression over the basis; supports the scalar algebra the schedulers need."""

    def __init__(self, v):
        self.v = np.asarray(v, dtype=np.float64)

    @staticmethod
    def basis(i, n):
        v = np.zeros(n)
        v[i] = 1.0
        return _Lin(v)

    def __add__(self, o):
        return _Lin(self.v + (o.v if isinstance(o, _Lin) else 0.0 if o == 0 else NotImplemented))

    __radd__ = __add__

    def __sub__(self, o):
        return _Lin(self.v - o.v)

    def __mul__(self, s):
        return _Lin(self.v * float(s))

    __rmul__ = __mul__

    def __truediv__(self, s):
        return _Lin(self.v / float(s))


class SchedulerBase:
    init_noise_sigma = 1.0
    order = 1
    history_buffers: Sequence[str] = ()

    def __init__(self, num_train_timesteps: int = 1000, steps_offset: int = 1, timestep_spacing: str = "leading", set_alpha_to_one: bool = False,
                 clip_sample: bool = False, clip_sample_range: float = 1.0):
        self.num_train = num_train_timesteps
        self.steps_offset = steps_offset
        self.spacing = timestep_spacing
        self.ac = alphas_cumprod(num_train_timesteps)
        # SD v1.5's scheduler_config.json pins set_alpha_to_one=False, clip_sample=False.  SD-XL-turbo ships an
        # EulerAncestral config without those keys, so DDIMScheduler.from_config (run_aug.py:228) falls back to DDIM's own
        # defaults set_alpha_to_one=True, clip_sample=True (range 1.0): see make_scheduler("ddim_sdxl_turbo").
        self.final_alpha_cumprod = 1.0 if set_alpha_to_one else self.ac[0]
        self.clip_sample, self.clip_sample_range = clip_sample, clip_sample_range
        self.timesteps: np.ndarray = np.zeros(0, dtype=np.int64)

    # img2img (diffusers get_timesteps): keep the last int(n*strength) steps
    def img2img_start(self, n: int, strength: float) -> int:
        init = min(int(n * strength), n)
        return max(n - init, 0)


class DDIMScheduler(SchedulerBase):
    def set_timesteps(self, n: int):
        self.n = n
        if self.spacing == "leading":
            ratio = self.num_train // n
            self.timesteps = (np.arange(0, n) * ratio).round()[::-1].copy().astype(np.int64) + self.steps_offset
        elif self.spacing == "trailing":
            self.timesteps = np.round(np.arange(self.num_train, 0, -self.num_train / n)).astype(np.int64) - 1
        else:
            raise ValueError(self.spacing)

    def begin(self, start_index: int = 0):
        self.idx = start_index

    def add_noise_coef(self, start_index: int):
        a = self.ac[int(self.timesteps[start_index])]
        return float(np.sqrt(a)), float(np.sqrt(1 - a))

    def plan(self, i: int) -> StepPlan:
        t = int(self.timesteps[i])
        prev_t = t - self.num_train // self.n
        a_t = self.ac[t]
        a_p = self.ac[prev_t] if prev_t >= 0 else self.final_alpha_cumprod
        x, e = _Lin.basis(0, 2), _Lin.basis(1, 2)
        x0 = (x - np.sqrt(1 - a_t) * e) / np.sqrt(a_t)
        if self.clip_sample:  # x_prev = sqrt(a_p) * clamp(x0) + sqrt(1-a_p) * e   (use_clipped_model_output=False: eps is kept)
            return StepPlan(["x", None], ["x"], [(np.sqrt(1 - a_p) * e).v.tolist()], clip_pre=x0.v.tolist(), clip_post=[float(np.sqrt(a_p))],
                            clip_range=float(self.clip_sample_range))
        nxt = np.sqrt(a_p) * x0 + np.sqrt(1 - a_p) * e
        return StepPlan(["x", None], ["x"], [nxt.v.tolist()])


class PNDMScheduler(SchedulerBase):
    """skip_prk_steps=True: PLMS.  n + 1 model evaluations (the second timestep is repeated)."""
    history_buffers = ("e1", "e2", "e3", "cur")

    def set_timesteps(self, n: int):
        self.n = n
        ratio = self.num_train // n
        _t = (np.arange(0, n) * ratio).round() + self.steps_offset
        self.timesteps = np.concatenate([_t[:-1], _t[-2:-1], _t[-1:]])[::-1].copy().astype(np.int64)

    def begin(self, start_index: int = 0):
        self.counter = 0
        self.n_ets = 0

    def plan(self, i: int) -> StepPlan:
        # basis: x, e, e1 (latest stored eps), e2, e3, cur (saved sample)
        nb = 6
        x, e, e1, e2, e3, cur = (_Lin.basis(k, nb) for k in range(nb))
        t = int(self.timesteps[i])
        ratio = self.num_train // self.n
        prev_t = t - ratio
        outs, names = [], []
        if self.counter != 1:
            # ets.append(e): shift history e3 <- e2 <- e1 <- e
            self.n_ets = min(self.n_ets + 1, 4)
            hist_new = [e, e1, e2]
        else:
            prev_t, t = t, t + ratio
            hist_new = [e1, e2, e3]
        n_ets = self.n_ets
        if n_ets == 1 and self.counter == 0:
            ee, xs = e, x
            save_cur = x
        elif n_ets == 1 and self.counter == 1:
            ee, xs = (e + e1) / 2, cur
            save_cur = cur
        elif n_ets == 2:
            ee, xs, save_cur = (3 * e - e1) / 2, x, cur
        elif n_ets == 3:
            ee, xs, save_cur = (23 * e - 16 * e1 + 5 * e2) / 12, x, cur
        else:
            ee, xs, save_cur = (55 * e - 59 * e1 + 37 * e2 - 9 * e3) * (1 / 24), x, cur
        a_t = self.ac[t]
        a_p = self.ac[prev_t] if prev_t >= 0 else self.final_alpha_cumprod
        b_t, b_p = 1 - a_t, 1 - a_p
        coeff = np.sqrt(a_p / a_t)
        denom = a_t * np.sqrt(b_p) + np.sqrt(a_t * b_t * a_p)
        nxt = coeff * xs - (a_p - a_t) / denom * ee
        self.counter += 1
        # outputs: x, e1, e2 (e3 is rewritten from e2 via a second tiny plan-free trick: 4 outputs max -> x, e1, e2, e3/cur)
        # cur only matters for counter 0 -> 1; e3 only once n_ets >= 3.  Pack: x, e1, e2, then e3 or cur.
        outs = [nxt, hist_new[0], hist_new[1]]
        names = ["x", "e1", "e2"]
        if self.counter == 1:  # just executed the very first step: save the sample
            outs.append(save_cur)
            names.append("cur")
        else:
            outs.append(hist_new[2])
            names.append("e3")
        return StepPlan(["x", None, "e1", "e2", "e3", "cur"], names, [o.v.tolist() for o in outs])


class UniPCMultistepScheduler(SchedulerBase):
    """solver_order 2, bh2, predict_x0, lower_order_final, final sigma 0, epsilon prediction."""
    history_buffers = ("m0", "m1", "last")

    def __init__(self, *a, solver_order: int = 2, **k):
        super().__init__(*a, **k)
        assert solver_order == 2, "only the reference's default order (2) is built"
        self.solver_order = solver_order

    def set_timesteps(self, n: int):
        self.n = n
        if self.spacing == "leading":
            ratio = self.num_train // (n + 1)
            ts = (np.arange(0, n + 1) * ratio).round()[::-1][:-1].copy().astype(np.int64) + self.steps_offset
        elif self.spacing == "trailing":
            ts = np.arange(self.num_train, 0, -self.num_train / n).round().copy().astype(np.int64) - 1
        else:
            raise ValueError(self.spacing)
        sig = ((1 - self.ac) / self.ac) ** 0.5
        sig = np.interp(ts, np.arange(0, len(sig)), sig)
        self.sigmas = np.concatenate([sig, [0.0]]).astype(np.float32).astype(np.float64)
        self.timesteps = ts

    @staticmethod
    def _as(sigma):
        alpha = 1.0 / np.sqrt(sigma * sigma + 1.0)
        return alpha, sigma * alpha

    def add_noise_coef(self, start_index: int):
        a, s = self._as(self.sigmas[start_index])
        return float(a), float(s)

    def begin(self, start_index: int = 0):
        self.step_index = start_index
        self.lower_order_nums = 0
        self.have_last = False
        self.n_hist = 0  # valid entries among m0, m1
        self.this_order = 1

    @staticmethod
    def _lam(sigma):
        a, s = UniPCMultistepScheduler._as(sigma)
        return np.log(a) - np.log(s)

    def _rhos(self, rks, h, order, corrector):
        hh = -h
        h_phi_1 = np.expm1(hh)
        h_phi_k = h_phi_1 / hh - 1
        B_h = np.expm1(hh)
        fact = 1
        R, b = [], []
        rks = np.asarray(rks, dtype=np.float64)
        for i in range(1, order + 1):
            R.append(rks ** (i - 1))
            b.append(h_phi_k * fact / B_h)
            fact *= i + 1
            h_phi_k = h_phi_k / hh - 1 / fact
        R, b = np.stack(R), np.asarray(b)
        if corrector:
            rhos = np.array([0.5]) if order == 1 else np.linalg.solve(R, b)
        else:
            rhos = np.array([0.5]) if order == 2 else (np.linalg.solve(R[:-1, :-1], b[:-1]) if order > 2 else np.zeros(0))
        return rhos, h_phi_1, B_h

    def plan(self, i: int) -> StepPlan:
        # basis: x, e, m0 (latest x0-pred), m1, last (last_sample)
        nb = 5
        x, e, m0, m1, last = (_Lin.basis(k, nb) for k in range(nb))
        si = self.step_index
        a_s, s_s = self._as(self.sigmas[si])
        m = (x - s_s * e) / a_s  # convert_model_output at the incoming sample
        xc = x
        if si > 0 and self.have_last:
            # UniC(m, last_sample, order = previous predictor's order); previous step index si-1 -> si
            order = self.this_order
            sig_t, sig_s0 = self.sigmas[si], self.sigmas[si - 1]
            a_t, s_t = self._as(sig_t)
            a_s0, s_s0 = self._as(sig_s0)
            h = self._lam(sig_t) - self._lam(sig_s0)
            rks, D1s = [], []
            if order == 2:
                lam_i = self._lam(self.sigmas[si - 2])
                rk = (lam_i - self._lam(sig_s0)) / h
                rks.append(rk)
                D1s.append((m1 - m0) / rk)
            rks.append(1.0)
            rhos, h_phi_1, B_h = self._rhos(rks, h, order, corrector=True)
            base = (s_t / s_s0) * last - a_t * h_phi_1 * m0
            corr = _Lin(np.zeros(nb))
            for r, d in zip(rhos[:-1], D1s):
                corr = corr + r * d
            xc = base - a_t * B_h * (corr + rhos[-1] * (m - m0))
        # history shift: m1 <- m0 <- m
        new_m1, new_m0 = m0, m
        self.n_hist = min(self.n_hist + 1, 2)
        order = min(self.solver_order, len(self.timesteps) - si)
        self.this_order = min(order, self.lower_order_nums + 1)
        order = self.this_order
        # UniP from xc with history (new_m0, new_m1)
        sig_t, sig_s0 = self.sigmas[si + 1], self.sigmas[si]
        a_t, s_t = self._as(sig_t)
        a_s0, s_s0 = self._as(sig_s0)
        if sig_t == 0.0:
            # final step: lambda_t = +inf; diffusers evaluates log(0) -> -inf => h = inf, expm1(-inf) = -1
            h_phi_1, B_h = -1.0, -1.0
            pred = (s_t / s_s0) * xc - a_t * h_phi_1 * new_m0  # = x0 prediction (order forced to 1 by lower_order_final)
            assert order == 1
            nxt = pred
        else:
            h = self._lam(sig_t) - self._lam(sig_s0)
            rks, D1s = [], []
            if order == 2:
                lam_i = self._lam(self.sigmas[si - 1])
                rk = (lam_i - self._lam(sig_s0)) / h
                rks.append(rk)
                D1s.append((new_m1 - new_m0) / rk)
            rks.append(1.0)
            rhos, h_phi_1, B_h = self._rhos(rks, h, order, corrector=False)
            nxt = (s_t / s_s0) * xc - a_t * h_phi_1 * new_m0
            if order == 2:
                nxt = nxt - a_t * B_h * (rhos[0] * D1s[0])
        if self.lower_order_nums < self.solver_order:
            self.lower_order_nums += 1
        self.have_last = True
        self.step_index += 1
        return StepPlan(["x", None, "m0", "m1", "last"], ["x", "m0", "m1", "last"], [nxt.v.tolist(), new_m0.v.tolist(), new_m1.v.tolist(), xc.v.tolist()])


def make_scheduler(name: str, **kw) -> SchedulerBase:
    name = name.lower()
    if name == "ddim":
        return DDIMScheduler(**kw)
    if name == "ddim_sdxl_turbo":  # DDIMScheduler.from_config(<sdxl-turbo EulerAncestral config>): trailing spacing + DDIM defaults
        return DDIMScheduler(**{**dict(timestep_spacing="trailing", set_alpha_to_one=True, clip_sample=True), **kw})
    if name == "unipc_sdxl_turbo":
        return UniPCMultistepScheduler(**{**dict(timestep_spacing="trailing"), **kw})
    if name in ("unipc", "unipcmultistep"):
        return UniPCMultistepScheduler(**kw)
    if name in ("pndm", "plms"):
        return PNDMScheduler(**kw)
    raise ValueError(f"unknown sampler {name!r} (reference supports ddim | unipcmultistep; BLIP keeps PNDM)")
