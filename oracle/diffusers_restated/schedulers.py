"""CPU restatement of the three diffusers 0.32.2 schedulers the reference can reach:
DDIMScheduler (reference default, run_aug/run_aug.py:128,220-221), UniPCMultistepScheduler
(`sampler="unipcmultistep"`, :218-219; BASELINE config 1) and PNDMScheduler/PLMS (BLIP-Diffusion keeps its
default scheduler, :217).  All built `.from_config(pipe.scheduler.config)` of SD v1.5's PNDM config:
beta_schedule scaled_linear 0.00085..0.012, 1000 train steps, steps_offset 1, timestep_spacing "leading",
clip_sample False, set_alpha_to_one False, skip_prk_steps True.

TEST INFRASTRUCTURE; **parity unpinned** (third-party diffusers, not installable here; formulas follow
schedulers/scheduling_{ddim,unipc_multistep,pndm}.py as published).  Host scalar math in
float64/float32 as diffusers does (numpy float64 for the tables, torch float32 for alphas_cumprod).
"""
from __future__ import annotations

import math
from typing import List, Optional

import numpy as np
import torch


def _alphas_cumprod(num_train: int = 1000, beta_start: float = 0.00085, beta_end: float = 0.012) -> torch.Tensor:
    betas = torch.linspace(beta_start ** 0.5, beta_end ** 0.5, num_train, dtype=torch.float32) ** 2
    return torch.cumprod(1.0 - betas, dim=0)


class DDIMScheduler:
    order = 1
    init_noise_sigma = 1.0

    def __init__(self, num_train_timesteps=1000, steps_offset=1, timestep_spacing="leading", set_alpha_to_one=False, clip_sample=False,
                 clip_sample_range=1.0):
        # diffusers' own DDIM defaults are set_alpha_to_one=True, clip_sample=True; SD v1.5's scheduler config pins both to
        # False (the defaults here).  sdxl-turbo's EulerAncestral config omits both keys => DDIM defaults apply there.
        self.clip_sample, self.clip_sample_range = clip_sample, clip_sample_range
        self.num_train = num_train_timesteps
        self.steps_offset = steps_offset
        self.spacing = timestep_spacing
        self.alphas_cumprod = _alphas_cumprod(num_train_timesteps)
        self.final_alpha_cumprod = torch.tensor(1.0) if set_alpha_to_one else self.alphas_cumprod[0]

    def set_timesteps(self, n: int):
        self.num_inference_steps = n
        if self.spacing == "leading":
            ratio = self.num_train // n
            ts = (np.arange(0, n) * ratio).round()[::-1].copy().astype(np.int64) + self.steps_offset
        elif self.spacing == "trailing":
            ratio = self.num_train / n
            ts = np.round(np.arange(self.num_train, 0, -ratio)).astype(np.int64) - 1
        else:
            raise ValueError(self.spacing)
        self.timesteps = torch.from_numpy(ts)

    def scale_model_input(self, x, t):
        return x

    def add_noise(self, x0, noise, t):
        a = self.alphas_cumprod[int(t)]
        return a.sqrt() * x0 + (1 - a).sqrt() * noise

    def step(self, eps, t, x):
        t = int(t)
        prev_t = t - self.num_train // self.num_inference_steps
        a_t = self.alphas_cumprod[t]
        a_prev = self.alphas_cumprod[prev_t] if prev_t >= 0 else self.final_alpha_cumprod
        x0 = (x - (1 - a_t).sqrt() * eps) / a_t.sqrt()
        if self.clip_sample:
            x0 = x0.clamp(-self.clip_sample_range, self.clip_sample_range)
        return a_prev.sqrt() * x0 + (1 - a_prev).sqrt() * eps  # eta = 0, use_clipped_model_output = False


class PNDMScheduler:
    """skip_prk_steps=True => pure PLMS."""
    order = 1
    init_noise_sigma = 1.0

    def __init__(self, num_train_timesteps=1000, steps_offset=1, set_alpha_to_one=False):
        self.num_train = num_train_timesteps
        self.steps_offset = steps_offset
        self.alphas_cumprod = _alphas_cumprod(num_train_timesteps)
        self.final_alpha_cumprod = torch.tensor(1.0) if set_alpha_to_one else self.alphas_cumprod[0]

    def set_timesteps(self, n: int):
        self.num_inference_steps = n
        ratio = self.num_train // n
        _t = (np.arange(0, n) * ratio).round() + self.steps_offset
        plms = np.concatenate([_t[:-1], _t[-2:-1], _t[-1:]])[::-1].copy()
        self.timesteps = torch.from_numpy(plms.astype(np.int64))
        self.ets: List[torch.Tensor] = []
        self.counter = 0
        self.cur_sample = None

    def scale_model_input(self, x, t):
        return x

    def step(self, eps, t, x):
        t = int(t)
        prev_t = t - self.num_train // self.num_inference_steps
        if self.counter != 1:
            self.ets = self.ets[-3:]
            self.ets.append(eps)
        else:
            prev_t = t
            t = t + self.num_train // self.num_inference_steps
        if len(self.ets) == 1 and self.counter == 0:
            e = eps
            self.cur_sample = x
        elif len(self.ets) == 1 and self.counter == 1:
            e = (eps + self.ets[-1]) / 2
            x = self.cur_sample
            self.cur_sample = None
        elif len(self.ets) == 2:
            e = (3 * self.ets[-1] - self.ets[-2]) / 2
        elif len(self.ets) == 3:
            e = (23 * self.ets[-1] - 16 * self.ets[-2] + 5 * self.ets[-3]) / 12
        else:
            e = (1 / 24) * (55 * self.ets[-1] - 59 * self.ets[-2] + 37 * self.ets[-3] - 9 * self.ets[-4])
        self.counter += 1
        return self._prev(x, t, prev_t, e)

    def _prev(self, x, t, prev_t, e):
        a_t = self.alphas_cumprod[t]
        a_prev = self.alphas_cumprod[prev_t] if prev_t >= 0 else self.final_alpha_cumprod
        b_t, b_prev = 1 - a_t, 1 - a_prev
        coeff = (a_prev / a_t) ** 0.5
        denom = a_t * b_prev ** 0.5 + (a_t * b_t * a_prev) ** 0.5
        return coeff * x - (a_prev - a_t) * e / denom


class UniPCMultistepScheduler:
    """solver_order 2, solver_type bh2, predict_x0, lower_order_final, final_sigmas_type "zero",
    prediction_type epsilon, thresholding off."""
    order = 1
    init_noise_sigma = 1.0

    def __init__(self, num_train_timesteps=1000, steps_offset=1, solver_order=2, timestep_spacing="leading"):
        self.num_train = num_train_timesteps
        self.steps_offset = steps_offset
        self.solver_order = solver_order
        self.spacing = timestep_spacing
        self.alphas_cumprod = _alphas_cumprod(num_train_timesteps)

    def set_timesteps(self, n: int):
        self.num_inference_steps = n
        if self.spacing == "leading":
            ratio = self.num_train // (n + 1)
            ts = (np.arange(0, n + 1) * ratio).round()[::-1][:-1].copy().astype(np.int64) + self.steps_offset
        elif self.spacing == "trailing":
            ratio = self.num_train / n
            ts = (np.arange(self.num_train, 0, -ratio)).round().copy().astype(np.int64) - 1
        else:
            raise ValueError(self.spacing)
        sig = np.array(((1 - self.alphas_cumprod) / self.alphas_cumprod) ** 0.5)
        sig = np.interp(ts, np.arange(0, len(sig)), sig)
        self.sigmas = torch.from_numpy(np.concatenate([sig, [0.0]]).astype(np.float32))
        self.timesteps = torch.from_numpy(ts)
        self.model_outputs: List[Optional[torch.Tensor]] = [None] * self.solver_order
        self.lower_order_nums = 0
        self.last_sample = None
        self.step_index: Optional[int] = None
        self.begin_index: Optional[int] = None
        self.this_order = 1

    def set_begin_index(self, i: int):
        self.begin_index = i

    def scale_model_input(self, x, t):
        return x

    @staticmethod
    def _alpha_sigma(sigma):
        alpha_t = 1 / ((sigma ** 2 + 1) ** 0.5)
        return alpha_t, sigma * alpha_t

    def add_noise(self, x0, noise, t):
        idx = self.begin_index if self.begin_index is not None else int((self.timesteps == int(t)).nonzero()[0])
        a, s = self._alpha_sigma(self.sigmas[idx])
        return a * x0 + s * noise

    def _convert(self, eps, x):
        a, s = self._alpha_sigma(self.sigmas[self.step_index])
        return (x - s * eps) / a

    def _uni_p(self, x, order):
        m0 = self.model_outputs[-1]
        sigma_t, sigma_s0 = self.sigmas[self.step_index + 1], self.sigmas[self.step_index]
        alpha_t, sigma_t = self._alpha_sigma(sigma_t)
        alpha_s0, sigma_s0 = self._alpha_sigma(sigma_s0)
        lambda_t = torch.log(alpha_t) - torch.log(sigma_t)
        lambda_s0 = torch.log(alpha_s0) - torch.log(sigma_s0)
        h = lambda_t - lambda_s0
        rks, D1s = [], []
        for i in range(1, order):
            si = self.step_index - i
            mi = self.model_outputs[-(i + 1)]
            a_si, s_si = self._alpha_sigma(self.sigmas[si])
            lambda_si = torch.log(a_si) - torch.log(s_si)
            rk = (lambda_si - lambda_s0) / h
            rks.append(rk)
            D1s.append((mi - m0) / rk)
        rks.append(1.0)
        rks = torch.tensor(rks)
        R, b = [], []
        hh = -h
        h_phi_1 = torch.expm1(hh)
        h_phi_k = h_phi_1 / hh - 1
        factorial_i = 1
        B_h = torch.expm1(hh)  # bh2
        for i in range(1, order + 1):
            R.append(torch.pow(rks, i - 1))
            b.append(h_phi_k * factorial_i / B_h)
            factorial_i *= i + 1
            h_phi_k = h_phi_k / hh - 1 / factorial_i
        R, b = torch.stack(R), torch.tensor(b)
        if len(D1s) > 0:
            D1s = torch.stack(D1s, dim=1)
            if order == 2:
                rhos_p = torch.tensor([0.5], dtype=x.dtype)
            else:
                rhos_p = torch.linalg.solve(R[:-1, :-1], b[:-1]).to(x.dtype)
        else:
            D1s = None
        x_t_ = sigma_t / sigma_s0 * x - alpha_t * h_phi_1 * m0
        if D1s is not None:
            pred_res = torch.einsum("k,bkc...->bc...", rhos_p, D1s)
        else:
            pred_res = 0
        return (x_t_ - alpha_t * B_h * pred_res).to(x.dtype)

    def _uni_c(self, this_model_output, last_sample, this_sample, order):
        m0 = self.model_outputs[-1]
        x = last_sample
        model_t = this_model_output
        sigma_t, sigma_s0 = self.sigmas[self.step_index], self.sigmas[self.step_index - 1]
        alpha_t, sigma_t = self._alpha_sigma(sigma_t)
        alpha_s0, sigma_s0 = self._alpha_sigma(sigma_s0)
        lambda_t = torch.log(alpha_t) - torch.log(sigma_t)
        lambda_s0 = torch.log(alpha_s0) - torch.log(sigma_s0)
        h = lambda_t - lambda_s0
        rks, D1s = [], []
        for i in range(1, order):
            si = self.step_index - (i + 1)
            mi = self.model_outputs[-(i + 1)]
            a_si, s_si = self._alpha_sigma(self.sigmas[si])
            lambda_si = torch.log(a_si) - torch.log(s_si)
            rk = (lambda_si - lambda_s0) / h
            rks.append(rk)
            D1s.append((mi - m0) / rk)
        rks.append(1.0)
        rks = torch.tensor(rks)
        R, b = [], []
        hh = -h
        h_phi_1 = torch.expm1(hh)
        h_phi_k = h_phi_1 / hh - 1
        factorial_i = 1
        B_h = torch.expm1(hh)
        for i in range(1, order + 1):
            R.append(torch.pow(rks, i - 1))
            b.append(h_phi_k * factorial_i / B_h)
            factorial_i *= i + 1
            h_phi_k = h_phi_k / hh - 1 / factorial_i
        R, b = torch.stack(R), torch.tensor(b)
        D1s = torch.stack(D1s, dim=1) if len(D1s) > 0 else None
        if order == 1:
            rhos_c = torch.tensor([0.5], dtype=x.dtype)
        else:
            rhos_c = torch.linalg.solve(R, b).to(x.dtype)
        x_t_ = sigma_t / sigma_s0 * x - alpha_t * h_phi_1 * m0
        corr_res = torch.einsum("k,bkc...->bc...", rhos_c[:-1], D1s) if D1s is not None else 0
        D1_t = model_t - m0
        return (x_t_ - alpha_t * B_h * (corr_res + rhos_c[-1] * D1_t)).to(x.dtype)

    def step(self, eps, t, x):
        if self.step_index is None:
            self.step_index = self.begin_index if self.begin_index is not None else int((self.timesteps == int(t)).nonzero()[0])
        use_corrector = self.step_index > 0 and self.last_sample is not None
        m = self._convert(eps, x)
        if use_corrector:
            x = self._uni_c(m, self.last_sample, x, self.this_order)
        for i in range(self.solver_order - 1):
            self.model_outputs[i] = self.model_outputs[i + 1]
        self.model_outputs[-1] = m
        this_order = min(self.solver_order, len(self.timesteps) - self.step_index)  # lower_order_final
        self.this_order = min(this_order, self.lower_order_nums + 1)
        self.last_sample = x
        prev = self._uni_p(x, self.this_order)
        if self.lower_order_nums < self.solver_order:
            self.lower_order_nums += 1
        self.step_index += 1
        return prev
