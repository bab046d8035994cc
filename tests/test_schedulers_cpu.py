"""CPU: the product's coefficient-table schedulers (host logic feeding saspa_cfg_sched_step) reproduce the
oracle restatement of diffusers' DDIM / UniPC / PNDM step() on random eps sequences (fp64-vs-fp32 scalar
algebra => tolerance 2e-5 relative to max|x|)."""
import numpy as np
import pytest
import torch

from oracle.diffusers_restated import schedulers as osched
from saspa_aug_b200 import schedulers as psched


def _apply(plan, bufs, eps):
    ins = [eps if nm is None else bufs[nm] for nm in plan.inputs]
    outs = []
    for row in plan.coef:
        acc = torch.zeros_like(eps)
        for c, t in zip(row, ins):
            acc = acc + c * t
        outs.append(acc)
    if plan.clip_range > 0:
        c = torch.zeros_like(eps)
        for w, t in zip(plan.clip_pre, ins):
            c = c + w * t
        c = c.clamp(-plan.clip_range, plan.clip_range)
        outs = [o + w * c for o, w in zip(outs, plan.clip_post)]
    for nm, o in zip(plan.outputs, outs):
        bufs[nm] = o


def _run(name, n, strength=None, seed=0):
    g = torch.Generator().manual_seed(seed)
    shape = (2, 4, 8, 8)
    from oracle.diffusers_restated.pipelines import make_scheduler as omake

    o = omake(name)
    p = psched.make_scheduler(name)
    o.set_timesteps(n)
    p.set_timesteps(n)
    assert np.array_equal(o.timesteps.numpy(), p.timesteps)
    start = 0
    ts = o.timesteps
    x = torch.randn(shape, generator=g)
    if strength is not None:
        start = p.img2img_start(n, strength)
        ts = ts[start:]
        if hasattr(o, "set_begin_index"):
            o.set_begin_index(start)
        z0, noise = torch.randn(shape, generator=g), torch.randn(shape, generator=g)
        xo = o.add_noise(z0, noise, ts[0])
        a, s = p.add_noise_coef(start)
        assert torch.allclose(xo, a * z0 + s * noise, atol=1e-5)
        x = xo
    bufs = {"x": x.clone()}
    for nm in p.history_buffers:
        bufs[nm] = torch.zeros(shape)
    p.begin(start)
    xo = x.clone()
    for k, t in enumerate(ts):
        eps = torch.randn(shape, generator=g)
        xo = o.step(eps, t, xo)
        _apply(p.plan(start + k), bufs, eps)
        err = (bufs["x"] - xo).abs().max().item() / max(xo.abs().max().item(), 1.0)
        assert err < 2e-5, (name, n, strength, k, err)
    return len(ts)


@pytest.mark.parametrize("name", ["ddim", "unipc", "pndm"])
@pytest.mark.parametrize("n", [2, 5, 20, 30])
def test_full_schedule(name, n):
    evals = _run(name, n)
    assert evals == (n + 1 if name == "pndm" else n)


@pytest.mark.parametrize("name", ["ddim_sdxl_turbo", "unipc_sdxl_turbo"])
@pytest.mark.parametrize("n,strength", [(1, None), (2, None), (4, None), (4, 0.5), (10, None)])
def test_sdxl_turbo_schedules(name, n, strength):
    """Trailing spacing; DDIM with diffusers' own defaults clip_sample=True / set_alpha_to_one=True (the clamp is active: x0 of
    random inputs leaves [-1,1])."""
    evals = _run(name, n, strength)
    assert evals == (n if strength is None else min(int(n * strength), n))


@pytest.mark.parametrize("name", ["ddim", "unipc"])
@pytest.mark.parametrize("n,strength", [(20, 0.5), (30, 0.85), (50, 0.15), (20, 1.0)])
def test_img2img_schedule(name, n, strength):
    evals = _run(name, n, strength)
    assert evals == min(int(n * strength), n)


def test_known_timesteps():
    p = psched.make_scheduler("unipc")
    p.set_timesteps(20)
    assert p.timesteps[0] == 941 and p.timesteps[1] == 894 and p.timesteps[-1] == 48
    d = psched.make_scheduler("ddim")
    d.set_timesteps(30)
    assert d.timesteps[0] == 958 and d.timesteps[-1] == 1
    t = psched.make_scheduler("ddim", timestep_spacing="trailing")
    t.set_timesteps(4)
    assert t.timesteps.tolist() == [999, 749, 499, 249]


@pytest.mark.parametrize("name,n", [("ddim", 20), ("ddim", 7), ("unipc", 20), ("unipc", 9), ("pndm", 20), ("pndm", 6)])
def test_exact_on_noise_consistent_trajectories(name, n):
    """Pins the restated schedulers (diffusers is not installable offline: "parity unpinned" at the reference boundary) to the one
    property all three solvers must have by construction: if the model returns the TRUE noise of a trajectory x_t = sqrt(a_t) x0 +
    sqrt(1 - a_t) eps (constant x0 and eps), every step lands exactly on that trajectory at the previous timestep -- DDIM (eta 0) by its
    closed form, UniPC (data prediction: the predicted x0 is constant, so predictor and corrector are exact at any order), PLMS (its
    transfer formula is exact for a constant eps).  Checked for the oracle and the product's coefficient tables, in fp64."""
    from oracle.diffusers_restated.pipelines import make_scheduler as omake

    g = torch.Generator().manual_seed(3)
    shape = (1, 4, 4, 4)
    x0 = torch.randn(shape, generator=g, dtype=torch.float64)
    eps = torch.randn(shape, generator=g, dtype=torch.float64)
    o, p = omake(name), psched.make_scheduler(name)
    o.set_timesteps(n)
    p.set_timesteps(n)
    betas = torch.linspace(0.00085 ** 0.5, 0.012 ** 0.5, 1000, dtype=torch.float64) ** 2  # scaled_linear
    acp = torch.cumprod(1.0 - betas, 0)

    def on_traj(t):
        a = acp[int(t)] if t >= 0 else torch.tensor(1.0, dtype=torch.float64)
        return a.sqrt() * x0 + (1 - a).sqrt() * eps

    ts = [int(t) for t in o.timesteps]
    x = on_traj(ts[0])
    bufs = {"x": x.clone()}
    for nm in p.history_buffers:
        bufs[nm] = torch.zeros(shape, dtype=torch.float64)
    p.begin(0)
    xo = x.clone()
    for k, t in enumerate(ts):
        xo = o.step(eps, torch.tensor(t), xo)
        _apply(p.plan(k), bufs, eps)
    # where the trajectory ends: DDIM / PLMS step to t_last - 1000/n (alpha of "timestep -1" = final_alpha_cumprod = a_0 with
    # set_alpha_to_one=False); UniPC's last step goes to sigma = 0, i.e. returns x0 itself
    if name == "unipc":
        want = x0
    else:
        want = acp[0].sqrt() * x0 + (1 - acp[0]).sqrt() * eps
    for got, tag in ((xo.double(), "oracle"), (bufs["x"], "product")):
        err = (got - want).abs().max().item()
        assert err < 2e-5, (name, n, tag, err)  # measured 0.6e-6 .. 2.6e-6 (fp32 coefficient tables)
