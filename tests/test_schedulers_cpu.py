"""CPU (host logic): the fused-step coefficient tables reproduce the oracle schedulers' `step` (DDIM and DPM-Solver++ 2M)
on random data, including a caller-truncated timestep list like train.py:760 passes."""
import os
import sys

import numpy as np
import pytest
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))

from oracle.composition import DDIMScheduler as ODDIM, DPMSolverMultistepScheduler as ODPM  # noqa: E402
from animate_anything_b200 import schedulers as S  # noqa: E402

KW = dict(beta_start=0.00085, beta_end=0.012, beta_schedule="scaled_linear")


def _simulate(coef, needs_hist, x, eps_list):
    hist = np.zeros_like(x)
    for k, e in zip(coef.astype(np.float64), eps_list):
        x0 = k[0] * x + k[1] * e
        xn = k[2] * x + k[3] * e + k[4] * x0 + (k[5] * hist if needs_hist else 0.0)
        hist = x0
        x = xn
    return x


@pytest.mark.parametrize("steps,trunc", [(50, 0), (25, 0), (10, 4), (3, 0)])
def test_ddim_table(steps, trunc):
    o = ODDIM(clip_sample=False, set_alpha_to_one=False, steps_offset=1, **KW)
    m = S.DDIMScheduler(clip_sample=False, set_alpha_to_one=False, steps_offset=1, **KW)
    o.set_timesteps(steps)
    m.set_timesteps(steps)
    assert o.timesteps.tolist() == m.timesteps.tolist()
    ts = m.timesteps.tolist()[trunc:]
    g = torch.Generator().manual_seed(0)
    x = torch.randn(64, generator=g, dtype=torch.float64)
    eps = [torch.randn(64, generator=g, dtype=torch.float64) for _ in ts]
    xo = x.clone()
    for t, e in zip(ts, eps):
        xo = o.step(e, t, xo).prev_sample
    coef, hist = m.step_coefficients(ts)
    xm = _simulate(coef, hist, x.numpy(), [e.numpy() for e in eps])
    assert np.allclose(xm, xo.numpy(), rtol=2e-5, atol=2e-5), np.abs(xm - xo.numpy()).max()


@pytest.mark.parametrize("steps,trunc", [(25, 0), (50, 0), (10, 3), (4, 0)])
def test_dpm_table(steps, trunc):
    base = ODDIM(clip_sample=False, set_alpha_to_one=False, steps_offset=1, **KW)
    o = ODPM.from_config(base.config)
    mb = S.DDIMScheduler(clip_sample=False, set_alpha_to_one=False, steps_offset=1, **KW)
    m = S.DPMSolverMultistepScheduler.from_config(mb.config)
    o.set_timesteps(steps)
    m.set_timesteps(steps)
    assert o.timesteps.tolist() == m.timesteps.tolist()
    ts = m.timesteps.tolist()[trunc:]
    g = torch.Generator().manual_seed(0)
    x = torch.randn(64, generator=g, dtype=torch.float64)
    eps = [torch.randn(64, generator=g, dtype=torch.float64) for _ in ts]
    xo = x.clone()
    for t, e in zip(ts, eps):
        xo = o.step(e, torch.tensor(t), xo).prev_sample
    coef, hist = m.step_coefficients(ts)
    xm = _simulate(coef, hist, x.numpy(), [e.numpy() for e in eps])
    assert np.allclose(xm, xo.numpy(), rtol=1e-4, atol=1e-4), np.abs(xm - xo.numpy()).max()


def test_add_noise_matches():
    o = ODDIM(clip_sample=False, set_alpha_to_one=False, steps_offset=1, **KW)
    m = S.DDIMScheduler(clip_sample=False, set_alpha_to_one=False, steps_offset=1, **KW)
    x = torch.randn(2, 4, 3, 8, 8)
    n = torch.randn(2, 4, 3, 8, 8)
    t = torch.tensor([481, 481])
    assert torch.allclose(o.add_noise(x, n, t), m.add_noise(x, n, t))
