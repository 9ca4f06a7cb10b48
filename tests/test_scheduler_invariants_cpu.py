"""CPU: first-principles checks of the scheduler arithmetic — an independent pin of leaf semantics that are otherwise only
*recalled* (the diffusers shim is parity-unpinned), and of the coefficient tables the fused CFG + step kernel consumes.

Invariant used (it does not depend on any implementation): if the network returns the TRUE noise (DDIM / DPM-Solver++, epsilon
prediction) or the TRUE v-target (Euler, v-prediction) of a sample built from a known clean x0 and a known noise eps, one
sampler step must land exactly on the same (x0, eps) trajectory at the next noise level:

    variance preserving (DDIM, DPM-Solver++):   x_t = sqrt(abar_t) x0 + sqrt(1 - abar_t) eps
    variance exploding  (Euler / EDM sigmas):   x_sigma = x0 + sigma eps

for every step, for every order of the multistep solver (a constant x0 prediction makes the second-order correction vanish), and
for caller-truncated timestep lists."""
import os
import sys

import numpy as np
import pytest
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))

from oracle.composition import SVD_SCHED, DDIMScheduler as ODDIM, DPMSolverMultistepScheduler as ODPM  # noqa: E402
from oracle.composition import EulerDiscreteScheduler as OEuler  # noqa: E402
from animate_anything_b200 import schedulers as S  # noqa: E402

KW = dict(beta_start=0.00085, beta_end=0.012, beta_schedule="scaled_linear", clip_sample=False, set_alpha_to_one=False,
          steps_offset=1)


def _abar():
    betas = torch.linspace(0.00085 ** 0.5, 0.012 ** 0.5, 1000, dtype=torch.float64) ** 2          # scaled_linear, from the SD config
    return torch.cumprod(1.0 - betas, dim=0)


def _on_trajectory(x0, eps, abar_t):
    return abar_t ** 0.5 * x0 + (1 - abar_t) ** 0.5 * eps


@pytest.mark.parametrize("steps", [50, 25, 7])
def test_ddim_true_noise_stays_on_trajectory(steps):
    abar = _abar()
    o = ODDIM(**KW)
    o.set_timesteps(steps)
    ts = o.timesteps.tolist()
    assert ts == [(steps - 1 - i) * (1000 // steps) + 1 for i in range(steps)]                   # "leading" spacing, steps_offset 1
    g = torch.Generator().manual_seed(0)
    x0, eps = torch.randn(257, generator=g, dtype=torch.float64), torch.randn(257, generator=g, dtype=torch.float64)
    x = _on_trajectory(x0, eps, abar[ts[0]])
    for i, t in enumerate(ts):
        x = o.step(eps, t, x).prev_sample
        prev = t - 1000 // steps
        target = abar[prev] if prev >= 0 else abar[0]                                             # set_alpha_to_one=False -> abar_0
        assert torch.allclose(x, _on_trajectory(x0, eps, target), rtol=5e-5, atol=5e-5), (i, t)   # the scheduler keeps float32 betas


def _linear_response(coef_row, hist_row=None):
    """The fused step is linear: x' = k2 x + k3 e + k4 x0_pred + k5 x0_prev with x0_pred = k0 x + k1 e."""
    k = coef_row.astype(np.float64)
    return k


@pytest.mark.parametrize("cls,steps,trunc", [("ddim", 50, 0), ("ddim", 10, 4), ("dpm", 25, 0), ("dpm", 10, 3), ("dpm", 4, 0)])
def test_product_coefficient_tables_keep_true_noise_on_trajectory(cls, steps, trunc):
    """The tables `step_coefficients` hands to `aab_cfg_scheduler_step`, simulated in fp64 with the kernel's formula."""
    abar = _abar().numpy()
    base = S.DDIMScheduler(**KW)
    m = base if cls == "ddim" else S.DPMSolverMultistepScheduler.from_config(base.config)
    m.set_timesteps(steps)
    ts = m.timesteps.tolist()[trunc:]
    coef, needs_hist = m.step_coefficients(ts)
    rng = np.random.default_rng(1)
    x0, eps = rng.standard_normal(129), rng.standard_normal(129)
    x = np.sqrt(abar[ts[0]]) * x0 + np.sqrt(1 - abar[ts[0]]) * eps
    hist = np.zeros_like(x)
    for i, (t, k) in enumerate(zip(ts, coef.astype(np.float64))):
        x0p = k[0] * x + k[1] * eps
        assert np.allclose(x0p, x0, rtol=1e-4, atol=1e-4) or (k[0] == 0 and k[1] == 0)           # DDIM rows leave k0, k1 unused
        x = k[2] * x + k[3] * eps + k[4] * x0p + (k[5] * hist if needs_hist else 0.0)
        hist = x0p
        nxt = ts[i + 1] if i + 1 < len(ts) else None
        if nxt is not None:
            want = np.sqrt(abar[nxt]) * x0 + np.sqrt(1 - abar[nxt]) * eps
            assert np.allclose(x, want, rtol=2e-4, atol=2e-4), (cls, i, t, np.abs(x - want).max())
    # the last step ends at abar_0 (DDIM: final_alpha_cumprod with set_alpha_to_one=False; DPM-Solver++ 0.24: sigma_last from abar_0)
    want = np.sqrt(abar[0]) * x0 + np.sqrt(1 - abar[0]) * eps
    assert np.allclose(x, want, rtol=2e-4, atol=2e-4), np.abs(x - want).max()


@pytest.mark.parametrize("steps,trunc", [(25, 0), (10, 3), (3, 0)])
def test_oracle_dpm_solver_true_noise_stays_on_trajectory(steps, trunc):
    abar = _abar()
    o = ODPM.from_config(ODDIM(**KW).config)
    o.set_timesteps(steps)
    ts = o.timesteps.tolist()[trunc:]
    g = torch.Generator().manual_seed(2)
    x0, eps = torch.randn(65, generator=g, dtype=torch.float64), torch.randn(65, generator=g, dtype=torch.float64)
    x = _on_trajectory(x0, eps, abar[ts[0]])
    for i, t in enumerate(ts):
        x = o.step(eps, torch.tensor(t), x).prev_sample
        nxt = ts[i + 1] if i + 1 < len(ts) else 0
        assert torch.allclose(x, _on_trajectory(x0, eps, abar[nxt]), rtol=1e-4, atol=1e-4), (i, t)


@pytest.mark.parametrize("steps", [25, 5])
def test_oracle_euler_v_prediction_true_target_stays_on_trajectory(steps):
    """EDM parameterisation of the SVD scheduler: x = x0 + sigma eps, network input x / sqrt(sigma^2 + 1), v-prediction
    x0 = -sigma / sqrt(sigma^2 + 1) * v + x / (sigma^2 + 1)  =>  the exact v for a known x0 is v* = (x / (sigma^2 + 1) - x0) * sqrt(sigma^2 + 1) / sigma."""
    o = OEuler(**SVD_SCHED)
    o.set_timesteps(steps)
    sig = [float(s) for s in o.sigmas]
    assert len(sig) == steps + 1 and sig[-1] == 0.0 and all(a > b for a, b in zip(sig, sig[1:]))
    assert abs(sig[0] - 700.0) < 1e-3 and abs(sig[-2] - 0.002) < 1e-6                              # Karras sigmas sigma_max .. sigma_min
    assert abs(float(o.init_noise_sigma) - (700.0 ** 2 + 1) ** 0.5) < 1e-2
    g = torch.Generator().manual_seed(3)
    x0, eps = torch.randn(33, generator=g, dtype=torch.float64), torch.randn(33, generator=g, dtype=torch.float64)
    x = x0 + sig[0] * eps
    for i, t in enumerate(o.timesteps):
        s = sig[i]
        scaled = o.scale_model_input(x, t)
        assert torch.allclose(scaled, x / (s * s + 1) ** 0.5, rtol=1e-6, atol=1e-9)
        v = (x / (s * s + 1) - x0) * (s * s + 1) ** 0.5 / s
        x = o.step(v, t, x).prev_sample
        assert torch.allclose(x, x0 + sig[i + 1] * eps, rtol=1e-5, atol=1e-5), (i, s)
    assert torch.allclose(x, x0, rtol=1e-5, atol=1e-5)                                             # sigma = 0 at the end


def test_product_euler_tables_equal_oracle_and_are_karras():
    m = S.EulerDiscreteScheduler(**SVD_SCHED)
    o = OEuler(**SVD_SCHED)
    for steps in (25, 4):
        m.set_timesteps(steps)
        o.set_timesteps(steps)
        sm, so = np.array([float(s) for s in m.sigmas]), np.array([float(s) for s in o.sigmas])
        assert np.allclose(sm, so, rtol=1e-6, atol=1e-9)
        rho = 7.0
        ramp = np.linspace(0, 1, steps)
        karras = (700.0 ** (1 / rho) + ramp * (0.002 ** (1 / rho) - 700.0 ** (1 / rho))) ** rho       # Karras et al. 2022, eq. 5
        assert np.allclose(sm[:-1], karras, rtol=1e-5)
        assert np.allclose(np.array([float(t) for t in m.timesteps]), 0.25 * np.log(sm[:-1]), rtol=1e-5, atol=1e-6)   # c_noise = ln(sigma) / 4
