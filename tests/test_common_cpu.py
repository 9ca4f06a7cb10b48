"""Host logic of animate_anything_b200.common (utils/common.py mirror) against the outputs of the VERBATIM reference
functions (tests/golden/common_ref.pt).  No GPU here: `ops.add_noise` (the aab_add_noise kernel) is replaced by a torch
restatement of its arithmetic -- product rounded, product rounded, sum rounded -- so that what is pinned is everything
AROUND the kernel: kept timesteps, RNG call, 16-bit coefficient roundings, frame repeat, mask resize and blend.
(The kernel itself is pinned bit-exactly on the GPU by tests/test_gpu_pipeline.py.)"""
import os
import sys

import pytest
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))


def _add_noise_restated(x0, noise, sa, sb):
    f = noise.shape[2]
    x = x0.expand(-1, -1, f, -1, -1) if x0.shape[2] == 1 else x0
    dt = noise.dtype
    t1 = (x.float() * sa).to(dt)
    t2 = (noise.float() * sb).to(dt)
    return (t1.float() + t2.float()).to(dt)


@pytest.fixture()
def common(monkeypatch):
    from animate_anything_b200 import common as C, ops
    monkeypatch.setattr(ops, "add_noise", _add_noise_restated)
    return C


def _sched():
    from animate_anything_b200.schedulers import DDIMScheduler
    s = DDIMScheduler(beta_start=0.00085, beta_end=0.012, beta_schedule="scaled_linear", clip_sample=False,
                      set_alpha_to_one=False, steps_offset=1)
    s.set_timesteps(10)
    return s


def test_ddpm_forward_timesteps_matches_reference(common):
    gold = torch.load(os.path.join(HERE, "golden", "common_ref.pt"))
    for name, dt in (("f32", torch.float32), ("bf16", torch.bfloat16), ("f16", torch.float16)):
        torch.manual_seed(3)
        xt, ts = common.DDPM_forward_timesteps(gold["x0"].to(dt), 4, 5, _sched())
        assert torch.equal(torch.as_tensor(ts), torch.as_tensor(gold["timesteps"]))
        assert torch.equal(xt, gold[f"xt_{name}"]), f"{name}: {(xt.float() - gold[f'xt_{name}'].float()).abs().max()}"


def test_ddpm_forward_mask_and_forward_match_reference(common):
    gold = torch.load(os.path.join(HERE, "golden", "common_ref.pt"))
    np_mask = gold["np_mask"].numpy()
    for name, dt in (("f32", torch.float32), ("bf16", torch.bfloat16)):
        torch.manual_seed(4)
        xm, ts = common.DDPM_forward_mask(gold["x0"].to(dt), 4, 5, _sched(), np_mask)
        assert xm.dtype == dt and torch.equal(torch.as_tensor(ts), torch.as_tensor(gold["timesteps"]))
        assert torch.equal(xm, gold[f"xmask_{name}"]), f"mask {name}: {(xm.float() - gold[f'xmask_{name}'].float()).abs().max()}"
        torch.manual_seed(6)
        xf, none = common.DDPM_forward(gold["x0"].to(dt), 4, 5, _sched())
        assert none is None
        assert torch.equal(xf, gold[f"xfwd_{name}"]), f"fwd {name}: {(xf.float() - gold[f'xfwd_{name}'].float()).abs().max()}"


def test_tensor_to_vae_latent_refuses_foreign_vae(common):
    with pytest.raises(TypeError):
        common.tensor_to_vae_latent(torch.zeros(1, 1, 3, 8, 8), object())
