"""GPU parity of the attention kernels vs torch fp32 softmax attention on the same 16-bit inputs."""
import math

import pytest
import torch

from util import check

pytestmark = pytest.mark.gpu
TOL = {torch.float16: (2e-3, 2e-3), torch.bfloat16: (1.6e-2, 1.6e-2)}


def _rand(shape, dtype, scale=1.0, seed=0):
    g = torch.Generator(device="cuda").manual_seed(seed)
    return (torch.randn(shape, device="cuda", generator=g) * scale).to(dtype)


def _ref_attn(q, k, v, scale):
    s = torch.einsum("bhqd,bhkd->bhqk", q.float(), k.float()) * scale
    p = torch.softmax(s, dim=-1)
    return torch.einsum("bhqk,bhkd->bhqd", p, v.float())


@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16])
@pytest.mark.parametrize("nb,l,heads", [(2, 256, 2), (3, 300, 5), (1, 1024, 3), (2, 64, 4), (1, 4096, 1)])
def test_flash_self(dtype, nb, l, heads):
    from animate_anything_b200 import ops
    c = heads * 64
    qkv = _rand((nb * l, 3 * c), dtype, 1.0, 1)
    out = ops.flash_attn_d64(qkv, 0, qkv, c, 2 * c, nb, l, l, heads)
    torch.cuda.synchronize()
    t = qkv.reshape(nb, l, 3, heads, 64).permute(2, 0, 3, 1, 4)
    ref = _ref_attn(t[0], t[1], t[2], 1 / 8.0).permute(0, 2, 1, 3).reshape(nb * l, c)
    check(f"flash self nb{nb} l{l} h{heads} {dtype}", out, ref, *TOL[dtype])


@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16])
def test_flash_cross_77(dtype):
    """text cross-attention: 77 keys (masked tail of the 128-key tile), K/V shared by the T frames of a batch item."""
    from animate_anything_b200 import ops
    b, t, l, heads, lk = 2, 3, 256, 5, 77
    c = heads * 64
    q = _rand((b * t * l, c), dtype, 1.0, 1)
    kv = _rand((b * lk, 2 * c), dtype, 1.0, 2)
    out = ops.flash_attn_d64(q, 0, kv, 0, c, b * t, l, lk, heads, kv_batch_div=t)
    torch.cuda.synchronize()
    qq = q.reshape(b, t, l, heads, 64).permute(0, 1, 3, 2, 4).reshape(b * t, heads, l, 64)
    kk = kv[:, :c].reshape(b, 1, lk, heads, 64).expand(b, t, lk, heads, 64).permute(0, 1, 3, 2, 4).reshape(b * t, heads, lk, 64)
    vv = kv[:, c:].reshape(b, 1, lk, heads, 64).expand(b, t, lk, heads, 64).permute(0, 1, 3, 2, 4).reshape(b * t, heads, lk, 64)
    ref = _ref_attn(qq, kk, vv, 1 / 8.0).permute(0, 2, 1, 3).reshape(b * t * l, c)
    check(f"flash cross {dtype}", out, ref, *TOL[dtype])


@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16])
@pytest.mark.parametrize("b,t,hw,heads", [(2, 17, 64, 5), (1, 9, 256, 8), (2, 3, 16, 2)])
def test_temporal(dtype, b, t, hw, heads):
    from animate_anything_b200 import ops
    c = heads * 64
    qkv = _rand((b * t * hw, 3 * c), dtype, 1.0, 1)
    out = ops.temporal_attn_d64(qkv, b, t, hw, heads, 0, c, 2 * c)
    torch.cuda.synchronize()
    x = qkv.reshape(b, t, hw, 3, heads, 64).permute(3, 0, 2, 4, 1, 5).reshape(3, b * hw, heads, t, 64)
    ref = _ref_attn(x[0], x[1], x[2], 1 / 8.0)                       # [b*hw, heads, t, 64]
    ref = ref.reshape(b, hw, heads, t, 64).permute(0, 3, 1, 2, 4).reshape(b * t * hw, c)
    check(f"temporal attn b{b} t{t} hw{hw} h{heads} {dtype}", out, ref, *TOL[dtype])
