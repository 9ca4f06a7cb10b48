"""Per-role wait-cycle breakdown of the implicit-GEMM kernel for a few shapes (uses AabIgemmDesc.debug_cycles)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from animate_anything_b200 import ops  # noqa: E402

dt = torch.bfloat16
g = torch.Generator(device="cuda").manual_seed(0)
rnd = lambda *s: torch.randn(*s, device="cuda", generator=g).to(dt)
NAMES = ["prod_wait_empty", "mma_wait_full", "mma_wait_tempty", "-", "-", "epi0_wait_tfull", "epi0_load_math",
         "epi0_drain_bar_store"]
cases = [("139264x320x320+res", 139264, 320, 320, True, False), ("139264x320x320", 139264, 320, 320, False, False),
         ("139264x960x320", 139264, 960, 320, False, False), ("geglu 139264x2560x320", 139264, 2560, 320, False, True)]
for name, m, n, k, res, geglu in cases:
    x, w, b = rnd(m, k), rnd(n, k), torch.randn(n, device="cuda")
    r = rnd(m, n) if res else None
    fn = lambda: ops.linear(x, w, b, residual=r, geglu=geglu)
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        fn()
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / 5 * 1e3
    ops.IGEMM_DEBUG = torch.zeros(16, device="cuda", dtype=torch.int64)
    fn()
    torch.cuda.synchronize()
    c = ops.IGEMM_DEBUG.tolist()
    ops.IGEMM_DEBUG = None
    ctas = 148
    life = c[15] / ctas
    print(f"{name:26s} {us:8.1f} us  {2 * m * n * k / us / 1e6:7.0f} TFLOP/s  producer-lifetime {life:9.0f} clk/CTA")
    print("    " + "  ".join(f"{nm}={100 * c[i] / ctas / max(life, 1):.0f}%" for i, nm in enumerate(NAMES)))
