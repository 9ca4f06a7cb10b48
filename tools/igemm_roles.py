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
for name, m, n, k, res, geglu in ([] if "small" in sys.argv[1:] else cases):
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
    ops.IGEMM_DEBUG = torch.zeros(320, device="cuda", dtype=torch.int64)
    fn()
    torch.cuda.synchronize()
    c = ops.IGEMM_DEBUG.tolist()
    ops.IGEMM_DEBUG = None
    ctas = 148
    life = c[15] / ctas
    print(f"{name:26s} {us:8.1f} us  {2 * m * n * k / us / 1e6:7.0f} TFLOP/s  producer-lifetime {life:9.0f} clk/CTA")
    print("    " + "  ".join(f"{nm}={100 * c[i] / ctas / max(life, 1):.0f}%" for i, nm in enumerate(NAMES)))

if "small" in sys.argv[1:]:
    # low-resolution levels (8x8 and 16x16 latents): few M tiles -> wave quantisation and L2->SM bandwidth
    def run(name, fn, flops):
        fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            fn()
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) / 5 * 1e3
        ops.IGEMM_DEBUG = torch.zeros(320, device="cuda", dtype=torch.int64)
        fn()
        torch.cuda.synchronize()
        c = ops.IGEMM_DEBUG.tolist()
        ops.IGEMM_DEBUG = None
        life = max(c[15] / 148, 1)
        print(f"{name:34s} {us:8.1f} us {flops / us / 1e6:6.0f} TFLOP/s  life {life:8.0f} clk  " +
              " ".join(f"{nm}={100 * c[i] / 148 / life:.0f}%" for i, nm in enumerate(NAMES) if nm != "-"), flush=True)

    if "trace_smallk" in sys.argv[1:]:
        # needs a library built with -DAAB_IGEMM_TRACE (AAB_LIB_PATH).  CTA 0 timeline for the epilogue-bound K=320 shape.
        m, n, k = 139264, 320, 320
        x, w, b, r = rnd(m, k), rnd(n, k), torch.randn(n, device="cuda"), rnd(m, n)
        for bn, res in ((128, True), (128, False)):
            fn = lambda: ops.linear(x, w, b, residual=(r if res else None), block_n=bn)
            fn()
            ops.IGEMM_DEBUG = torch.zeros(1024, device="cuda", dtype=torch.int64)
            fn()
            torch.cuda.synchronize()
            c_ = ops.IGEMM_DEBUG.tolist()
            ops.IGEMM_DEBUG = None
            t0 = c_[16]
            kpt = 5
            ntile = 96 // kpt
            print(f"== linear {m}x{n}x{k} res={res} bn{bn}: tile | first stage empty_ready | first full_ready | last issued | "
                  f"epi g0 start/end | g1 | g2 | g3   (clk rel. to start)")
            for tl in range(min(ntile, 19)):
                it0, it1 = tl * kpt, tl * kpt + kpt - 1
                e = c_[16 + it0] - t0
                f_ = c_[16 + 96 + it0] - t0
                i_ = c_[16 + 192 + it1] - t0
                eg = [(c_[16 + 288 + g_ * 64 + 2 * tl] - t0, c_[16 + 288 + g_ * 64 + 2 * tl + 1] - t0) for g_ in range(4)]
                print(f"   {tl:3d} {e:8d} {f_:8d} {i_:8d}   " + "  ".join(f"{a_:7d}/{b_:7d}" for a_, b_ in eg))
        sys.exit(0)
    if "trace" in sys.argv[1:]:
        # per-k-block timestamps of CTA 0 (clock64): producer saw the stage empty / MMA warp saw it full / MMAs issued
        hw, c = 16, 1280
        x = rnd(34, hw, hw, c)
        w9 = rnd(c, 9 * c)
        b = torch.randn(c, device="cuda")
        for bn in (256, 128):
            for nm, fl in (("full", 0), ("no_load", 128)):
                ops.IGEMM_DBG_FLAGS = fl
                fn = lambda: ops.conv3x3(x, w9, b, block_n=bn)
                fn()
                ops.IGEMM_DEBUG = torch.zeros(320, device="cuda", dtype=torch.int64)
                fn()
                torch.cuda.synchronize()
                c_ = ops.IGEMM_DEBUG.tolist()
                ops.IGEMM_DEBUG = None
                t0 = c_[16]
                print(f"== conv3x3 8704x1280x11520 bn{bn} {nm}: it  empty_ready  full_ready  issued   (d_issued)")
                prev = None
                for it in range(96):
                    e, f, i_ = c_[16 + it] - t0, c_[16 + 96 + it] - t0, c_[16 + 192 + it] - t0
                    print(f"   {it:3d} {e:10d} {f:10d} {i_:10d}   {'' if prev is None else i_ - prev}")
                    prev = i_
        ops.IGEMM_DBG_FLAGS = 0
        sys.exit(0)
    if "diag" in sys.argv[1:]:
        for hw, c in ((16, 1280), (64, 320)):
            x = rnd(34, hw, hw, c)
            w9, w1 = rnd(c, 9 * c), rnd(c, c)
            b = torch.randn(c, device="cuda")
            rows = 34 * hw * hw
            for bn in (128, 256):
                for nm, fl in (("full", 0), ("no_load", 128), ("no_load_no_sync", 128 + 256)):
                    ops.IGEMM_DBG_FLAGS = fl
                    run(f"conv3x3 {rows}x{c}x{9 * c} bn{bn} {nm}", lambda: ops.conv3x3(x, w9, b, block_n=bn), 2 * rows * c * 9 * c)
                    run(f"linear  {rows}x{c}x{c} bn{bn} {nm}", lambda: ops.linear(x.view(rows, c), w1, b, block_n=bn), 2 * rows * c * c)
            ops.IGEMM_DBG_FLAGS = 0
        sys.exit(0)
    for hw, c in ((8, 1280), (16, 1280), (32, 640)):
        x = rnd(34, hw, hw, c)
        w9, w3, w1 = rnd(c, 9 * c), rnd(c, 3 * c), rnd(c, c)
        b = torch.randn(c, device="cuda")
        rows = 34 * hw * hw
        for bn in (64, 128, 256):
            run(f"conv3x3 {rows}x{c}x{9 * c} bn{bn}", lambda: ops.conv3x3(x, w9, b, block_n=bn), 2 * rows * c * 9 * c)
            run(f"tconv3  {rows}x{c}x{3 * c} bn{bn}", lambda: ops.tconv3(x.view(rows, c), 2, 17, hw * hw, w3, b, block_n=bn),
                2 * rows * c * 3 * c)
            run(f"linear  {rows}x{c}x{c} bn{bn}", lambda: ops.linear(x.view(rows, c), w1, b, block_n=bn), 2 * rows * c * c)
