"""Tile-width / CTA-pair sweep of the implicit GEMM at the launch shapes that sit furthest below the roofline
(graph-timed like tools/kernel_ab.py).   python tools/igemm_bn_sweep.py"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
from animate_anything_b200 import ops  # noqa: E402

dt = torch.bfloat16
dev = "cuda"


def timeit(fn, iters=16):
    for _ in range(2):
        fn(0)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        fn(0)
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    with torch.cuda.graph(g):
        for i in range(iters):
            fn(i)
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / (3 * iters) * 1e3


print("| rows | N | K | epilogue | variant | us | TFLOP/s |")
print("|---|---|---|---|---|---|---|")
cases = [(139264, 320, 320, "res"), (139264, 320, 320, "plain"), (34816, 640, 640, "res"), (139264, 960, 320, "plain"),
         (139264, 320, 1280, "res"), (8704, 1280, 1280, "res"), (139264, 2560, 320, "geglu")]
for m, n, k, epi in cases:
    nbuf = 3
    xs = [torch.randn(m, k, device=dev).to(dt) for _ in range(nbuf)]
    w = (torch.randn(n, k, device=dev) * k ** -0.5).to(dt)
    b = torch.randn(n, device=dev)
    res = torch.randn(m, n if epi != "geglu" else n // 2, device=dev).to(dt)
    variants = [("bn64", 64, False), ("bn128", 128, False), ("bn256", 256, False), ("bn256 pair", 256, True)]
    if epi == "geglu":
        variants = [("bn128", 128, False), ("bn256", 256, False), ("bn256 pair", 256, True)]
    for name, bn, pair in variants:
        ops.IGEMM_PAIR = "all" if pair else False
        kw = dict(block_n=bn)
        if epi == "res":
            kw["residual"] = res
        if epi == "geglu":
            kw["geglu"] = True
        try:
            us = timeit(lambda i: ops.linear(xs[i % nbuf], w, b, **kw))
        except Exception as ex:
            print(f"| {m} | {n} | {k} | {epi} | {name} | failed: {ex} | |")
            continue
        print(f"| {m} | {n} | {k} | {epi} | {name} | {us:.1f} | {2.0 * m * n * k / us / 1e6:.0f} |")
    del xs
ops.IGEMM_PAIR = True
