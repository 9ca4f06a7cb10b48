"""One full-size UNet3D forward at config-2 shapes (B=2 CFG pair, T=17, 64x64 latents, bf16) for profiling.
  python tools/profile_unet.py events   -> per-launch CUDA-event timing of every implicit-GEMM launch (JSON to gpurun_out/)
  ncu --profile-from-start off ... python tools/profile_unet.py ncu   -> cudaProfilerStart/Stop around ONE forward
"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import build_models, synth_inputs  # noqa: E402
from animate_anything_b200 import ops  # noqa: E402

mode = sys.argv[1] if len(sys.argv) > 1 else "events"
dev = torch.device("cuda", 0)
pipe = build_models(dev, torch.bfloat16)
inp = {k: v.to(dev) for k, v in synth_inputs(0, pinned=False).items()}
sample = inp["latents"].expand(2, -1, -1, -1, -1)
cond2 = torch.cat([inp["condition_latent"]] * 2)
ehs = torch.cat([inp["negative_prompt_embeds"], inp["prompt_embeds"]])
tt = torch.tensor([500.0], device=dev)
mot = torch.tensor([4.0], device=dev)


def fwd():
    return pipe.unet(sample, tt, ehs, condition_latent=cond2, mask=inp["mask"], motion=mot, _raw_eps=True)


for _ in range(2):
    fwd()
torch.cuda.synchronize()
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
if mode == "ncu":
    torch.cuda.profiler.start()
    fwd()
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
elif mode == "kernels":
    # per-shape CUDA-event timing of the non-GEMM kernels of one forward (algorithmic bytes -> GB/s, flops -> TFLOP/s)
    ops.KERNEL_PROFILE = []
    fwd()
    torch.cuda.synchronize()
    kp = ops.KERNEL_PROFILE
    ops.KERNEL_PROFILE = None
    agg = {}
    for p in kp:
        key = (p["name"], p["shape"])
        a = agg.setdefault(key, {"n": 0, "ms": 0.0, "bytes": 0.0, "flops": 0.0})
        a["n"] += 1
        a["ms"] += p["ev"][0].elapsed_time(p["ev"][1])
        a["bytes"] += p["bytes"]
        a["flops"] += p["flops"]
    print("| kernel | shape | launches | ms total | us / launch | algorithmic GB/s | TFLOP/s |")
    print("|---|---|---|---|---|---|---|")
    tot = {}
    for (name, shape), a in sorted(agg.items(), key=lambda kv: -kv[1]["ms"]):
        tot[name] = tot.get(name, 0.0) + a["ms"]
        print(f"| {name} | {shape} | {a['n']} | {a['ms']:.3f} | {1e3 * a['ms'] / a['n']:.1f} | "
              f"{a['bytes'] / a['ms'] / 1e6:.0f} | {a['flops'] / a['ms'] / 1e9:.0f} |")
    print("totals (ms):", {k: round(v, 3) for k, v in tot.items()})
elif mode == "vae_time":
    lat = inp["latents"]
    for _ in range(2):
        pipe.vae.decode_video(lat)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        pipe.vae.decode_video(lat)
    e1.record()
    torch.cuda.synchronize()
    print(f"vae decode 16 frames 512x512: {e0.elapsed_time(e1) / 3:.2f} ms  ({40.2e12 / (e0.elapsed_time(e1) / 3) / 1e9:.0f} TFLOP/s)")
    img = torch.randn(1, 3, 512, 512, device=dev).to(torch.bfloat16)
    pipe.vae.encode(img)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(3):
        pipe.vae.encode(img)
    e1.record()
    torch.cuda.synchronize()
    print(f"vae encode 1 frame 512x512: {e0.elapsed_time(e1) / 3:.2f} ms")
elif mode == "vae":
    lat = inp["latents"]
    pipe.vae.decode_video(lat)
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    pipe.vae.decode_video(lat)
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
else:
    ops.IGEMM_PROFILE = []
    fwd()
    torch.cuda.synchronize()
    rows = []
    for p in ops.IGEMM_PROFILE:
        ms = p["ev"][0].elapsed_time(p["ev"][1])
        rows.append({k: v for k, v in p.items() if k != "ev"} | {"ms": ms, "tflops": p["flops"] / ms / 1e9})
    ops.IGEMM_PROFILE = None
    agg = {}
    for r in rows:
        key = (r["rows"], r["n"], r["k"], r["taps"], r["block_n"], r.get("pair", False))
        a = agg.setdefault(key, {"count": 0, "ms": 0.0, "flops": 0.0})
        a["count"] += 1
        a["ms"] += r["ms"]
        a["flops"] += r["flops"]
    table = sorted(({"rows": k[0], "n": k[1], "k": k[2], "taps": k[3], "block_n": k[4], "pair": k[5], **v,
                     "tflops": v["flops"] / v["ms"] / 1e9} for k, v in agg.items()), key=lambda d: -d["ms"])
    tot = sum(t["ms"] for t in table)
    print(f"igemm total {tot:.2f} ms over {len(rows)} launches")
    for t in table[:40]:
        print(f"rows={t['rows']:7d} n={t['n']:5d} k={t['k']:6d} taps={t['taps']} bn={t['block_n']:3d}{'p' if t['pair'] else ' '} x{t['count']:3d} "
              f"{t['ms']:8.3f} ms  {t['tflops']:7.1f} TFLOP/s")
    json.dump(table, open(os.path.join(ROOT, "gpurun_out", "igemm_shapes%s.json" % os.environ.get("AAB_PROFILE_TAG", "")), "w"), indent=1)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        fwd()
    e1.record()
    torch.cuda.synchronize()
    print(f"unet forward {e0.elapsed_time(e1) / 3:.2f} ms")
