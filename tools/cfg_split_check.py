"""2-GPU check (torchrun --nproc-per-node 2): CFG halves split over two GPUs give bit-identical latents to the
single-GPU pipeline, and the per-clip latency.  Small model for the equality check, full-size for the timing."""
import os
import sys
import time

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import build_models, synth_inputs  # noqa: E402

rank = int(os.environ["RANK"])
local = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
dev = torch.device("cuda", local)
pipe = build_models(dev, torch.bfloat16)
inp = {k: v.to(dev) for k, v in synth_inputs(0, pinned=False).items()}      # same clip on both ranks
kw = dict(motion=[4], guidance_scale=9.0, num_inference_steps=int(os.environ.get("STEPS", "50")), output_type="pt",
          return_dict=False)


def run():
    torch.cuda.synchronize()
    dist.barrier()
    t0 = time.perf_counter()
    v, lat = pipe(prompt_embeds=inp["prompt_embeds"], negative_prompt_embeds=inp["negative_prompt_embeds"],
                  latents=inp["latents"], condition_latent=inp["condition_latent"], mask=inp["mask"], **kw)
    torch.cuda.synchronize()
    return lat, time.perf_counter() - t0


lat1, _ = run()
lat1, t_single = run()
pipe.cfg_group = dist.group.WORLD
lat2, _ = run()
lat2, t_split = run()
same = bool(torch.equal(lat1, lat2))
maxdiff = float((lat1.float() - lat2.float()).abs().max())
if rank == 0:
    print(f"cfg-split over 2 GPUs: identical={same} maxdiff={maxdiff:.3e}  single-GPU clip {t_single:.2f} s -> split {t_split:.2f} s "
          f"({t_single / t_split:.2f}x)", flush=True)
dist.destroy_process_group()
