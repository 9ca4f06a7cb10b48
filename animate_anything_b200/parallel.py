"""Multi-GPU partitioning of the sampling path (one process per GPU, torch.distributed; NCCL on GPUs, gloo in CPU tests).

What shards: batch elements of the UNet (prompts / CFG halves) and the frames of the VAE; the
diffusion timesteps of one clip do NOT (latents at step i+1 depend on step i, models/pipeline.py:163-192).

  * prompt sharding (weak scaling, config 3): rank r owns prompts  r, r+W, ...  with both CFG halves co-located, so the
    denoising loop needs no communication at all; ONE all-gather of the decoded uint8 frames at the end.
  * frame-sharded VAE decode: rank r decodes a contiguous frame range of the final latents, one all-gather of frames.
  * cfg split (single clip on 2 GPUs): rank 0 runs the unconditional, rank 1 the text half; the fp32 noise prediction
    (1.1 MB at 16x64x64) is all-gathered every step and both ranks apply the same fused CFG + scheduler step.
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist


def shard_indices(n_items: int, rank: int, world: int) -> List[int]:
    """Round-robin ownership: item i belongs to rank i % world."""
    return [i for i in range(n_items) if i % world == rank]


def frame_range(n_frames: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous frame range [lo, hi) of rank `rank`; the first (n_frames % world) ranks get one extra frame."""
    base, extra = divmod(n_frames, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def to_uint8_frames(video: torch.Tensor) -> torch.Tensor:
    """[-1, 1] float video [b, 3, f, H, W] -> uint8 (same mapping as diffusers tensor2vid: (x * 0.5 + 0.5) * 255)."""
    return video.mul(127.5).add_(127.5).clamp_(0, 255).to(torch.uint8)


def all_gather_clips(frames_local: torch.Tensor, group=None) -> torch.Tensor:
    """frames_local [b_local, ...] (same shape on every rank) -> [world * b_local, ...] ordered by rank: the one
    collective of the prompt-sharded path."""
    world = dist.get_world_size(group)
    out = torch.empty((world,) + tuple(frames_local.shape), dtype=frames_local.dtype, device=frames_local.device)
    if frames_local.is_cuda:
        dist.all_gather_into_tensor(out, frames_local.contiguous().unsqueeze(0), group=group)
    else:
        parts = [torch.empty_like(frames_local) for _ in range(world)]
        dist.all_gather(parts, frames_local.contiguous(), group=group)
        out = torch.stack(parts)
    return out.reshape((-1,) + tuple(frames_local.shape[1:]))


def gather_round_robin(items_local: torch.Tensor, n_items: int, group=None) -> torch.Tensor:
    """Inverse of `shard_indices` when every rank owns the same count: returns the items in global order."""
    world = dist.get_world_size(group)
    g = all_gather_clips(items_local, group)                     # [world * per, ...] rank-major
    per = items_local.shape[0]
    order = [r * per + j for j in range(per) for r in range(world)]
    return g[order][:n_items]


def decode_video_frame_sharded(vae, latents: torch.Tensor, group=None) -> torch.Tensor:
    """Each rank decodes its contiguous frame range of latents [b, 4, f, h, w]; returns the full uint8 video
    [b, 3, f, 8h, 8w] on every rank (one all-gather; ranks are padded to the same frame count)."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    f = latents.shape[2]
    lo, hi = frame_range(f, rank, world)
    per = -(-f // world)
    local = to_uint8_frames(vae.decode_video(latents[:, :, lo:hi].contiguous()))
    if local.shape[2] < per:
        pad = torch.zeros(local.shape[:2] + (per - local.shape[2],) + local.shape[3:], dtype=local.dtype,
                          device=local.device)
        local = torch.cat([local, pad], dim=2)
    g = all_gather_clips(local.unsqueeze(0), group)              # [world, b, 3, per, H, W]
    pieces = []
    for r in range(world):
        l2, h2 = frame_range(f, r, world)
        pieces.append(g[r][:, :, : h2 - l2])
    return torch.cat(pieces, dim=2)


class PromptShardedPipeline:
    """Weak-scaling driver: every rank runs `LatentToVideoPipeline` on its own prompts (CFG pair co-located), decoded
    frames are all-gathered once.  All ranks must call with the same global inputs."""

    def __init__(self, pipe, group=None):
        self.pipe = pipe
        self.group = group

    @torch.no_grad()
    def __call__(self, prompt_embeds, negative_prompt_embeds, latents, condition_latent, mask=None, motion=None, **kw):
        world = dist.get_world_size(self.group)
        rank = dist.get_rank(self.group)
        n = prompt_embeds.shape[0]
        if n % world:
            raise ValueError("number of prompts must be a multiple of the world size")
        mine = shard_indices(n, rank, world)
        vids, lats = [], []
        for i in mine:
            m = None if mask is None else (mask if mask.shape[0] == 1 else mask[i: i + 1])
            v, l = self.pipe(prompt_embeds=prompt_embeds[i: i + 1],
                             negative_prompt_embeds=None if negative_prompt_embeds is None else negative_prompt_embeds[i: i + 1],
                             latents=latents[i: i + 1], condition_latent=condition_latent[i: i + 1], mask=m, motion=motion,
                             output_type="pt", return_dict=False, **kw)
            vids.append(to_uint8_frames(v))
            lats.append(l)
        frames = gather_round_robin(torch.cat(vids, dim=0), n, self.group)
        return frames, torch.cat(lats, dim=0)
