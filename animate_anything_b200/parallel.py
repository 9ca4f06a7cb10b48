"""Multi-GPU partitioning of the sampling path (one process per GPU, torch.distributed; NCCL on GPUs, gloo in CPU tests).

What shards: batch elements of the UNet (prompts / CFG halves) and the frames of the VAE; the
diffusion timesteps of one clip do NOT (latents at step i+1 depend on step i, models/pipeline.py:163-192).

  * prompt sharding (weak scaling, config 3): rank r owns prompts  r, r+W, ...  with both CFG halves co-located, so the
    denoising loop needs no communication at all; ONE all-gather of the decoded uint8 frames at the end.
  * cfg split (one clip on 2 GPUs): rank 0 of a pair runs the unconditional, rank 1 the text half; the fp32 noise
    prediction (1.1 MB per half at 16x64x64) is all-gathered every step and both ranks apply the same fused CFG +
    scheduler step, so the latents stay bit-identical on both (every kernel is batch-invariant).
  * frame-sharded VAE decode: rank r decodes a contiguous frame range of the final latents through the SAME fused
    decoder tail as the single-GPU path (`AutoencoderKL.decode_frames_uint8`: bit-equal uint8 frames in diffusers'
    tensor2vid layout [f, H, b*W, 3]); one all-gather of frames.
  * `LatencyShardedPipeline` = config 3 on 4 / 8 GPUs: prompts over pair-groups; inside a pair-group of 2 ranks the CFG
    split + frame-sharded decode, inside a group of 1 rank the plain pipeline; one all-gather of all clips' frames.
"""
from __future__ import annotations

from typing import List, Optional, Tuple

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


def decode_frames_uint8_sharded(vae, latents: torch.Tensor, group=None) -> torch.Tensor:
    """Each rank of `group` decodes its contiguous frame range of latents [b, 4, f, h, w] with the fused uint8 decoder
    tail; returns all frames [f, 8h, b*8w, 3] uint8 (tensor2vid layout) on every rank.  One all-gather; ranks are padded
    to the same frame count.  Bit-identical to `vae.decode_frames_uint8(latents)` on one GPU (frames are independent)."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    f = latents.shape[2]
    lo, hi = frame_range(f, rank, world)
    per = -(-f // world)
    if hi > lo:
        local = vae.decode_frames_uint8(latents[:, :, lo:hi].contiguous())            # [hi-lo, H, bW, 3]
    else:
        probe = vae.decode_frames_uint8(latents[:, :, :1].contiguous())
        local = probe[:0]
    if local.shape[0] < per:
        pad = torch.zeros((per - local.shape[0],) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
        local = torch.cat([local, pad], dim=0)
    g = all_gather_clips(local.unsqueeze(0), group)              # [world, per, H, bW, 3]
    pieces = []
    for r in range(world):
        l2, h2 = frame_range(f, r, world)
        pieces.append(g[r][: h2 - l2])
    return torch.cat(pieces, dim=0)


# round-1 name; same contract (frames layout) as above
decode_video_frame_sharded = decode_frames_uint8_sharded


class PromptShardedPipeline:
    """Weak-scaling driver: every rank runs `LatentToVideoPipeline` on its own prompts (CFG pair co-located), decoded
    frames are all-gathered once.  All ranks must call with the same global inputs.  Returns (frames uint8
    [n_prompts, f, H, W, 3] in global prompt order, this rank's final latents)."""

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
                             output_type="u8", return_dict=False, **kw)
            vids.append(v.unsqueeze(0))                           # [1, f, H, W, 3]
            lats.append(l)
        frames = gather_round_robin(torch.cat(vids, dim=0), n, self.group)
        return frames, torch.cat(lats, dim=0)


class LatencyShardedPipeline:
    """BASELINE config 3 as written: P prompts x (uncond, text) over W GPUs with W = P (CFG pairs co-located, no per-step
    traffic) or W = 2P (one batch element per GPU: each prompt's two halves on a pair of ranks, one all-gather of the
    fp32 noise prediction per step inside the pair, then the pair frame-shards its VAE decode).  ONE world all-gather of
    decoded frames at the end.  `pipe.cfg_group` is set here; every rank must construct this object (new_group is
    collective).  Returns (frames uint8 [P, f, H, W, 3] on every rank, this rank's final latents [1, 4, f, h, w])."""

    def __init__(self, pipe, n_prompts: int):
        self.pipe = pipe
        world = dist.get_world_size()
        rank = dist.get_rank()
        if world == n_prompts:
            self.ranks_per_prompt = 1
        elif world == 2 * n_prompts:
            self.ranks_per_prompt = 2
        else:
            raise ValueError(f"{n_prompts} prompts need {n_prompts} or {2 * n_prompts} ranks, got {world}")
        self.n_prompts = n_prompts
        self.prompt = rank // self.ranks_per_prompt
        self.pair = None
        if self.ranks_per_prompt == 2:
            for p in range(n_prompts):                            # every rank creates every group (collective)
                grp = dist.new_group([2 * p, 2 * p + 1])
                if p == self.prompt:
                    self.pair = grp
            pipe.cfg_group = self.pair

    @torch.no_grad()
    def __call__(self, prompt_embeds, negative_prompt_embeds, latents, condition_latent, mask=None, motion=None, **kw):
        i = self.prompt
        m = None if mask is None else (mask if mask.shape[0] == 1 else mask[i: i + 1])
        args = dict(prompt_embeds=prompt_embeds[i: i + 1], negative_prompt_embeds=negative_prompt_embeds[i: i + 1],
                    latents=latents[i: i + 1], condition_latent=condition_latent[i: i + 1], mask=m, motion=motion,
                    return_dict=False, **kw)
        if self.ranks_per_prompt == 1:
            frames, lat = self.pipe(output_type="u8", **args)     # [f, H, W, 3]
            allf = all_gather_clips(frames.unsqueeze(0))          # [P, f, H, W, 3]
            return allf, lat
        _, lat = self.pipe(output_type="latent", **args)          # both ranks of the pair hold identical latents
        f = lat.shape[2]
        r = dist.get_rank(self.pair)
        lo, hi = frame_range(f, r, 2)
        per = -(-f // 2)
        if hi > lo:
            local = self.pipe.vae.decode_frames_uint8(lat[:, :, lo:hi].contiguous())
        else:                                                     # a one-frame clip: the second rank of the pair has nothing to decode
            sf = self.pipe.vae_scale_factor
            local = torch.zeros((0, lat.shape[3] * sf, lat.shape[0] * lat.shape[4] * sf, 3), dtype=torch.uint8, device=lat.device)
        if local.shape[0] < per:
            local = torch.cat([local, torch.zeros((per - local.shape[0],) + tuple(local.shape[1:]), dtype=local.dtype,
                                                  device=local.device)], dim=0)
        g = all_gather_clips(local.unsqueeze(0))                  # world all-gather: [2P, per, H, W, 3]
        clips = []
        for p in range(self.n_prompts):
            l0, h0 = frame_range(f, 0, 2)
            l1, h1 = frame_range(f, 1, 2)
            clips.append(torch.cat([g[2 * p][: h0 - l0], g[2 * p + 1][: h1 - l1]], dim=0))
        return torch.stack(clips), lat
