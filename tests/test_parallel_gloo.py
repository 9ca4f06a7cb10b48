"""CPU, world_size 2 over gloo: the host-side partitioning logic of the N>1 path (prompt sharding, frame sharding, the
single gather) with a stand-in pipeline/VAE (the CUDA kernels themselves are covered by the -m gpu tests)."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


class FakeVAE:
    def decode_video(self, lat):
        b, c, f, h, w = lat.shape
        return (lat[:, :3].float().repeat_interleave(2, -1).repeat_interleave(2, -2) / 4).clamp(-1, 1)

    def decode_frames_uint8(self, lat):
        """tensor2vid layout and rounding: [f, H, b*W, 3] uint8."""
        v = self.decode_video(lat).mul(0.5).add(0.5).clamp(0, 1)
        b, c, f, h, w = v.shape
        return (v.permute(2, 3, 0, 4, 1).reshape(f, h, b * w, c) * 255).to(torch.uint8)


class FakePipe:
    vae = FakeVAE()
    cfg_group = None

    def __call__(self, prompt_embeds, negative_prompt_embeds, latents, condition_latent, mask, motion, output_type,
                 return_dict, **kw):
        lat = latents + prompt_embeds.mean()
        if output_type == "latent":
            return lat, lat
        return FakeVAE().decode_frames_uint8(lat), latents + 1


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from animate_anything_b200 import parallel as P
    try:
        assert P.shard_indices(5, rank, world) == [i for i in range(5) if i % world == rank]
        # frame-sharded decode equals the single-process decode
        g = torch.Generator().manual_seed(0)
        lat = torch.randn(1, 4, 5, 4, 4, generator=g)
        full = FakeVAE().decode_frames_uint8(lat)
        got = P.decode_frames_uint8_sharded(FakeVAE(), lat)
        assert torch.equal(got, full)
        # prompt-sharded pipeline: frames come back in global prompt order on every rank
        pe = torch.randn(4, 7, 8, generator=g)
        lats = torch.randn(4, 4, 3, 4, 4, generator=g)
        cond = torch.randn(4, 4, 1, 4, 4, generator=g)
        frames, my_lat = P.PromptShardedPipeline(FakePipe())(pe, pe, lats, cond)
        ref = torch.stack([FakePipe()(pe[i:i + 1], None, lats[i:i + 1], None, None, None, "u8", False)[0]
                           for i in range(4)])
        assert torch.equal(frames, ref)
        assert my_lat.shape[0] == 2
        # config-3 driver: 2 prompts on 2 ranks (pairs co-located) and 1 prompt on 2 ranks (CFG halves split + frame-
        # sharded decode): both return every clip's frames on every rank
        lp = P.LatencyShardedPipeline(FakePipe(), 2)
        f2, l2 = lp(pe[:2], pe[:2], lats[:2], cond[:2])
        assert torch.equal(f2, ref[:2]) and l2.shape[0] == 1
        fp = FakePipe()
        lp1 = P.LatencyShardedPipeline(fp, 1)
        assert fp.cfg_group is not None
        f1, _ = lp1(pe[:1], pe[:1], lats[:1], cond[:1])
        assert torch.equal(f1, ref[:1])
        q.put((rank, "ok"))
    except Exception as e:  # pragma: no cover
        q.put((rank, repr(e)))
    finally:
        dist.destroy_process_group()


def test_world2_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 1000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert all(r[1] == "ok" for r in res), res


def test_frame_range_partition():
    from animate_anything_b200.parallel import frame_range
    for f in (16, 17, 5, 1):
        for w in (1, 2, 4, 8):
            ranges = [frame_range(f, r, w) for r in range(w)]
            assert ranges[0][0] == 0 and ranges[-1][1] == f
            assert all(ranges[i][1] == ranges[i + 1][0] for i in range(w - 1))
