"""GPU parity of AutoencoderKL encode/decode and of the full LatentToVideoPipeline.__call__ loop against the oracle
(fp32 math on the same 16-bit-rounded weights).  Full-loop errors are reported as measured and bounded relative to the
stock PyTorch 16-bit execution of the same loop (see test_gpu_unet.py docstring)."""
import os
import sys

import pytest
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from util import report  # noqa: E402

pytestmark = pytest.mark.gpu

UNET = dict(sample_size=16, block_out_channels=(64, 128, 256, 256), attention_head_dim=64, cross_attention_dim=128,
            motion_mask=True, motion_strength=True)
VAE = dict(block_out_channels=(64, 64, 128, 128), layers_per_block=1, sample_size=128)
SCHED = dict(beta_start=0.00085, beta_end=0.012, beta_schedule="scaled_linear", clip_sample=False,
             set_alpha_to_one=False, steps_offset=1)


def _pair(oracle_cls, ours_cls, cfg, dtype, seed, drop=("sample_size",)):
    from oracle.composition import fill_deterministic
    ocfg = {k: v for k, v in cfg.items() if k not in drop} if oracle_cls.__name__.startswith("Oracle") else cfg
    oracle = fill_deterministic(oracle_cls(**ocfg).eval(), seed=seed)
    sd16 = {k: v.to(dtype) for k, v in oracle.state_dict().items()}
    oracle.load_state_dict({k: v.float() for k, v in sd16.items()})
    ours = ours_cls(**cfg).eval()
    ours.load_state_dict(sd16, strict=True)
    return oracle.cuda(), ours.to(dtype).cuda()


def _no_tf32():
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False


@pytest.mark.parametrize("dtype,hw", [(torch.float16, (64, 96)), (torch.bfloat16, (64, 96)), (torch.float16, (120, 136))])
def test_vae_encode_decode(dtype, hw):
    """(120, 136): latent 15 x 17 -- what train.py:738-742 produces for a 170x150 prompt image; H*W of the mid-block
    attention is then not a multiple of 8 (zero-padded K of the P.V product)."""
    from oracle.composition import AutoencoderKL as OVAE, oracle_decode_latents, oracle_encode_image
    from animate_anything_b200.autoencoder_kl import AutoencoderKL
    _no_tf32()
    ovae, vae = _pair(OVAE, AutoencoderKL, VAE, dtype, seed=1)
    g = torch.Generator().manual_seed(3)
    img = torch.randn(2, 3, hw[0], hw[1], generator=g).to(dtype).cuda()
    lat = torch.randn(1, 4, 3, hw[0] // 8, hw[1] // 8, generator=g).to(dtype).cuda()
    with torch.no_grad():
        ref_mean = ovae.encode(img.float()).latent_dist.mode()
        ref_vid = oracle_decode_latents(ovae, lat.float())
        st_mean = ovae.to(dtype).encode(img).latent_dist.mode().float()
        st_vid = oracle_decode_latents(ovae, lat)
    mean = vae.encode(img).latent_dist.mode()
    vid = vae.decode_video(lat)
    assert vid.dtype == torch.float32 and vid.shape == ref_vid.shape
    for name, got, ref, stock in (("encode", mean, ref_mean, st_mean), ("decode", vid, ref_vid, st_vid)):
        e = (got.float() - ref).abs()
        es = (stock.float() - ref).abs()
        sc = ref.abs().mean().item()
        print(f"vae {name} {dtype}: ref|mean|={sc:.4f} ours max={e.max().item():.3e} mean={e.mean().item():.3e} | "
              f"stock max={es.max().item():.3e} mean={es.mean().item():.3e}")
        assert e.mean().item() <= 1.5 * es.mean().item() + 2e-4 * sc
        assert e.max().item() <= 2.5 * es.max().item() + 2e-3 * sc
    # diffusers-surface decode(): [N,4,h,w] -> .sample [N,3,H,W] in model dtype
    img2 = vae.decode(lat[0].permute(1, 0, 2, 3)).sample
    assert img2.shape == (3, 3, hw[0], hw[1]) and img2.dtype == dtype


@pytest.mark.parametrize("sched_name,steps", [("ddim", 3), ("dpm", 4)])
def test_full_loop(sched_name, steps):
    from oracle.composition import (AutoencoderKL as OVAE, DDIMScheduler as ODDIM, DPMSolverMultistepScheduler as ODPM,
                                    OracleUNet3D, oracle_sampling_loop)
    from animate_anything_b200 import schedulers as S
    from animate_anything_b200.autoencoder_kl import AutoencoderKL
    from animate_anything_b200.pipeline import LatentToVideoPipeline
    from animate_anything_b200.unet_3d_condition_mask import UNet3DConditionModel
    _no_tf32()
    dtype = torch.float16
    ounet, unet = _pair(OracleUNet3D, UNet3DConditionModel, UNET, dtype, seed=0)
    ovae, vae = _pair(OVAE, AutoencoderKL, VAE, dtype, seed=1)
    osched = ODDIM(**SCHED)
    sched = S.DDIMScheduler(**SCHED)
    if sched_name == "dpm":
        osched = ODPM.from_config(osched.config)
        sched = S.DPMSolverMultistepScheduler.from_config(sched.config)
    g = torch.Generator().manual_seed(5)
    lat = torch.randn(1, 4, 4, 16, 16, generator=g).to(dtype).cuda()
    cond = torch.randn(1, 4, 1, 16, 16, generator=g).to(dtype).cuda()
    pe = torch.randn(1, 77, 128, generator=g).to(dtype).cuda()
    ne = torch.randn(1, 77, 128, generator=g).to(dtype).cuda()
    mask = torch.ones(1, 1, 1, 16, 16, dtype=dtype, device="cuda")
    ref_vid, ref_lat = oracle_sampling_loop(ounet, osched, lat.float(), pe.float(), ne.float(), cond.float(),
                                            mask.float(), [4], 9.0, steps, vae=ovae)
    pipe = LatentToVideoPipeline(vae=vae, text_encoder=None, tokenizer=None, unet=unet, scheduler=sched)
    vid, out_lat = pipe(prompt_embeds=pe, negative_prompt_embeds=ne, latents=lat, condition_latent=cond, mask=mask,
                        motion=[4], guidance_scale=9.0, num_inference_steps=steps, output_type="pt", return_dict=False)
    assert pipe.last_gpu_launches > 100
    # CUDA-graph replay must give bit-identical latents
    pipe.use_cuda_graph = True
    vid_g, lat_g = pipe(prompt_embeds=pe, negative_prompt_embeds=ne, latents=lat, condition_latent=cond, mask=mask,
                        motion=[4], guidance_scale=9.0, num_inference_steps=steps, output_type="pt", return_dict=False)
    assert torch.equal(lat_g, out_lat), float((lat_g.float() - out_lat.float()).abs().max())
    # stock 16-bit torch loop as yard-stick
    if sched_name == "dpm":
        osched = ODPM.from_config(osched.config)
    st_vid, st_lat = oracle_sampling_loop(ounet.to(dtype), osched, lat, pe, ne, cond, mask, [4], 9.0, steps, vae=None)
    e = (out_lat.float() - ref_lat).abs()
    es = (st_lat.float() - ref_lat).abs()
    sc = ref_lat.abs().mean().item()
    print(f"loop {sched_name} x{steps}: latents ref|mean|={sc:.4f} ours max={e.max().item():.3e} mean={e.mean().item():.3e}"
          f" | stock fp16 max={es.max().item():.3e} mean={es.mean().item():.3e}")
    ev = (vid - ref_vid).abs()
    print(f"  video: ref|mean|={ref_vid.abs().mean().item():.4f} max={ev.max().item():.3e} mean={ev.mean().item():.3e}")
    # the loop amplifies rounding noise (CFG 9 x 1/sqrt(alpha_t)); bound relative to the stock 16-bit loop
    assert e.mean().item() <= 3.0 * es.mean().item() + 5e-4 * sc
    assert e.max().item() <= 4.0 * es.max().item() + 5e-3 * sc
    # uint8 frames path (tensor2vid) as the reference returns them
    frames, _ = pipe(prompt_embeds=pe, negative_prompt_embeds=ne, latents=lat, condition_latent=cond, mask=mask,
                     motion=[4], guidance_scale=9.0, num_inference_steps=steps, return_dict=False)
    assert len(frames) == 4 and frames[0].shape == (128, 128, 3) and frames[0].dtype.name == "uint8"
    # the fused uint8 path is bit-identical to diffusers' tensor2vid applied to the float video
    from animate_anything_b200.pipeline import tensor2vid
    ref_frames = tensor2vid(vid.clone())
    import numpy as np
    for a, b2 in zip(frames, ref_frames):
        assert np.array_equal(a, b2)


@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16])
def test_common_ddpm_forward_timesteps_bit_exact(dtype):
    """utils/common.py:32-48 through aab_add_noise: same seed -> same torch.randn noise -> bit-identical x_t (the
    kernel reproduces the three 16-bit roundings of repeat + scheduler.add_noise), same kept timesteps."""
    from oracle.composition import DDIMScheduler as OSched, oracle_ddpm_forward_timesteps
    from animate_anything_b200.schedulers import DDIMScheduler
    from animate_anything_b200.common import DDPM_forward_timesteps
    osched, sched = OSched(**SCHED), DDIMScheduler(**SCHED)
    osched.set_timesteps(25)
    sched.set_timesteps(25)
    g = torch.Generator().manual_seed(2)
    for shape in ((1, 4, 1, 64, 64), (2, 4, 1, 24, 40), (1, 4, 6, 16, 16)):
        x0 = torch.randn(*shape, generator=g).to(dtype).cuda()
        torch.manual_seed(7)
        xt, ts = DDPM_forward_timesteps(x0, 10, 16, sched)
        torch.manual_seed(7)
        ref, ots = oracle_ddpm_forward_timesteps(x0, 10, 16, osched)
        assert xt.dtype == dtype and xt.shape == ref.shape
        assert torch.equal(torch.as_tensor(ts).cpu(), torch.as_tensor(ots).cpu())
        assert torch.equal(xt, ref), f"{shape} {dtype}: max diff {(xt.float() - ref.float()).abs().max().item()}"


@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16])
def test_common_tensor_to_vae_latent(dtype):
    """utils/common.py:12-20 fused into the encoder tail: [b, f, 3, H, W] -> [b, 4, f, h, w] * 0.18215."""
    from oracle.composition import AutoencoderKL as OVAE, oracle_encode_image
    from animate_anything_b200.autoencoder_kl import AutoencoderKL
    from animate_anything_b200.common import tensor_to_vae_latent
    _no_tf32()
    ovae, vae = _pair(OVAE, AutoencoderKL, VAE, dtype, seed=1)
    g = torch.Generator().manual_seed(4)
    frames = torch.randn(2, 3, 3, 64, 32, generator=g).clamp(-1, 1).to(dtype).cuda()
    with torch.no_grad():
        ref = oracle_encode_image(ovae, frames.float())
        stock = oracle_encode_image(ovae.to(dtype), frames).float()
    lat = tensor_to_vae_latent(frames, vae)
    assert lat.shape == ref.shape == (2, 4, 3, 8, 4) and lat.dtype == dtype
    e, es, sc = (lat.float() - ref).abs(), (stock - ref).abs(), ref.abs().mean().item()
    print(f"tensor_to_vae_latent {dtype}: ref|mean|={sc:.4f} ours max={e.max().item():.3e} mean={e.mean().item():.3e} | "
          f"stock max={es.max().item():.3e} mean={es.mean().item():.3e}")
    assert e.mean().item() <= 1.5 * es.mean().item() + 2e-4 * sc
    assert e.max().item() <= 2.5 * es.max().item() + 2e-3 * sc
    # frame chunking path (f > frame_chunk) gives the same latents
    vae.frame_chunk = 2
    assert torch.equal(tensor_to_vae_latent(frames, vae), lat)


def test_eval_flow_like_train_eval():
    """The sequence of train.py:731-770 `eval()` on small models: image -> VaeImageProcessor.preprocess ->
    tensor_to_vae_latent -> DDPM_forward_timesteps(forward_t) -> pipeline(latents=, condition_latent=, mask=, motion=,
    timesteps=kept) with DPM-Solver++ (what train.py:806 installs), against the same sequence on the fp32 oracle."""
    import numpy as np
    from oracle.composition import (AutoencoderKL as OVAE, DDIMScheduler as ODDIM, DPMSolverMultistepScheduler as ODPM,
                                    OracleUNet3D, oracle_ddpm_forward_timesteps, oracle_encode_image,
                                    oracle_sampling_loop)
    from animate_anything_b200 import schedulers as S
    from animate_anything_b200.autoencoder_kl import AutoencoderKL
    from animate_anything_b200.common import DDPM_forward_timesteps, tensor_to_vae_latent
    from animate_anything_b200.image_processor import VaeImageProcessor
    from animate_anything_b200.pipeline import LatentToVideoPipeline
    from animate_anything_b200.unet_3d_condition_mask import UNet3DConditionModel
    _no_tf32()
    dtype = torch.float16
    ounet, unet = _pair(OracleUNet3D, UNet3DConditionModel, UNET, dtype, seed=0)
    ovae, vae = _pair(OVAE, AutoencoderKL, VAE, dtype, seed=1)
    osched = ODPM.from_config(ODDIM(**SCHED).config)
    sched = S.DPMSolverMultistepScheduler.from_config(S.DDIMScheduler(**SCHED).config)
    steps, forward_t, frames = 6, 4, 4
    osched.set_timesteps(steps, device="cuda")
    sched.set_timesteps(steps, device="cuda")
    rng = np.random.default_rng(0)
    img = rng.random((128, 128, 3)).astype(np.float32)                 # H, W, C in [0, 1]
    inp = VaeImageProcessor().preprocess(img, 128, 128)                # [1, 3, 128, 128] in [-1, 1]
    inp = inp.unsqueeze(0).to(dtype).cuda()                            # b f c h w (train.py:746)
    cond = tensor_to_vae_latent(inp, vae)
    torch.manual_seed(11)
    init, ts = DDPM_forward_timesteps(cond, forward_t, frames, sched)
    assert init.shape == (1, 4, frames, 16, 16) and len(ts) == forward_t
    g = torch.Generator().manual_seed(5)
    pe = torch.randn(1, 77, 128, generator=g).to(dtype).cuda()
    ne = torch.randn(1, 77, 128, generator=g).to(dtype).cuda()
    mask = (torch.rand(1, 1, 1, 16, 16, generator=g) > 0.4).to(dtype).cuda()
    pipe = LatentToVideoPipeline(vae=vae, text_encoder=None, tokenizer=None, unet=unet, scheduler=sched)
    video, lat = pipe(prompt_embeds=pe, negative_prompt_embeds=ne, latents=init, condition_latent=cond, mask=mask,
                      motion=[5], guidance_scale=7.5, num_inference_steps=steps, timesteps=ts, return_dict=False)
    assert len(video) == frames and video[0].shape == (128, 128, 3)
    # the same sequence on the oracle: fp32 reference and stock 16-bit yard-stick, fed the same noise
    torch.manual_seed(11)
    noise = torch.randn(init.shape, dtype=dtype, device="cuda")
    with torch.no_grad():
        ocond = oracle_encode_image(ovae, inp.float())
        oinit, ots = oracle_ddpm_forward_timesteps(ocond, forward_t, frames, osched, noise=noise.float())
        assert [int(t) for t in ots] == [int(t) for t in ts]
        _, ref = oracle_sampling_loop(ounet, osched, oinit, pe.float(), ne.float(), ocond, mask.float(), [5], 7.5, steps,
                                      timesteps=ots)
        osched2 = ODPM.from_config(ODDIM(**SCHED).config)
        _, stock = oracle_sampling_loop(ounet.to(dtype), osched2, init, pe, ne, cond, mask, [5], 7.5, steps, timesteps=ots)
    e, es, sc = (lat.float() - ref).abs(), (stock.float() - ref).abs(), ref.abs().mean().item()
    print(f"eval flow: latents ref|mean|={sc:.4f} ours max={e.max().item():.3e} mean={e.mean().item():.3e} | "
          f"stock fp16 max={es.max().item():.3e} mean={es.mean().item():.3e}")
    assert e.mean().item() <= 3.0 * es.mean().item() + 5e-4 * sc
    assert e.max().item() <= 4.0 * es.max().item() + 5e-3 * sc
