"""CPU: host logic of the transparent-video mirror classes (animate_anything_b200/layerdiffuse_VAE.py, row f4) — state_dict
keys, constructor checks, weight-layout conversion (zero-padded attention heads, 32 -> 64 channel padding in front of the
stride-2 convs, the 1x1 latent conv as a residual GEMM), skip bookkeeping and geometry — executed with the kernels replaced by
the plain-torch stand-ins of tests/ops_emulation.py and compared with the outputs of the VERBATIM reference classes
(tests/golden/transparent_ref.pt).  Kernel numerics are the `-m gpu` tests' business (tests/test_gpu_transparent.py)."""
import os
import sys

import pytest
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

from ops_emulation import emulated_ops  # noqa: E402
from oracle.composition import fill_deterministic, oracle_rgba_postprocess  # noqa: E402

GOLD = os.path.join(HERE, "golden", "transparent_ref.pt")


def _host_prepared(model):
    """fp32 / CPU weight prep, installed where `_prepared()` looks for it (the real one insists on a 16-bit CUDA model)."""
    prep = model._build_prepared(torch.float32, torch.device("cpu"))
    model.__dict__["_aab_prepared"] = prep
    return prep


def _decoder(gold):
    from animate_anything_b200.layerdiffuse_VAE import UNet384
    dec = UNet384().eval()
    assert sorted(dec.state_dict().keys()) == gold["dec_keys"], "state_dict keys differ from models/layerdiffuse_VAE.py:UNet384"
    fill_deterministic(dec, seed=7)
    dec.load_state_dict({k: v.bfloat16().float() for k, v in dec.state_dict().items()})
    return dec


def test_unet384_mirror_wiring_matches_reference():
    gold = torch.load(GOLD)
    dec = _decoder(gold)
    for k, v in gold["dec_config"].items():
        got = dec.config[k]
        assert (tuple(got) if isinstance(got, (list, tuple)) else got) == (tuple(v) if isinstance(v, (list, tuple)) else v), k
    with emulated_ops():
        _host_prepared(dec)
        y = dec(gold["dec_x"], gold["dec_latent"])
    assert y.shape == gold["dec_out"].shape
    err = float((y - gold["dec_out"]).abs().max())
    assert err < 2e-4, err


def test_unet384_frame_chunks_and_rgba_tail():
    """decode_rgba_u8 (the fused tail of models/pipeline_stage2.py:303-324): frame chunking, the strided latent view and
    the uint8 RGBA conversion against the oracle's post-processing of the same decoder output."""
    gold = torch.load(GOLD)
    dec = _decoder(gold)
    g = torch.Generator().manual_seed(4)
    f = 3
    video = torch.randn(1, 3, f, 32, 48, generator=g).clamp(-1, 1)
    latents = torch.randn(1, 4, f, 4, 6, generator=g)
    with emulated_ops():
        _host_prepared(dec)
        dec.frame_chunk = 2
        pngs = dec.decode_rgba_u8(video, latents).numpy()
        rgba = dec(video[0].permute(1, 0, 2, 3), latents[0].permute(1, 0, 2, 3))
    ref = oracle_rgba_postprocess(rgba.half(), 1, f)
    assert pngs.shape == ref.shape == (f, 32, 48, 4)
    d = abs(pngs.astype(int) - ref.astype(int))
    assert d[..., :3].max() <= 1 and (d[..., :3] != 0).mean() < 1e-3 and (d[..., 3] != 0).mean() < 1e-3
    assert set(map(int, set(pngs[..., 3].ravel().tolist()))) <= {0, 255}


def test_offset_encoder_mirror_wiring_matches_reference():
    from animate_anything_b200.layerdiffuse_VAE import LatentTransparencyOffsetEncoder
    gold = torch.load(GOLD)
    enc = LatentTransparencyOffsetEncoder().eval()
    assert sorted(enc.state_dict().keys()) == gold["enc_keys"]
    assert float(enc.blocks[16].weight.abs().max()) == 0.0                       # zero_module, :38
    fill_deterministic(enc, seed=8)
    enc.load_state_dict({k: v.bfloat16().float() for k, v in enc.state_dict().items()})
    with emulated_ops():
        _host_prepared(enc)
        e = enc(gold["enc_in"])
    assert e.shape == gold["enc_out"].shape
    assert float((e - gold["enc_out"]).abs().max()) < 2e-5


def test_constructor_and_shape_errors():
    from animate_anything_b200.layerdiffuse_VAE import LatentTransparencyOffsetEncoder, UNet384
    with pytest.raises(ValueError):
        UNet384(down_block_types=("DownBlock2D", "DownBlock2D", "DownBlock2D", "CrossAttnDownBlock2D"))
    with pytest.raises(ValueError):
        UNet384(block_out_channels=(32, 64, 128))
    dec = UNet384().eval()
    assert float(dec.latent_conv_in.weight.abs().max()) == 0.0                   # zero_module, :70
    with pytest.raises(TypeError):                                               # fp32 / CPU model: no fallback
        dec(torch.zeros(1, 3, 64, 64), torch.zeros(1, 4, 8, 8))
    with emulated_ops():
        _host_prepared(dec)
        with pytest.raises(ValueError):
            dec(torch.zeros(1, 3, 60, 64), torch.zeros(1, 4, 8, 8))
    with pytest.raises(TypeError):
        LatentTransparencyOffsetEncoder()(torch.zeros(1, 4, 64, 64))


def test_masked_pipeline_signature_matches_reference():
    """Same keyword names, order and defaults as models/pipeline_stage2.py:172-199 (the trainer calls it with keywords)."""
    import inspect
    from animate_anything_b200.pipeline_stage2 import MaskedLatentToVideoPipeline
    sig = inspect.signature(MaskedLatentToVideoPipeline.__call__)
    expect = [("clean_latents", None), ("vae_alpha_decoder", None), ("prompt", None), ("height", None), ("width", None),
              ("num_frames", 16), ("num_inference_steps", 50), ("guidance_scale", 9.0), ("negative_prompt", None),
              ("eta", 0.0), ("generator", None), ("latents", None), ("condition_latent", None), ("prompt_embeds", None),
              ("negative_prompt_embeds", None), ("output_type", "np"), ("return_dict", True), ("callback", None),
              ("callback_steps", 1), ("cross_attention_kwargs", None), ("timesteps", None), ("mask", None), ("motion", None),
              ("image_embeds", None)]
    got = [(n, p.default) for n, p in sig.parameters.items() if n != "self"]
    assert got == expect


def test_masked_pipeline_host_logic_end_to_end():
    """`MaskedLatentToVideoPipeline.__call__` called UNBOUND on a base pipeline object (train_transparent_i2v_stage2.py:500-515) with
    every kernel emulated: UNet3D loop -> one VAE decoder pass with both tails -> UNet384 -> RGBA bytes, against
    `oracle_masked_sampling_loop` (pinned to the verbatim reference call by tests/test_oracle_golden.py)."""
    import numpy as np
    from oracle.composition import (AutoencoderKL as OVAE, DDIMScheduler as ODDIM, OracleUNet3D, OracleUNet384,
                                    oracle_masked_sampling_loop)
    from animate_anything_b200 import schedulers as S
    from animate_anything_b200.autoencoder_kl import AutoencoderKL
    from animate_anything_b200.layerdiffuse_VAE import UNet384
    from animate_anything_b200.pipeline_stage2 import MaskedLatentToVideoPipeline, TextToVideoSDPipeline
    from animate_anything_b200.unet_3d_condition_mask import UNet3DConditionModel
    ucfg = dict(sample_size=16, block_out_channels=(64, 64, 128, 128), attention_head_dim=64, cross_attention_dim=64,
                motion_mask=True, motion_strength=True)
    vcfg = dict(block_out_channels=(64, 64, 64, 64), layers_per_block=1, sample_size=64)
    sk = dict(beta_start=0.00085, beta_end=0.012, beta_schedule="scaled_linear", clip_sample=False, set_alpha_to_one=False,
              steps_offset=1)
    ounet = fill_deterministic(OracleUNet3D(**{k: v for k, v in ucfg.items() if k != "sample_size"}).eval(), 0)
    ovae = fill_deterministic(OVAE(**vcfg).eval(), 1)
    odec = fill_deterministic(OracleUNet384().eval(), 7)
    unet, vae, dec = UNet3DConditionModel(**ucfg).eval(), AutoencoderKL(**vcfg).eval(), UNet384().eval()
    unet.load_state_dict(ounet.state_dict())
    vae.load_state_dict(ovae.state_dict())
    dec.load_state_dict(odec.state_dict())
    g = torch.Generator().manual_seed(6)
    lat = torch.randn(1, 4, 2, 8, 8, generator=g)
    cond = torch.randn(1, 4, 1, 8, 8, generator=g)
    pe, ne = torch.randn(1, 7, 64, generator=g), torch.randn(1, 7, 64, generator=g)
    mask1 = (torch.rand(1, 1, 1, 8, 8, generator=g) > 0.5).float()
    with torch.no_grad():
        ref_vid, ref_lat, ref_png = oracle_masked_sampling_loop(ounet, ODDIM(**sk), ovae, odec, lat, pe, ne, cond, mask1, [5],
                                                                guidance_scale=9.0, num_inference_steps=2)
    pipeline = TextToVideoSDPipeline(vae=vae, text_encoder=None, tokenizer=None, unet=unet, scheduler=S.DDIMScheduler(**sk))
    pipeline.use_cuda_graph = False
    with emulated_ops():
        for m in (unet, vae, dec):
            _host_prepared(m)
        video, latents, pngs, alpha_png, pngs_rgb = MaskedLatentToVideoPipeline.__call__(
            pipeline, clean_latents=lat.clone(), vae_alpha_decoder=dec, prompt_embeds=pe, negative_prompt_embeds=ne, latents=lat,
            width=64, height=64, num_frames=2, num_inference_steps=2, guidance_scale=9.0, motion=[5], return_dict=False,
            condition_latent=cond, mask=mask1)
        with pytest.raises(TypeError):
            MaskedLatentToVideoPipeline.__call__(pipeline, vae_alpha_decoder=torch.nn.Identity(), prompt_embeds=pe,
                                                 negative_prompt_embeds=ne, latents=lat, condition_latent=cond, mask=mask1)
    assert float((latents - ref_lat).abs().max()) < 2e-3 * max(1.0, float(ref_lat.abs().mean()))
    assert len(video) == 2 and video[0].shape == (64, 64, 3) and video[0].dtype == np.uint8
    want_frames = ((ref_vid * 0.5 + 0.5).clamp(0, 1) * 255).to(torch.uint8)[0].permute(1, 2, 3, 0).numpy()
    assert np.abs(np.stack(video).astype(int) - want_frames.astype(int)).max() <= 1
    assert pngs.shape == ref_png.shape == (2, 64, 64, 4)
    d = np.abs(pngs.astype(int) - ref_png.astype(int))
    # the emulated RGBA tail rounds through fp16 like the 16-bit product does; the oracle here ran in fp32
    assert d[..., :3].mean() < 0.6 and (d[..., 3] != 0).mean() < 5e-3
    assert (alpha_png == pngs[..., 3]).all() and (pngs_rgb == pngs[..., :3]).all()
