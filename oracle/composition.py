"""ORACLE (test infrastructure, not product): plain-PyTorch restatement of the reference's *composition* of the hot
path, written on top of the diffusers shim (`oracle/shim/diffusers/_impl.py`, parity unpinned — see its header).

  OracleUNet3D          <- models/unet_3d_condition_mask.py:54-526  (UNet3DConditionModel)
  _OBlock               <- models/unet_3d_blocks.py:234-842         (the five block classes, one table-driven class)
  oracle_sampling_loop  <- models/pipeline.py:14-214                (LatentToVideoPipeline.__call__)

It travels to the GPU box (where /root/reference does not exist) and is pinned in THIS container against the verbatim
reference files: `tests/golden/make_golden.py` imports /root/reference/models/*.py over the shim, runs both on the same
seeded weights/inputs and stores the reference outputs as fixtures; `tests/test_oracle_golden.py` re-checks the
restatement against those fixtures everywhere.  Sub-module names equal the reference's, so one state_dict fits both.
"""
from __future__ import annotations

import os
import sys
from typing import List, Optional

import torch
import torch.nn as nn

_SHIM = os.path.join(os.path.dirname(os.path.abspath(__file__)), "shim")
if _SHIM not in sys.path:
    sys.path.insert(0, _SHIM)

from diffusers._impl import (AutoencoderKL, DDIMScheduler, DPMSolverMultistepScheduler, Downsample2D,  # noqa: E402
                             ResnetBlock2D, TemporalConvLayer, TimestepEmbedding, Timesteps, Transformer2DModel,
                             TransformerTemporalModel, Upsample2D, tensor2vid)


class _OBlock(nn.Module):
    """One UNet stage.  kind: 'down' | 'mid' | 'up'; cross=True adds the (spatial, temporal) transformer pair."""

    def __init__(self, kind, cross, resnet_io, temb_ch, eps, groups, head_ch, cross_dim, scale=1.0, resample_ch=None,
                 down_padding=1):
        super().__init__()
        self.kind, self.cross = kind, cross
        self.resnets = nn.ModuleList([ResnetBlock2D(in_channels=i, out_channels=o, temb_channels=temb_ch, eps=eps,
                                                    groups=groups, output_scale_factor=scale) for i, o in resnet_io])
        self.temp_convs = nn.ModuleList([TemporalConvLayer(o, o, dropout=0.1) for _, o in resnet_io])
        if cross:
            n_attn = len(resnet_io) - (1 if kind == "mid" else 0)
            ch = resnet_io[-1][1]
            self.attentions = nn.ModuleList([
                Transformer2DModel(ch // head_ch, head_ch, in_channels=ch, num_layers=1, cross_attention_dim=cross_dim,
                                   norm_num_groups=groups, use_linear_projection=True) for _ in range(n_attn)])
            self.temp_attentions = nn.ModuleList([
                TransformerTemporalModel(ch // head_ch, head_ch, in_channels=ch, num_layers=1,
                                         cross_attention_dim=cross_dim, norm_num_groups=groups) for _ in range(n_attn)])
        if kind == "down":
            self.downsamplers = (nn.ModuleList([Downsample2D(resample_ch, use_conv=True, out_channels=resample_ch,
                                                             padding=down_padding, name="op")])
                                 if resample_ch else None)
        if kind == "up":
            self.upsamplers = (nn.ModuleList([Upsample2D(resample_ch, use_conv=True, out_channels=resample_ch)])
                               if resample_ch else None)

    def _spatio_temporal(self, i, h, ehs, nf):
        h = self.attentions[i](h, encoder_hidden_states=ehs).sample
        if nf > 1:
            h = self.temp_attentions[i](h, num_frames=nf).sample
        return h

    def forward(self, h, temb, ehs, nf, skips: Optional[List[torch.Tensor]] = None, upsample_size=None):
        if self.kind == "mid":
            h = self.resnets[0](h, temb)
            h = self.temp_convs[0](h, num_frames=nf)
            for i in range(len(self.attentions)):
                h = self._spatio_temporal(i, h, ehs, nf)
                h = self.resnets[i + 1](h, temb)
                if nf > 1:
                    h = self.temp_convs[i + 1](h, num_frames=nf)
            return h
        produced = []
        for i, (res, tconv) in enumerate(zip(self.resnets, self.temp_convs)):
            if self.kind == "up":
                h = torch.cat([h, skips.pop()], dim=1)
            h = res(h, temb)
            if nf > 1:
                h = tconv(h, num_frames=nf)
            if self.cross:
                h = self._spatio_temporal(i, h, ehs, nf)
            produced.append(h)
        if self.kind == "down":
            if self.downsamplers is not None:
                h = self.downsamplers[0](h)
                produced.append(h)
            return h, produced
        if self.upsamplers is not None:
            h = self.upsamplers[0](h, upsample_size)
        return h


class OracleUNet3D(nn.Module):
    def __init__(self, sample_size=None, in_channels=4, out_channels=4, block_out_channels=(320, 640, 1280, 1280),
                 layers_per_block=2, downsample_padding=1, mid_block_scale_factor=1, norm_num_groups=32, norm_eps=1e-5,
                 cross_attention_dim=1024, attention_head_dim=64, motion_mask=False, motion_strength=False,
                 cross_stages=(True, True, True, False)):
        super().__init__()
        self.motion_mask, self.motion_strength = motion_mask, motion_strength
        ch = list(block_out_channels)
        c0, temb_ch, g, e = ch[0], ch[0] * 4, norm_num_groups, norm_eps
        self.conv_in = nn.Conv2d(in_channels, c0, 3, padding=1)
        self.conv_in2 = nn.Conv2d(5, c0, 3, padding=1)
        self.time_proj = Timesteps(c0, True, 0)
        self.time_embedding = TimestepEmbedding(c0, temb_ch, act_fn="silu", cond_proj_dim=c0)
        self.motion_proj = Timesteps(c0, True, 0)
        self.motion_embedding = nn.Sequential(nn.Linear(c0, temb_ch), nn.SiLU(), nn.Linear(temb_ch, temb_ch))
        nn.init.zeros_(self.motion_embedding[-1].weight)
        nn.init.zeros_(self.motion_embedding[-1].bias)
        self.transformer_in = TransformerTemporalModel(num_attention_heads=8, attention_head_dim=attention_head_dim,
                                                       in_channels=c0, num_layers=1)
        hd = attention_head_dim
        nst = len(ch)
        self.down_blocks = nn.ModuleList()
        prev = c0
        for i in range(nst):
            io = [(prev if j == 0 else ch[i], ch[i]) for j in range(layers_per_block)]
            self.down_blocks.append(_OBlock("down", cross_stages[i], io, temb_ch, e, g, hd, cross_attention_dim,
                                            resample_ch=ch[i] if i < nst - 1 else None,
                                            down_padding=downsample_padding))
            prev = ch[i]
        self.mid_block = _OBlock("mid", True, [(ch[-1], ch[-1])] * 2, temb_ch, e, g, hd, cross_attention_dim,
                                 scale=mid_block_scale_factor)
        self.up_blocks = nn.ModuleList()
        rev = ch[::-1]
        rev_cross = list(cross_stages)[::-1]
        prev_out = rev[0]
        for i in range(nst):
            out_c = rev[i]
            in_c = rev[min(i + 1, nst - 1)]
            nl = layers_per_block + 1
            io = [((prev_out if j == 0 else out_c) + (in_c if j == nl - 1 else out_c), out_c) for j in range(nl)]
            self.up_blocks.append(_OBlock("up", rev_cross[i], io, temb_ch, e, g, hd, cross_attention_dim,
                                          resample_ch=out_c if i < nst - 1 else None))
            prev_out = out_c
        self.conv_norm_out = nn.GroupNorm(g, c0, eps=e)
        self.conv_act = nn.SiLU()
        self.conv_out = nn.Conv2d(c0, out_channels, 3, padding=1)

    @property
    def dtype(self):
        return self.conv_out.weight.dtype

    def forward(self, sample, timestep, encoder_hidden_states, condition_latent, mask, motion=None, timestep_cond=None):
        x = torch.cat([condition_latent, sample], dim=2)                      # b c T h w, T = F + 1
        b, _, nf, hh, ww = x.shape
        ts = timestep if torch.is_tensor(timestep) else torch.tensor([timestep], device=x.device)
        ts = ts.reshape(-1).to(x.device).expand(b)
        t_emb = self.time_proj(ts).to(self.dtype)
        cond = timestep_cond                                                  # :418-419: used as given when no motion value
        if self.motion_strength and motion is not None:
            cond = self.motion_proj(motion).to(self.dtype)
        emb = self.time_embedding(t_emb, cond).repeat_interleave(nf, dim=0)
        ehs = encoder_hidden_states.repeat_interleave(nf, dim=0)
        if self.motion_mask and mask is not None:
            rep = b // mask.shape[0]
            m = mask.repeat(rep, 1, nf, 1, 1)                                # '(t b) 1 f h w'
            x = torch.cat([m, x], dim=1)
            x = self.conv_in2(x.permute(0, 2, 1, 3, 4).reshape(b * nf, -1, hh, ww))
        else:
            x = self.conv_in(x.permute(0, 2, 1, 3, 4).reshape(b * nf, -1, hh, ww))
        if nf > 1:
            x = self.transformer_in(x, num_frames=nf).sample
        skips = [x]
        for blk in self.down_blocks:
            x, produced = blk(x, emb, ehs, nf)
            skips.extend(produced)
        x = self.mid_block(x, emb, ehs, nf)
        # models/unet_3d_condition_mask.py:377-383,486-491: latent sizes that are not multiples of 2**num_upsamplers make
        # every non-final up block interpolate to the size of the next skip tensor instead of x2
        fwd_size = any(s % (2 ** (len(self.up_blocks) - 1)) != 0 for s in (hh, ww))
        for i, blk in enumerate(self.up_blocks):
            up_size = None
            if i < len(self.up_blocks) - 1 and fwd_size:
                up_size = skips[-len(blk.resnets) - 1].shape[2:]
            x = blk(x, emb, ehs, nf, skips, upsample_size=up_size)
        x = self.conv_out(self.conv_act(self.conv_norm_out(x)))
        x = x.reshape(b, nf, -1, hh, ww).permute(0, 2, 1, 3, 4)
        return x[:, :, 1:]


def oracle_decode_latents(vae, latents):
    """diffusers TextToVideoSDPipeline.decode_latents (called at models/pipeline.py:200)."""
    z = latents / vae.config.scaling_factor
    b, c, f, h, w = z.shape
    img = vae.decode(z.permute(0, 2, 1, 3, 4).reshape(b * f, c, h, w)).sample
    return img.reshape(b, f, -1, img.shape[-2], img.shape[-1]).permute(0, 2, 1, 3, 4).float()


def oracle_encode_image(vae, frames):
    """utils/common.py:12-20 tensor_to_vae_latent: frames [b, f, 3, H, W] -> latents [b, 4, f, h, w] * 0.18215."""
    b, f = frames.shape[:2]
    lat = vae.encode(frames.reshape(b * f, *frames.shape[2:])).latent_dist.mode()
    return lat.reshape(b, f, *lat.shape[1:]).permute(0, 2, 1, 3, 4) * 0.18215


def oracle_ddpm_forward_timesteps(x0, step, num_frames, scheduler, noise=None):
    """utils/common.py:32-48 DDPM_forward_timesteps restated (noise drawn with torch.randn when not given, as the
    reference does): keep the last `step` timesteps, repeat the single-frame latent, add noise at the first kept one."""
    timesteps = scheduler.timesteps[len(scheduler.timesteps) - step:]
    t = timesteps[0]
    xt = x0.repeat(1, 1, num_frames, 1, 1) if x0.shape[2] == 1 else x0
    if noise is None:
        noise = torch.randn(xt.shape, dtype=xt.dtype, device=x0.device)
    tt = torch.tensor([int(t)] * xt.shape[0], device=x0.device)
    return scheduler.add_noise(xt, noise, tt), timesteps


@torch.no_grad()
def oracle_sampling_loop(unet, scheduler, latents, prompt_embeds, negative_prompt_embeds, condition_latent, mask,
                         motion, guidance_scale=9.0, num_inference_steps=50, timesteps=None, vae=None,
                         output_type="pt"):
    """Restates LatentToVideoPipeline.__call__ (models/pipeline.py:107-212) for pre-computed prompt embeddings."""
    cfg = guidance_scale > 1.0
    ehs = torch.cat([negative_prompt_embeds, prompt_embeds]) if cfg else prompt_embeds
    scheduler.set_timesteps(num_inference_steps, device=latents.device)
    if timesteps is None:
        timesteps = scheduler.timesteps
    cond = torch.cat([condition_latent, condition_latent]) if cfg else condition_latent
    mot = None if motion is None else torch.tensor(motion, device=latents.device)
    for t in timesteps:
        inp = torch.cat([latents] * 2) if cfg else latents
        inp = scheduler.scale_model_input(inp, t)
        eps = unet(inp, t, ehs, condition_latent=cond, mask=mask, motion=mot)
        if cfg:
            eu, et = eps.chunk(2)
            eps = eu + guidance_scale * (et - eu)
        b, c, f, h, w = latents.shape
        flat = latents.permute(0, 2, 1, 3, 4).reshape(b * f, c, h, w)
        eflat = eps.permute(0, 2, 1, 3, 4).reshape(b * f, c, h, w)
        flat = scheduler.step(eflat, t, flat).prev_sample
        latents = flat.reshape(b, f, c, h, w).permute(0, 2, 1, 3, 4)
    if vae is None:
        return None, latents
    video = oracle_decode_latents(vae, latents)
    if output_type != "pt":
        video = tensor2vid(video)
    return video, latents


# ---------------------------------------------------------------------------------------------- shared test helper
def fill_deterministic(model: nn.Module, seed: int = 0, scale: float = 0.05):
    """Key-order-independent deterministic weights so that the verbatim reference model, this oracle and the product
    model (all sharing state_dict keys) hold identical parameters without shipping checkpoints.  Zero-initialised layers
    are re-drawn too so that they are exercised."""
    import hashlib
    sd = model.state_dict()
    new = {}
    for k, v in sd.items():
        if not torch.is_floating_point(v):
            new[k] = v
            continue
        h = int.from_bytes(hashlib.sha256(f"{seed}:{k}".encode()).digest()[:8], "little") % (2 ** 31)
        g = torch.Generator().manual_seed(h)
        t = torch.randn(v.shape, generator=g, dtype=torch.float32)
        if v.dim() == 1 and (("norm" in k and k.endswith("weight")) or ".0.weight" in k and v.dim() == 1):
            t = 1.0 + 0.1 * t                                     # norm gains
        elif v.dim() == 1:
            t = 0.1 * t                                           # biases
        else:
            fan_in = v[0].numel()
            t = t * (1.0 / fan_in) ** 0.5
        new[k] = t.to(v.dtype)
    model.load_state_dict(new)
    return model


# ---------------------------------------------------------------------------------------------- SVD path (config 4)
from diffusers._svd import (AutoencoderKLTemporalDecoder, EulerDiscreteScheduler,  # noqa: E402,F401
                            UNetSpatioTemporalConditionModel)

SVD_SCHED = dict(beta_start=0.00085, beta_end=0.012, beta_schedule="scaled_linear", prediction_type="v_prediction",
                 interpolation_type="linear", use_karras_sigmas=True, sigma_min=0.002, sigma_max=700.0,
                 timestep_spacing="leading", timestep_type="continuous", steps_offset=1)


@torch.no_grad()
def oracle_svd_sampling_loop(unet, scheduler, vae, image_embeddings, image_latents, mask, latents, num_inference_steps=25,
                             min_guidance_scale=1.0, max_guidance_scale=3.0, fps=6, motion_bucket_id=127,
                             noise_aug_strength=0.02, decode_chunk_size=None, decode=True, frame_mask=None,
                             condition_latent=None):
    """Restates the loop of MaskStableVideoDiffusionPipeline.__call__ (models/pipeline.py:375-459) from the point where the
    image has been encoded: `image_embeddings` [B, 1, D] (CLIP image embedding, positive half), `image_latents` [B, 4, h, w]
    (`vae.encode(image + noise).latent_dist.mode()`, NOT scaled), `mask` [1, h, w], `latents` [B, F, 4, h, w] unit
    noise.  Classifier-free guidance with zeroed negative conditioning (:343,_encode_vae_image), per-frame guidance
    scale linspace(min, max, F) (:405-408), 9-channel input cat([mask, latents, image_latents], dim=2) (:422), Euler
    step (:439), chunked temporal-VAE decode (:456).  Returns (frames [B, 3, F, H, W] fp32 or None, latents).
    TextStableVideoDiffusionPipeline.__call__ with condition_type="image" (models/pipeline.py:468-731, the call of
    app_svd.py:120-133) is the same loop with `frame_mask` [B, F, 1, h, w] used as given for both halves (:590) instead of `mask`,
    and, when `condition_latent` [B, F, 4, h, w] is passed, that tensor for BOTH halves (:602-604) instead of the image latents."""
    b, nf = latents.shape[:2]
    cfg = max_guidance_scale > 1.0
    emb = torch.cat([torch.zeros_like(image_embeddings), image_embeddings]) if cfg else image_embeddings
    il = torch.cat([torch.zeros_like(image_latents), image_latents]) if cfg else image_latents
    il = il.unsqueeze(1).repeat(1, nf, 1, 1, 1)
    if condition_latent is not None:
        il = torch.cat([condition_latent] * 2) if cfg else condition_latent
    if frame_mask is not None:
        m = torch.cat([frame_mask] * 2) if cfg else frame_mask
    else:
        m = mask[None, None].repeat(2, nf, 1, 1, 1).reshape(2, nf, 1, *mask.shape[-2:])      # '1 h w -> 2 f 1 h w'
    ids = torch.tensor([[fps, motion_bucket_id, noise_aug_strength]], dtype=emb.dtype).repeat(b, 1)
    ids = (torch.cat([ids, ids]) if cfg else ids).to(latents.device)
    scheduler.set_timesteps(num_inference_steps, device=latents.device)
    latents = latents * scheduler.init_noise_sigma
    gs = torch.linspace(min_guidance_scale, max_guidance_scale, nf).unsqueeze(0).to(latents.device, latents.dtype)
    gs = gs.repeat(b, 1)[:, :, None, None, None]
    for t in scheduler.timesteps:
        x = torch.cat([latents] * 2) if cfg else latents
        x = scheduler.scale_model_input(x, t)
        x = torch.cat([m.to(x.dtype), x, il], dim=2)
        pred = unet(x, t, encoder_hidden_states=emb, added_time_ids=ids, return_dict=False)[0]
        if cfg:
            pu, pc = pred.chunk(2)
            pred = pu + gs * (pc - pu)
        latents = scheduler.step(pred, t, latents).prev_sample
    if not decode:
        return None, latents
    chunk = decode_chunk_size or nf
    z = latents.flatten(0, 1) / vae.config.scaling_factor
    frames = torch.cat([vae.decode(z[i: i + chunk], num_frames=z[i: i + chunk].shape[0]).sample
                        for i in range(0, z.shape[0], chunk)], dim=0)
    frames = frames.reshape(-1, nf, *frames.shape[1:]).permute(0, 2, 1, 3, 4).float()
    return frames, latents


# ---------------------------------------------------------------------------------------------- transparent-video branch (row f4)
from diffusers._unet2d import (AttnDownBlock2D, AttnUpBlock2D, DownBlock2D, UpBlock2D,  # noqa: E402
                               UNetMidBlock2D as UNetMidBlock2D_024)


class OracleLatentTransparencyOffsetEncoder(nn.Module):
    """models/layerdiffuse_VAE.py:17-41: RGBA image [b, 4, H, W] -> latent offset [b, 4, H/8, W/8]; nine 3x3 convs
    (strides 1,1,2,1,2,1,2,1,1) with SiLU between them, none after the last."""

    def __init__(self):
        super().__init__()
        spec = [(4, 32, 1), (32, 32, 1), (32, 64, 2), (64, 64, 1), (64, 128, 2), (128, 128, 1), (128, 256, 2),
                (256, 256, 1), (256, 4, 1)]
        mods = []
        for i, (ci, co, st) in enumerate(spec):
            mods.append(nn.Conv2d(ci, co, 3, padding=1, stride=st))
            if i < len(spec) - 1:
                mods.append(nn.SiLU())
        self.blocks = nn.Sequential(*mods)

    def forward(self, x):
        return self.blocks(x)


class OracleUNet384(nn.Module):
    """models/layerdiffuse_VAE.py:44-177 (`UNet384`, the LayerDiffuse alpha decoder): a 2-D UNet without time embedding,
    GroupNorm(4), 3 DownBlock2D + 1 AttnDownBlock2D (head dim 8), mid block with attention, AttnUpBlock2D + 3 UpBlock2D;
    the SD latent enters through a 1x1 conv and is added before the last down block (:152-153, the 8x-downsampled level)."""

    def __init__(self, in_channels=3, out_channels=4, block_out_channels=(32, 64, 128, 256), layers_per_block=2,
                 attention_head_dim=8, norm_num_groups=4, norm_eps=1e-5):
        super().__init__()
        ch, g, e, hd = list(block_out_channels), norm_num_groups, norm_eps, attention_head_dim
        n = len(ch)
        self.conv_in = nn.Conv2d(in_channels, ch[0], 3, padding=1)
        self.latent_conv_in = nn.Conv2d(4, ch[2], 1)
        self.down_blocks = nn.ModuleList()
        prev = ch[0]
        for i in range(n):
            kw = dict(num_layers=layers_per_block, in_channels=prev, out_channels=ch[i], temb_channels=None, resnet_eps=e,
                      resnet_act_fn="silu", resnet_groups=g, downsample_padding=1)
            if i < n - 1:
                self.down_blocks.append(DownBlock2D(add_downsample=True, **kw))
            else:
                self.down_blocks.append(AttnDownBlock2D(attention_head_dim=hd, downsample_type=None, **kw))
            prev = ch[i]
        self.mid_block = UNetMidBlock2D_024(in_channels=ch[-1], temb_channels=None, resnet_eps=e, resnet_act_fn="silu",
                                            output_scale_factor=1, attention_head_dim=hd, resnet_groups=g)
        self.up_blocks = nn.ModuleList()
        rev = ch[::-1]
        out_c = rev[0]
        for i in range(n):
            prev_out, out_c = out_c, rev[i]
            in_c = rev[min(i + 1, n - 1)]
            kw = dict(num_layers=layers_per_block + 1, in_channels=in_c, out_channels=out_c, prev_output_channel=prev_out,
                      temb_channels=None, resnet_eps=e, resnet_act_fn="silu", resnet_groups=g)
            if i == 0:
                self.up_blocks.append(AttnUpBlock2D(attention_head_dim=hd, upsample_type="conv", **kw))
            else:
                self.up_blocks.append(UpBlock2D(add_upsample=i < n - 1, **kw))
        self.conv_norm_out = nn.GroupNorm(num_channels=ch[0], num_groups=g, eps=e)
        self.conv_act = nn.SiLU()
        self.conv_out = nn.Conv2d(ch[0], out_channels, 3, padding=1)

    def forward(self, x, latent):
        sample_latent = self.latent_conv_in(latent)
        sample = self.conv_in(x)
        skips = (sample,)
        for i, blk in enumerate(self.down_blocks):
            if i == 3:
                sample = sample + sample_latent
            sample, res = blk(hidden_states=sample, temb=None)
            skips += res
        sample = self.mid_block(sample, None)
        for blk in self.up_blocks:
            res = skips[-len(blk.resnets):]
            skips = skips[:-len(blk.resnets)]
            sample = blk(sample, res, None)
        return self.conv_out(self.conv_act(self.conv_norm_out(sample)))


def oracle_rgba_postprocess(decoded_rgba_bf, b, f):
    """models/pipeline_stage2.py:300-318: decoder output [(b f), 4, H, W] (model dtype) -> uint8 RGBA frames [f, H, W, 4]:
    alpha * 255 thresholded at 127 to {0, 255}; foreground (fg + 1) * 127.5; float -> clip(0, 255) -> truncate."""
    import numpy as np
    h, w = decoded_rgba_bf.shape[-2:]
    d = decoded_rgba_bf.reshape(b, f, 4, h, w).permute(0, 2, 1, 3, 4)
    alpha = d[:, 3:] * 255.0
    alpha[alpha > 127] = 255
    alpha[alpha <= 127] = 0
    fg = (d[:, :3] + 1.0) * 127.5
    pngs = torch.cat((fg, alpha), dim=1)[0].permute(1, 0, 2, 3).permute(0, 2, 3, 1)
    return pngs.detach().cpu().float().numpy().clip(0, 255).astype(np.uint8)


@torch.no_grad()
def oracle_masked_sampling_loop(unet, scheduler, vae, vae_alpha_decoder, latents, prompt_embeds, negative_prompt_embeds,
                                condition_latent, mask, motion, guidance_scale=9.0, num_inference_steps=50, timesteps=None):
    """Restates MaskedLatentToVideoPipeline.__call__ (models/pipeline_stage2.py:171-337) as train_transparent_i2v_stage2.py:500-515
    calls it (`image_embeds=None`): the denoising loop is LatentToVideoPipeline's (:252-296 equals models/pipeline.py:156-197 once
    `encode_prompt`'s tuple is re-concatenated [negative, positive] at :232); then decode_latents (:299), the alpha decoder on the
    decoded frames + final latents (:305-309) and the RGBA post-processing (:311-324).
    Returns (video fp32 [b,3,f,H,W], latents, pngs uint8 [f,H,W,4])."""
    video, latents = oracle_sampling_loop(unet, scheduler, latents, prompt_embeds, negative_prompt_embeds, condition_latent,
                                          mask, motion, guidance_scale=guidance_scale,
                                          num_inference_steps=num_inference_steps, timesteps=timesteps, vae=vae,
                                          output_type="pt")
    b, c, f, h, w = video.shape
    dtype = latents.dtype
    x = video.permute(0, 2, 1, 3, 4).reshape(b * f, c, h, w).to(dtype)
    lat = latents.permute(0, 2, 1, 3, 4).reshape(b * f, 4, *latents.shape[-2:])
    rgba = vae_alpha_decoder(x, lat)
    return video, latents, oracle_rgba_postprocess(rgba, b, f)
