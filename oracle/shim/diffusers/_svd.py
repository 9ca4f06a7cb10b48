"""ORACLE (test infrastructure, not product) — SVD leg of the diffusers shim: a pure-PyTorch restatement, FROM MEMORY of
`diffusers==0.24.0` (requirements.txt:4 of the reference; not vendored, not installable here), of the classes that
`/root/reference/models/pipeline.py:223-466` (`MaskStableVideoDiffusionPipeline`) and `/root/reference/train_svd.py:726-826`
drive:

    UNetSpatioTemporalConditionModel   (diffusers/models/unet_spatio_temporal_condition.py, unet_3d_blocks.py)
    TransformerSpatioTemporalModel / TemporalBasicTransformerBlock        (transformer_temporal.py, attention.py)
    SpatioTemporalResBlock / TemporalResnetBlock / AlphaBlender           (resnet.py)
    AutoencoderKLTemporalDecoder / TemporalDecoder                        (autoencoder_kl_temporal_decoder.py)
    EulerDiscreteScheduler                                               (scheduling_euler_discrete.py)
    StableVideoDiffusionPipeline helpers                                  (pipeline_stable_video_diffusion.py)

PARITY UNPINNED at the leaf level (SURVEY.md 8c marks these "lowest confidence RECALLED"): it could not be diffed against
the real package.  The *composition* the reference owns (its `__call__`, the 9-channel input cat at :422, the per-frame
guidance vector at :405-408,436) is pinned by importing the reference file verbatim over this shim
(tests/golden/make_golden.py svd).
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Any, Optional

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

from ._impl import (Attention, BaseOutput, BasicTransformerBlock, ConfigMixin, DiagonalGaussianDistribution,
                    DiffusionPipeline, Downsample2D, Encoder, FeedForward, ModelMixin, ResnetBlock2D, SchedulerOutput,
                    TimestepEmbedding, Timesteps, Upsample2D, AutoencoderKLOutput, DecoderOutput, randn_tensor,
                    register_to_config, _make_betas)


# ------------------------------------------------------------------------------------------------ resnet.py
class AlphaBlender(nn.Module):
    """alpha * x_spatial + (1 - alpha) * x_temporal; "learned": alpha = sigmoid(mix_factor); "learned_with_images": 1 where
    image_only_indicator is set."""

    def __init__(self, alpha: float, merge_strategy: str = "learned_with_images", switch_spatial_to_temporal_mix=False):
        super().__init__()
        self.merge_strategy = merge_strategy
        self.switch_spatial_to_temporal_mix = switch_spatial_to_temporal_mix
        if merge_strategy == "fixed":
            self.register_buffer("mix_factor", torch.Tensor([alpha]))
        else:
            self.register_parameter("mix_factor", nn.Parameter(torch.Tensor([alpha])))

    def get_alpha(self, image_only_indicator, ndims):
        if self.merge_strategy == "fixed":
            alpha = self.mix_factor
        elif self.merge_strategy == "learned":
            alpha = torch.sigmoid(self.mix_factor)
        else:                                                            # learned_with_images
            alpha = torch.where(image_only_indicator.bool(), torch.ones(1, 1, device=image_only_indicator.device),
                                torch.sigmoid(self.mix_factor)[..., None])
            if ndims == 5:
                alpha = alpha[:, None, :, None, None]
            elif ndims == 3:
                alpha = alpha.reshape(-1)[:, None, None]
        return alpha

    def forward(self, x_spatial, x_temporal, image_only_indicator=None):
        alpha = self.get_alpha(image_only_indicator, x_spatial.ndim).to(x_spatial.dtype)
        if self.switch_spatial_to_temporal_mix:
            alpha = 1.0 - alpha
        return alpha * x_spatial + (1.0 - alpha) * x_temporal


class TemporalResnetBlock(nn.Module):
    def __init__(self, in_channels, out_channels=None, temb_channels=512, eps=1e-6):
        super().__init__()
        out_channels = in_channels if out_channels is None else out_channels
        self.in_channels, self.out_channels = in_channels, out_channels
        self.norm1 = nn.GroupNorm(32, in_channels, eps=eps, affine=True)
        self.conv1 = nn.Conv3d(in_channels, out_channels, (3, 1, 1), stride=1, padding=(1, 0, 0))
        self.time_emb_proj = nn.Linear(temb_channels, out_channels) if temb_channels is not None else None
        self.norm2 = nn.GroupNorm(32, out_channels, eps=eps, affine=True)
        self.dropout = nn.Dropout(0.0)
        self.conv2 = nn.Conv3d(out_channels, out_channels, (3, 1, 1), stride=1, padding=(1, 0, 0))
        self.nonlinearity = nn.SiLU()
        self.use_in_shortcut = in_channels != out_channels
        self.conv_shortcut = nn.Conv3d(in_channels, out_channels, 1) if self.use_in_shortcut else None

    def forward(self, input_tensor, temb):                               # [B, C, F, H, W], [B, F, Ct]
        h = self.conv1(self.nonlinearity(self.norm1(input_tensor)))
        if self.time_emb_proj is not None:
            t = self.time_emb_proj(self.nonlinearity(temb))[:, :, :, None, None].permute(0, 2, 1, 3, 4)
            h = h + t
        h = self.conv2(self.dropout(self.nonlinearity(self.norm2(h))))
        if self.conv_shortcut is not None:
            input_tensor = self.conv_shortcut(input_tensor)
        return input_tensor + h


class SpatioTemporalResBlock(nn.Module):
    def __init__(self, in_channels, out_channels=None, temb_channels=512, eps=1e-6, temporal_eps=None,
                 merge_factor=0.5, merge_strategy="learned_with_images", switch_spatial_to_temporal_mix=False):
        super().__init__()
        out_channels = out_channels if out_channels is not None else in_channels
        self.spatial_res_block = ResnetBlock2D(in_channels=in_channels, out_channels=out_channels,
                                               temb_channels=temb_channels, eps=eps)
        self.temporal_res_block = TemporalResnetBlock(out_channels, out_channels, temb_channels=temb_channels,
                                                      eps=temporal_eps if temporal_eps is not None else eps)
        self.time_mixer = AlphaBlender(alpha=merge_factor, merge_strategy=merge_strategy,
                                       switch_spatial_to_temporal_mix=switch_spatial_to_temporal_mix)

    def forward(self, hidden_states, temb=None, image_only_indicator=None):
        nf = image_only_indicator.shape[-1]
        hidden_states = self.spatial_res_block(hidden_states, temb)
        bf, c, hh, ww = hidden_states.shape
        b = bf // nf
        mix = hidden_states[None, :].reshape(b, nf, c, hh, ww).permute(0, 2, 1, 3, 4)
        hidden_states = hidden_states[None, :].reshape(b, nf, c, hh, ww).permute(0, 2, 1, 3, 4)
        if temb is not None:
            temb = temb.reshape(b, nf, -1)
        hidden_states = self.temporal_res_block(hidden_states, temb)
        hidden_states = self.time_mixer(x_spatial=mix, x_temporal=hidden_states, image_only_indicator=image_only_indicator)
        return hidden_states.permute(0, 2, 1, 3, 4).reshape(bf, c, hh, ww)


# ------------------------------------------------------------------------------------------------ attention.py
class TemporalBasicTransformerBlock(nn.Module):
    def __init__(self, dim, time_mix_inner_dim, num_attention_heads, attention_head_dim, cross_attention_dim=None):
        super().__init__()
        self.is_res = dim == time_mix_inner_dim
        self.norm_in = nn.LayerNorm(dim)
        self.ff_in = FeedForward(dim, dim_out=time_mix_inner_dim, activation_fn="geglu")
        self.norm1 = nn.LayerNorm(time_mix_inner_dim)
        self.attn1 = Attention(query_dim=time_mix_inner_dim, heads=num_attention_heads, dim_head=attention_head_dim,
                               cross_attention_dim=None)
        if cross_attention_dim is not None:
            self.norm2 = nn.LayerNorm(time_mix_inner_dim)
            self.attn2 = Attention(query_dim=time_mix_inner_dim, cross_attention_dim=cross_attention_dim,
                                   heads=num_attention_heads, dim_head=attention_head_dim)
        else:
            self.norm2, self.attn2 = None, None
        self.norm3 = nn.LayerNorm(time_mix_inner_dim)
        self.ff = FeedForward(time_mix_inner_dim, activation_fn="geglu")

    def forward(self, hidden_states, num_frames, encoder_hidden_states=None):
        bf, s, c = hidden_states.shape
        b = bf // num_frames
        h = hidden_states[None, :].reshape(b, num_frames, s, c).permute(0, 2, 1, 3).reshape(b * s, num_frames, c)
        residual = h
        h = self.ff_in(self.norm_in(h))
        if self.is_res:
            h = h + residual
        h = self.attn1(self.norm1(h), encoder_hidden_states=None) + h
        if self.attn2 is not None:
            h = self.attn2(self.norm2(h), encoder_hidden_states=encoder_hidden_states) + h
        ff = self.ff(self.norm3(h))
        h = ff + h if self.is_res else ff
        return h[None, :].reshape(b, s, num_frames, c).permute(0, 2, 1, 3).reshape(bf, s, c)


@dataclass
class TransformerTemporalModelOutput(BaseOutput):
    sample: Any


class TransformerSpatioTemporalModel(nn.Module):
    def __init__(self, num_attention_heads=16, attention_head_dim=88, in_channels=320, out_channels=None, num_layers=1,
                 cross_attention_dim=None):
        super().__init__()
        self.num_attention_heads, self.attention_head_dim = num_attention_heads, attention_head_dim
        inner = num_attention_heads * attention_head_dim
        self.inner_dim = inner
        self.in_channels = in_channels
        self.norm = nn.GroupNorm(32, in_channels, eps=1e-6)
        self.proj_in = nn.Linear(in_channels, inner)
        self.transformer_blocks = nn.ModuleList([
            BasicTransformerBlock(inner, num_attention_heads, attention_head_dim, cross_attention_dim=cross_attention_dim)
            for _ in range(num_layers)])
        self.temporal_transformer_blocks = nn.ModuleList([
            TemporalBasicTransformerBlock(inner, inner, num_attention_heads, attention_head_dim,
                                          cross_attention_dim=cross_attention_dim) for _ in range(num_layers)])
        self.time_pos_embed = TimestepEmbedding(in_channels, in_channels * 4, out_dim=in_channels)
        self.time_proj = Timesteps(in_channels, True, 0)
        self.time_mixer = AlphaBlender(alpha=0.5, merge_strategy="learned_with_images")
        self.out_channels = in_channels if out_channels is None else out_channels
        self.proj_out = nn.Linear(inner, in_channels)

    def forward(self, hidden_states, encoder_hidden_states=None, image_only_indicator=None, return_dict=True):
        bf, _, hh, ww = hidden_states.shape
        nf = image_only_indicator.shape[-1]
        b = bf // nf
        time_context = encoder_hidden_states
        first = time_context[None, :].reshape(b, nf, -1, time_context.shape[-1])[:, 0]
        # diffusers broadcasts to (h*w, batch, ...) and flattens in THAT order, while the temporal block flattens its
        # tokens as (batch, h*w): kept as is (bug-compatible; with one key per row the cross-attention is a per-row vector)
        time_context = first[None, :].broadcast_to(hh * ww, b, 1, time_context.shape[-1])
        time_context = time_context.reshape(hh * ww * b, 1, time_context.shape[-1])
        residual = hidden_states
        hidden_states = self.norm(hidden_states)
        inner = hidden_states.shape[1]
        hidden_states = hidden_states.permute(0, 2, 3, 1).reshape(bf, hh * ww, inner)
        hidden_states = self.proj_in(hidden_states)
        num_frames_emb = torch.arange(nf, device=hidden_states.device).repeat(b, 1).reshape(-1)
        t_emb = self.time_proj(num_frames_emb).to(dtype=hidden_states.dtype)
        emb = self.time_pos_embed(t_emb)[:, None, :]
        for block, tblock in zip(self.transformer_blocks, self.temporal_transformer_blocks):
            hidden_states = block(hidden_states, encoder_hidden_states=encoder_hidden_states)
            mix = hidden_states + emb
            mix = tblock(mix, num_frames=nf, encoder_hidden_states=time_context)
            hidden_states = self.time_mixer(x_spatial=hidden_states, x_temporal=mix,
                                            image_only_indicator=image_only_indicator)
        hidden_states = self.proj_out(hidden_states)
        hidden_states = hidden_states.reshape(bf, hh, ww, inner).permute(0, 3, 1, 2).contiguous()
        out = hidden_states + residual
        if not return_dict:
            return (out,)
        return TransformerTemporalModelOutput(sample=out)


# ------------------------------------------------------------------------------------------------ unet_3d_blocks.py
class DownBlockSpatioTemporal(nn.Module):
    def __init__(self, in_channels, out_channels, temb_channels, num_layers=1, add_downsample=True):
        super().__init__()
        self.resnets = nn.ModuleList([SpatioTemporalResBlock(in_channels if i == 0 else out_channels, out_channels,
                                                             temb_channels, eps=1e-5) for i in range(num_layers)])
        self.downsamplers = (nn.ModuleList([Downsample2D(out_channels, use_conv=True, out_channels=out_channels,
                                                         name="op")]) if add_downsample else None)

    def forward(self, hidden_states, temb=None, image_only_indicator=None):
        outs = ()
        for r in self.resnets:
            hidden_states = r(hidden_states, temb, image_only_indicator=image_only_indicator)
            outs += (hidden_states,)
        if self.downsamplers is not None:
            for d in self.downsamplers:
                hidden_states = d(hidden_states)
            outs += (hidden_states,)
        return hidden_states, outs


class CrossAttnDownBlockSpatioTemporal(nn.Module):
    has_cross_attention = True

    def __init__(self, in_channels, out_channels, temb_channels, num_layers=1, transformer_layers_per_block=1,
                 num_attention_heads=1, cross_attention_dim=1280, add_downsample=True):
        super().__init__()
        if isinstance(transformer_layers_per_block, int):
            transformer_layers_per_block = [transformer_layers_per_block] * num_layers
        self.resnets = nn.ModuleList([SpatioTemporalResBlock(in_channels if i == 0 else out_channels, out_channels,
                                                             temb_channels, eps=1e-6) for i in range(num_layers)])
        self.attentions = nn.ModuleList([
            TransformerSpatioTemporalModel(num_attention_heads, out_channels // num_attention_heads,
                                           in_channels=out_channels, num_layers=transformer_layers_per_block[i],
                                           cross_attention_dim=cross_attention_dim) for i in range(num_layers)])
        self.downsamplers = (nn.ModuleList([Downsample2D(out_channels, use_conv=True, out_channels=out_channels,
                                                         padding=1, name="op")]) if add_downsample else None)

    def forward(self, hidden_states, temb=None, encoder_hidden_states=None, image_only_indicator=None):
        outs = ()
        for r, a in zip(self.resnets, self.attentions):
            hidden_states = r(hidden_states, temb, image_only_indicator=image_only_indicator)
            hidden_states = a(hidden_states, encoder_hidden_states=encoder_hidden_states,
                              image_only_indicator=image_only_indicator, return_dict=False)[0]
            outs += (hidden_states,)
        if self.downsamplers is not None:
            for d in self.downsamplers:
                hidden_states = d(hidden_states)
            outs += (hidden_states,)
        return hidden_states, outs


class UNetMidBlockSpatioTemporal(nn.Module):
    has_cross_attention = True

    def __init__(self, in_channels, temb_channels, num_layers=1, transformer_layers_per_block=1, num_attention_heads=1,
                 cross_attention_dim=1280):
        super().__init__()
        if isinstance(transformer_layers_per_block, int):
            transformer_layers_per_block = [transformer_layers_per_block] * num_layers
        resnets = [SpatioTemporalResBlock(in_channels, in_channels, temb_channels, eps=1e-5)]
        attentions = []
        for i in range(num_layers):
            attentions.append(TransformerSpatioTemporalModel(num_attention_heads, in_channels // num_attention_heads,
                                                             in_channels=in_channels,
                                                             num_layers=transformer_layers_per_block[i],
                                                             cross_attention_dim=cross_attention_dim))
            resnets.append(SpatioTemporalResBlock(in_channels, in_channels, temb_channels, eps=1e-5))
        self.attentions = nn.ModuleList(attentions)
        self.resnets = nn.ModuleList(resnets)

    def forward(self, hidden_states, temb=None, encoder_hidden_states=None, image_only_indicator=None):
        hidden_states = self.resnets[0](hidden_states, temb, image_only_indicator=image_only_indicator)
        for a, r in zip(self.attentions, self.resnets[1:]):
            hidden_states = a(hidden_states, encoder_hidden_states=encoder_hidden_states,
                              image_only_indicator=image_only_indicator, return_dict=False)[0]
            hidden_states = r(hidden_states, temb, image_only_indicator=image_only_indicator)
        return hidden_states


class UpBlockSpatioTemporal(nn.Module):
    def __init__(self, in_channels, prev_output_channel, out_channels, temb_channels, resolution_idx=None, num_layers=1,
                 resnet_eps=1e-6, add_upsample=True):
        super().__init__()
        resnets = []
        for i in range(num_layers):
            skip = in_channels if (i == num_layers - 1) else out_channels
            cin = prev_output_channel if i == 0 else out_channels
            resnets.append(SpatioTemporalResBlock(cin + skip, out_channels, temb_channels, eps=resnet_eps))
        self.resnets = nn.ModuleList(resnets)
        self.upsamplers = (nn.ModuleList([Upsample2D(out_channels, use_conv=True, out_channels=out_channels)])
                           if add_upsample else None)

    def forward(self, hidden_states, res_hidden_states_tuple, temb=None, image_only_indicator=None):
        for r in self.resnets:
            res = res_hidden_states_tuple[-1]
            res_hidden_states_tuple = res_hidden_states_tuple[:-1]
            hidden_states = torch.cat([hidden_states, res], dim=1)
            hidden_states = r(hidden_states, temb, image_only_indicator=image_only_indicator)
        if self.upsamplers is not None:
            for u in self.upsamplers:
                hidden_states = u(hidden_states)
        return hidden_states


class CrossAttnUpBlockSpatioTemporal(nn.Module):
    has_cross_attention = True

    def __init__(self, in_channels, out_channels, prev_output_channel, temb_channels, resolution_idx=None, num_layers=1,
                 transformer_layers_per_block=1, resnet_eps=1e-6, num_attention_heads=1, cross_attention_dim=1280,
                 add_upsample=True):
        super().__init__()
        if isinstance(transformer_layers_per_block, int):
            transformer_layers_per_block = [transformer_layers_per_block] * num_layers
        resnets, attentions = [], []
        for i in range(num_layers):
            skip = in_channels if (i == num_layers - 1) else out_channels
            cin = prev_output_channel if i == 0 else out_channels
            resnets.append(SpatioTemporalResBlock(cin + skip, out_channels, temb_channels, eps=resnet_eps))
            attentions.append(TransformerSpatioTemporalModel(num_attention_heads, out_channels // num_attention_heads,
                                                             in_channels=out_channels,
                                                             num_layers=transformer_layers_per_block[i],
                                                             cross_attention_dim=cross_attention_dim))
        self.resnets = nn.ModuleList(resnets)
        self.attentions = nn.ModuleList(attentions)
        self.upsamplers = (nn.ModuleList([Upsample2D(out_channels, use_conv=True, out_channels=out_channels)])
                           if add_upsample else None)

    def forward(self, hidden_states, res_hidden_states_tuple, temb=None, encoder_hidden_states=None,
                image_only_indicator=None):
        for r, a in zip(self.resnets, self.attentions):
            res = res_hidden_states_tuple[-1]
            res_hidden_states_tuple = res_hidden_states_tuple[:-1]
            hidden_states = torch.cat([hidden_states, res], dim=1)
            hidden_states = r(hidden_states, temb, image_only_indicator=image_only_indicator)
            hidden_states = a(hidden_states, encoder_hidden_states=encoder_hidden_states,
                              image_only_indicator=image_only_indicator, return_dict=False)[0]
        if self.upsamplers is not None:
            for u in self.upsamplers:
                hidden_states = u(hidden_states)
        return hidden_states


# ------------------------------------------------------------------------------------------------ unet
@dataclass
class UNetSpatioTemporalConditionOutput(BaseOutput):
    sample: Any = None


class UNetSpatioTemporalConditionModel(ModelMixin, ConfigMixin):
    @register_to_config
    def __init__(self, sample_size=None, in_channels=8, out_channels=4,
                 down_block_types=("CrossAttnDownBlockSpatioTemporal", "CrossAttnDownBlockSpatioTemporal",
                                   "CrossAttnDownBlockSpatioTemporal", "DownBlockSpatioTemporal"),
                 up_block_types=("UpBlockSpatioTemporal", "CrossAttnUpBlockSpatioTemporal",
                                 "CrossAttnUpBlockSpatioTemporal", "CrossAttnUpBlockSpatioTemporal"),
                 block_out_channels=(320, 640, 1280, 1280), addition_time_embed_dim=256,
                 projection_class_embeddings_input_dim=768, layers_per_block=2, cross_attention_dim=1024,
                 transformer_layers_per_block=1, num_attention_heads=(5, 10, 10, 20), num_frames=25):
        super().__init__()
        self.sample_size = sample_size
        n = len(down_block_types)
        if len(up_block_types) != n or len(block_out_channels) != n:
            raise ValueError("down_block_types, up_block_types and block_out_channels must have the same length")
        if isinstance(num_attention_heads, int):
            num_attention_heads = (num_attention_heads,) * n
        if isinstance(cross_attention_dim, int):
            cross_attention_dim = (cross_attention_dim,) * n
        if isinstance(layers_per_block, int):
            layers_per_block = [layers_per_block] * n
        if isinstance(transformer_layers_per_block, int):
            transformer_layers_per_block = [transformer_layers_per_block] * n
        c0 = block_out_channels[0]
        self.conv_in = nn.Conv2d(in_channels, c0, kernel_size=3, padding=1)
        ted = c0 * 4
        self.time_proj = Timesteps(c0, True, downscale_freq_shift=0)
        self.time_embedding = TimestepEmbedding(c0, ted)
        self.add_time_proj = Timesteps(addition_time_embed_dim, True, downscale_freq_shift=0)
        self.add_embedding = TimestepEmbedding(projection_class_embeddings_input_dim, ted)
        self.down_blocks = nn.ModuleList()
        self.up_blocks = nn.ModuleList()
        out_c = c0
        for i, t in enumerate(down_block_types):
            in_c, out_c = out_c, block_out_channels[i]
            final = i == n - 1
            if t == "DownBlockSpatioTemporal":
                blk = DownBlockSpatioTemporal(in_c, out_c, ted, num_layers=layers_per_block[i], add_downsample=not final)
            elif t == "CrossAttnDownBlockSpatioTemporal":
                blk = CrossAttnDownBlockSpatioTemporal(in_c, out_c, ted, num_layers=layers_per_block[i],
                                                       transformer_layers_per_block=transformer_layers_per_block[i],
                                                       num_attention_heads=num_attention_heads[i],
                                                       cross_attention_dim=cross_attention_dim[i], add_downsample=not final)
            else:
                raise ValueError(f"{t} does not exist.")
            self.down_blocks.append(blk)
        self.mid_block = UNetMidBlockSpatioTemporal(block_out_channels[-1], ted,
                                                    transformer_layers_per_block=transformer_layers_per_block[-1],
                                                    cross_attention_dim=cross_attention_dim[-1],
                                                    num_attention_heads=num_attention_heads[-1])
        self.num_upsamplers = 0
        rev_c = list(reversed(block_out_channels))
        rev_h = list(reversed(num_attention_heads))
        rev_l = list(reversed(layers_per_block))
        rev_x = list(reversed(cross_attention_dim))
        rev_t = list(reversed(transformer_layers_per_block))
        out_c = rev_c[0]
        for i, t in enumerate(up_block_types):
            final = i == n - 1
            prev, out_c = out_c, rev_c[i]
            in_c = rev_c[min(i + 1, n - 1)]
            if not final:
                self.num_upsamplers += 1
            if t == "UpBlockSpatioTemporal":
                blk = UpBlockSpatioTemporal(in_c, prev, out_c, ted, resolution_idx=i, num_layers=rev_l[i] + 1,
                                            add_upsample=not final)
            elif t == "CrossAttnUpBlockSpatioTemporal":
                blk = CrossAttnUpBlockSpatioTemporal(in_c, out_c, prev, ted, resolution_idx=i, num_layers=rev_l[i] + 1,
                                                     transformer_layers_per_block=rev_t[i], num_attention_heads=rev_h[i],
                                                     cross_attention_dim=rev_x[i], add_upsample=not final)
            else:
                raise ValueError(f"{t} does not exist.")
            self.up_blocks.append(blk)
        self.conv_norm_out = nn.GroupNorm(num_channels=c0, num_groups=32, eps=1e-5)
        self.conv_act = nn.SiLU()
        self.conv_out = nn.Conv2d(c0, out_channels, kernel_size=3, padding=1)

    def forward(self, sample, timestep, encoder_hidden_states, added_time_ids, return_dict=True):
        timesteps = timestep
        if not torch.is_tensor(timesteps):
            dt = torch.float32 if isinstance(timestep, float) else torch.int64
            timesteps = torch.tensor([timesteps], dtype=dt, device=sample.device)
        elif timesteps.ndim == 0:
            timesteps = timesteps[None].to(sample.device)
        b, nf = sample.shape[:2]
        timesteps = timesteps.expand(b)
        t_emb = self.time_proj(timesteps).to(dtype=sample.dtype)
        emb = self.time_embedding(t_emb)
        time_embeds = self.add_time_proj(added_time_ids.flatten()).reshape((b, -1)).to(emb.dtype)
        emb = emb + self.add_embedding(time_embeds)
        sample = sample.flatten(0, 1)
        emb = emb.repeat_interleave(nf, dim=0)
        encoder_hidden_states = encoder_hidden_states.repeat_interleave(nf, dim=0)
        sample = self.conv_in(sample)
        ioi = torch.zeros(b, nf, dtype=sample.dtype, device=sample.device)
        skips = (sample,)
        for blk in self.down_blocks:
            if getattr(blk, "has_cross_attention", False):
                sample, res = blk(sample, temb=emb, encoder_hidden_states=encoder_hidden_states, image_only_indicator=ioi)
            else:
                sample, res = blk(sample, temb=emb, image_only_indicator=ioi)
            skips += res
        sample = self.mid_block(sample, temb=emb, encoder_hidden_states=encoder_hidden_states, image_only_indicator=ioi)
        for blk in self.up_blocks:
            res = skips[-len(blk.resnets):]
            skips = skips[: -len(blk.resnets)]
            if getattr(blk, "has_cross_attention", False):
                sample = blk(sample, res, temb=emb, encoder_hidden_states=encoder_hidden_states, image_only_indicator=ioi)
            else:
                sample = blk(sample, res, temb=emb, image_only_indicator=ioi)
        sample = self.conv_out(self.conv_act(self.conv_norm_out(sample)))
        sample = sample.reshape(b, nf, *sample.shape[1:])
        if not return_dict:
            return (sample,)
        return UNetSpatioTemporalConditionOutput(sample=sample)


# ------------------------------------------------------------------------------------------------ temporal VAE
class MidBlockTemporalDecoder(nn.Module):
    def __init__(self, in_channels, out_channels, attention_head_dim=512, num_layers=1, upcast_attention=False):
        super().__init__()
        self.resnets = nn.ModuleList([
            SpatioTemporalResBlock(in_channels if i == 0 else out_channels, out_channels, temb_channels=None, eps=1e-6,
                                   temporal_eps=1e-5, merge_factor=0.0, merge_strategy="learned",
                                   switch_spatial_to_temporal_mix=True) for i in range(num_layers)])
        self.attentions = nn.ModuleList([Attention(query_dim=in_channels, heads=in_channels // attention_head_dim,
                                                   dim_head=attention_head_dim, eps=1e-6,
                                                   upcast_attention=upcast_attention, norm_num_groups=32, bias=True,
                                                   residual_connection=True)])

    def forward(self, hidden_states, image_only_indicator):
        hidden_states = self.resnets[0](hidden_states, image_only_indicator=image_only_indicator)
        for r, a in zip(self.resnets[1:], self.attentions):
            hidden_states = a(hidden_states)
            hidden_states = r(hidden_states, image_only_indicator=image_only_indicator)
        return hidden_states


class UpBlockTemporalDecoder(nn.Module):
    def __init__(self, in_channels, out_channels, num_layers=1, add_upsample=True):
        super().__init__()
        self.resnets = nn.ModuleList([
            SpatioTemporalResBlock(in_channels if i == 0 else out_channels, out_channels, temb_channels=None, eps=1e-6,
                                   temporal_eps=1e-5, merge_factor=0.0, merge_strategy="learned",
                                   switch_spatial_to_temporal_mix=True) for i in range(num_layers)])
        self.upsamplers = (nn.ModuleList([Upsample2D(out_channels, use_conv=True, out_channels=out_channels)])
                           if add_upsample else None)

    def forward(self, hidden_states, image_only_indicator):
        for r in self.resnets:
            hidden_states = r(hidden_states, image_only_indicator=image_only_indicator)
        if self.upsamplers is not None:
            for u in self.upsamplers:
                hidden_states = u(hidden_states)
        return hidden_states


class TemporalDecoder(nn.Module):
    def __init__(self, in_channels=4, out_channels=3, block_out_channels=(128, 256, 512, 512), layers_per_block=2):
        super().__init__()
        self.layers_per_block = layers_per_block
        self.conv_in = nn.Conv2d(in_channels, block_out_channels[-1], kernel_size=3, stride=1, padding=1)
        self.mid_block = MidBlockTemporalDecoder(num_layers=layers_per_block, in_channels=block_out_channels[-1],
                                                 out_channels=block_out_channels[-1],
                                                 attention_head_dim=block_out_channels[-1])
        self.up_blocks = nn.ModuleList()
        rev = list(reversed(block_out_channels))
        out_c = rev[0]
        for i in range(len(block_out_channels)):
            prev, out_c = out_c, rev[i]
            self.up_blocks.append(UpBlockTemporalDecoder(prev, out_c, num_layers=layers_per_block + 1,
                                                         add_upsample=i != len(block_out_channels) - 1))
        self.conv_norm_out = nn.GroupNorm(num_channels=block_out_channels[0], num_groups=32, eps=1e-6)
        self.conv_act = nn.SiLU()
        self.conv_out = nn.Conv2d(block_out_channels[0], out_channels, kernel_size=3, padding=1)
        self.time_conv_out = nn.Conv3d(out_channels, out_channels, (3, 1, 1), padding=(1, 0, 0))

    def forward(self, sample, image_only_indicator, num_frames=1):
        sample = self.conv_in(sample)
        sample = self.mid_block(sample, image_only_indicator=image_only_indicator)
        for u in self.up_blocks:
            sample = u(sample, image_only_indicator=image_only_indicator)
        sample = self.conv_out(self.conv_act(self.conv_norm_out(sample)))
        bf, c, hh, ww = sample.shape
        b = bf // num_frames
        sample = sample[None, :].reshape(b, num_frames, c, hh, ww).permute(0, 2, 1, 3, 4)
        sample = self.time_conv_out(sample)
        return sample.permute(0, 2, 1, 3, 4).reshape(bf, c, hh, ww)


class AutoencoderKLTemporalDecoder(ModelMixin, ConfigMixin):
    @register_to_config
    def __init__(self, in_channels=3, out_channels=3, down_block_types=("DownEncoderBlock2D",) * 4,
                 block_out_channels=(128, 256, 512, 512), layers_per_block=2, latent_channels=4, sample_size=768,
                 scaling_factor=0.18215, force_upcast=True):
        super().__init__()
        self.encoder = Encoder(in_channels=in_channels, out_channels=latent_channels, down_block_types=down_block_types,
                               block_out_channels=block_out_channels, layers_per_block=layers_per_block, double_z=True)
        self.decoder = TemporalDecoder(in_channels=latent_channels, out_channels=out_channels,
                                       block_out_channels=block_out_channels, layers_per_block=layers_per_block)
        self.quant_conv = nn.Conv2d(2 * latent_channels, 2 * latent_channels, 1)

    def encode(self, x, return_dict=True):
        posterior = DiagonalGaussianDistribution(self.quant_conv(self.encoder(x)))
        if not return_dict:
            return (posterior,)
        return AutoencoderKLOutput(latent_dist=posterior)

    def decode(self, z, num_frames, return_dict=True):
        b = z.shape[0] // num_frames
        ioi = torch.zeros(b, num_frames, dtype=z.dtype, device=z.device)
        decoded = self.decoder(z, num_frames=num_frames, image_only_indicator=ioi)
        if not return_dict:
            return (decoded,)
        return DecoderOutput(sample=decoded)


# ------------------------------------------------------------------------------------------------ scheduler
class EulerDiscreteScheduler(ConfigMixin):
    order = 1

    @register_to_config
    def __init__(self, num_train_timesteps=1000, beta_start=0.0001, beta_end=0.02, beta_schedule="linear",
                 prediction_type="epsilon", interpolation_type="linear", use_karras_sigmas=False, sigma_min=None,
                 sigma_max=None, timestep_spacing="linspace", timestep_type="discrete", steps_offset=0):
        self.betas = _make_betas(num_train_timesteps, beta_start, beta_end, beta_schedule)
        self.alphas = 1.0 - self.betas
        self.alphas_cumprod = torch.cumprod(self.alphas, dim=0)
        sigmas = np.array(((1 - self.alphas_cumprod) / self.alphas_cumprod) ** 0.5)
        timesteps = np.linspace(0, num_train_timesteps - 1, num_train_timesteps, dtype=float)[::-1].copy()
        sigmas = torch.from_numpy(sigmas[::-1].copy()).to(dtype=torch.float32)
        self.num_inference_steps = None
        if timestep_type == "continuous" and prediction_type == "v_prediction":
            self.timesteps = torch.Tensor([0.25 * s.log() for s in sigmas])
        else:
            self.timesteps = torch.from_numpy(timesteps).to(dtype=torch.float32)
        self.sigmas = torch.cat([sigmas, torch.zeros(1)])
        self.is_scale_input_called = False
        self._step_index = None

    @property
    def init_noise_sigma(self):
        max_sigma = max(self.sigmas) if isinstance(self.sigmas, list) else self.sigmas.max()
        if self.config.timestep_spacing in ("linspace", "trailing"):
            return max_sigma
        return (max_sigma ** 2 + 1) ** 0.5

    @property
    def step_index(self):
        return self._step_index

    def _init_step_index(self, timestep):
        if isinstance(timestep, torch.Tensor):
            timestep = timestep.to(self.timesteps.device)
        idx = (self.timesteps == timestep).nonzero()
        self._step_index = (idx[1] if len(idx) > 1 else idx[0]).item()

    def scale_model_input(self, sample, timestep):
        if self._step_index is None:
            self._init_step_index(timestep)
        sigma = self.sigmas[self._step_index]
        self.is_scale_input_called = True
        return sample / ((sigma ** 2 + 1) ** 0.5)

    def _sigma_to_t(self, sigma, log_sigmas):
        log_sigma = np.log(np.maximum(sigma, 1e-10))
        dists = log_sigma - log_sigmas[:, np.newaxis]
        low_idx = np.cumsum((dists >= 0), axis=0).argmax(axis=0).clip(max=log_sigmas.shape[0] - 2)
        high_idx = low_idx + 1
        low, high = log_sigmas[low_idx], log_sigmas[high_idx]
        w = np.clip((low - log_sigma) / (low - high), 0, 1)
        return ((1 - w) * low_idx + w * high_idx).reshape(sigma.shape)

    def _convert_to_karras(self, in_sigmas, num_inference_steps):
        sigma_min = self.config.sigma_min if self.config.sigma_min is not None else in_sigmas[-1].item()
        sigma_max = self.config.sigma_max if self.config.sigma_max is not None else in_sigmas[0].item()
        rho = 7.0
        ramp = np.linspace(0, 1, num_inference_steps)
        min_inv_rho, max_inv_rho = sigma_min ** (1 / rho), sigma_max ** (1 / rho)
        return (max_inv_rho + ramp * (min_inv_rho - max_inv_rho)) ** rho

    def set_timesteps(self, num_inference_steps, device=None):
        self.num_inference_steps = num_inference_steps
        n = self.config.num_train_timesteps
        sp = self.config.timestep_spacing
        if sp == "linspace":
            timesteps = np.linspace(0, n - 1, num_inference_steps, dtype=np.float32)[::-1].copy()
        elif sp == "leading":
            step_ratio = n // num_inference_steps
            timesteps = (np.arange(0, num_inference_steps) * step_ratio).round()[::-1].copy().astype(np.float32)
            timesteps += self.config.steps_offset
        else:
            step_ratio = n / num_inference_steps
            timesteps = (np.arange(n, 0, -step_ratio)).round().copy().astype(np.float32) - 1
        sigmas = np.array(((1 - self.alphas_cumprod) / self.alphas_cumprod) ** 0.5)
        log_sigmas = np.log(sigmas)
        sigmas = np.interp(timesteps, np.arange(0, len(sigmas)), sigmas)
        if self.config.use_karras_sigmas:
            sigmas = self._convert_to_karras(in_sigmas=sigmas, num_inference_steps=num_inference_steps)
            timesteps = np.array([self._sigma_to_t(s, log_sigmas) for s in sigmas])
        sigmas = torch.from_numpy(sigmas).to(dtype=torch.float32, device=device)
        if self.config.timestep_type == "continuous" and self.config.prediction_type == "v_prediction":
            self.timesteps = torch.Tensor([0.25 * s.log() for s in sigmas]).to(device=device)
        else:
            self.timesteps = torch.from_numpy(np.asarray(timesteps, dtype=np.float32)).to(device=device)
        self.sigmas = torch.cat([sigmas, torch.zeros(1, device=sigmas.device)])
        self._step_index = None

    def step(self, model_output, timestep, sample, s_churn=0.0, s_tmin=0.0, s_tmax=float("inf"), s_noise=1.0,
             generator=None, return_dict=True):
        if self._step_index is None:
            self._init_step_index(timestep)
        sample = sample.to(torch.float32)
        sigma = self.sigmas[self._step_index]
        gamma = min(s_churn / (len(self.sigmas) - 1), 2 ** 0.5 - 1) if s_tmin <= sigma <= s_tmax else 0.0
        noise = randn_tensor(model_output.shape, dtype=model_output.dtype, device=model_output.device, generator=generator)
        eps = noise * s_noise
        sigma_hat = sigma * (gamma + 1)
        if gamma > 0:
            sample = sample + eps * (sigma_hat ** 2 - sigma ** 2) ** 0.5
        pt = self.config.prediction_type
        if pt in ("original_sample", "sample"):
            pred = model_output
        elif pt == "epsilon":
            pred = sample - sigma_hat * model_output
        elif pt == "v_prediction":
            pred = model_output * (-sigma / (sigma ** 2 + 1) ** 0.5) + (sample / (sigma ** 2 + 1))
        else:
            raise ValueError(pt)
        derivative = (sample - pred) / sigma_hat
        dt = self.sigmas[self._step_index + 1] - sigma_hat
        prev = (sample + derivative * dt).to(model_output.dtype)
        self._step_index += 1
        if not return_dict:
            return (prev,)
        return SchedulerOutput(prev_sample=prev, pred_original_sample=pred)


# ------------------------------------------------------------------------------------------------ pipeline
@dataclass
class StableVideoDiffusionPipelineOutput(BaseOutput):
    frames: Any = None


class _SvdImageProcessor:
    """The slice of VaeImageProcessor the SVD pipeline touches for tensor inputs."""

    def preprocess(self, image, height=None, width=None):
        if not torch.is_tensor(image):
            raise NotImplementedError("oracle SVD pipeline takes image tensors in [-1, 1] ([B, 3, H, W])")
        if image.shape[-2:] != (height, width):
            image = F.interpolate(image, size=(height, width), mode="bilinear", align_corners=False)
        return image

    def postprocess(self, image, output_type="pt"):
        image = (image / 2 + 0.5).clamp(0, 1)
        if output_type == "pt":
            return image
        return image.cpu().permute(0, 2, 3, 1).float().numpy()


def svd_tensor2vid(video, processor, output_type="np"):
    outputs = [processor.postprocess(video[b].permute(1, 0, 2, 3), output_type) for b in range(video.shape[0])]
    return outputs


class StableVideoDiffusionPipeline(DiffusionPipeline):
    def __init__(self, vae, image_encoder, unet, scheduler, feature_extractor=None):
        self.vae, self.image_encoder, self.unet, self.scheduler = vae, image_encoder, unet, scheduler
        self.feature_extractor = feature_extractor
        self.vae_scale_factor = 2 ** (len(self.vae.config.block_out_channels) - 1)
        self.image_processor = _SvdImageProcessor()

    @property
    def _execution_device(self):
        return next(self.unet.parameters()).device

    @property
    def guidance_scale(self):
        return self._guidance_scale

    def maybe_free_model_hooks(self):
        pass

    def check_inputs(self, image, height, width):
        if not torch.is_tensor(image):
            raise ValueError(f"`image` has to be a torch.Tensor in the oracle pipeline but is {type(image)}")
        if height % 8 != 0 or width % 8 != 0:
            raise ValueError(f"`height` and `width` have to be divisible by 8 but are {height} and {width}.")

    def _encode_image(self, image, device, num_videos_per_prompt, do_classifier_free_guidance):
        """CLIP vision tower + projection in diffusers (outside SURVEY 8's hot path).  Here `image_encoder` is any module
        mapping the [-1, 1] image batch to `[B, D]` embeddings (tests use a fixed random projection)."""
        dtype = next(self.image_encoder.parameters()).dtype
        emb = self.image_encoder(image.to(device=device, dtype=dtype))
        emb = emb.unsqueeze(1)
        bs, seq, _ = emb.shape
        emb = emb.repeat(1, num_videos_per_prompt, 1).view(bs * num_videos_per_prompt, seq, -1)
        if do_classifier_free_guidance:
            emb = torch.cat([torch.zeros_like(emb), emb])
        return emb

    def _encode_vae_image(self, image, device, num_videos_per_prompt, do_classifier_free_guidance):
        image = image.to(device=device)
        lat = self.vae.encode(image).latent_dist.mode()
        if do_classifier_free_guidance:
            lat = torch.cat([torch.zeros_like(lat), lat])
        return lat.repeat(num_videos_per_prompt, 1, 1, 1)

    def _get_add_time_ids(self, fps, motion_bucket_id, noise_aug_strength, dtype, batch_size, num_videos_per_prompt,
                          do_classifier_free_guidance):
        ids = torch.tensor([[fps, motion_bucket_id, noise_aug_strength]], dtype=dtype)
        ids = ids.repeat(batch_size * num_videos_per_prompt, 1)
        if do_classifier_free_guidance:
            ids = torch.cat([ids, ids])
        return ids

    def prepare_latents(self, batch_size, num_frames, num_channels_latents, height, width, dtype, device, generator,
                        latents=None):
        shape = (batch_size, num_frames, num_channels_latents // 2, height // self.vae_scale_factor,
                 width // self.vae_scale_factor)
        if latents is None:
            latents = randn_tensor(shape, generator=generator, device=device, dtype=dtype)
        else:
            latents = latents.to(device)
        return latents * self.scheduler.init_noise_sigma

    def decode_latents(self, latents, num_frames, decode_chunk_size=14):
        latents = latents.flatten(0, 1)
        latents = 1 / self.vae.config.scaling_factor * latents
        frames = []
        for i in range(0, latents.shape[0], decode_chunk_size):
            n_in = latents[i: i + decode_chunk_size].shape[0]
            frames.append(self.vae.decode(latents[i: i + decode_chunk_size], num_frames=n_in).sample)
        frames = torch.cat(frames, dim=0)
        frames = frames.reshape(-1, num_frames, *frames.shape[1:]).permute(0, 2, 1, 3, 4)
        return frames.float()
