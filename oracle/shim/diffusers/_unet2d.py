"""ORACLE (test infrastructure, not product): restatement of the `diffusers==0.24.0` `models/unet_2d_blocks.py`
classes that the reference's transparent-video branch imports (`models/layerdiffuse_VAE.py:7`:
`UNetMidBlock2D, get_down_block, get_up_block`, used with "DownBlock2D" / "AttnDownBlock2D" / "UpBlock2D" /
"AttnUpBlock2D", `temb_channels=None`, `resnet_time_scale_shift="default"`, conv resampling).

PARITY UNPINNED at this leaf level, like the rest of the shim (`_impl.py` header): recalled from the published 0.24.0
source, not diffable here.  Recalled facts this file fixes (one assertion each in tests/test_oracle_golden.py):
  * every block returns / consumes `output_states` in creation order; a down block appends the downsampler output too;
  * `Attn*Block2D` attentions are `Attention(C, heads=C // attention_head_dim, dim_head=attention_head_dim,
    norm_num_groups=resnet_groups, eps=resnet_eps, residual_connection=True, bias=True, upcast_softmax=True,
    rescale_output_factor=output_scale_factor, _from_deprecated_attn_block=True)`, run after each resnet;
  * `get_down_block("AttnDownBlock2D", add_downsample=False)` -> `downsample_type=None` (no downsampler); with
    add_downsample it defaults to "conv"; same for `get_up_block("AttnUpBlock2D")`;
  * `UpBlock2D` / `AttnUpBlock2D` resnet i takes `prev_output_channel if i == 0 else out_channels` plus the skip
    (`in_channels if i == num_layers - 1 else out_channels`) and concatenates `[hidden, skip]` in that order;
  * `UNetMidBlock2D(attn_groups=None, resnet_time_scale_shift="default")` -> attention GroupNorm uses `resnet_groups`.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from ._impl import Attention, Downsample2D, ResnetBlock2D, Upsample2D
from ._impl import UNetMidBlock2D as _MidBase


def _resnet(cin, cout, temb_channels, eps, groups, act, scale, dropout):
    return ResnetBlock2D(in_channels=cin, out_channels=cout, temb_channels=temb_channels, eps=eps, groups=groups,
                         dropout=dropout, non_linearity=act, output_scale_factor=scale)


def _attn(ch, head_dim, eps, groups, scale):
    return Attention(ch, heads=ch // head_dim, dim_head=head_dim, rescale_output_factor=scale, eps=eps,
                     norm_num_groups=groups, residual_connection=True, bias=True, upcast_softmax=True,
                     _from_deprecated_attn_block=True)


class DownBlock2D(nn.Module):
    def __init__(self, in_channels, out_channels, temb_channels, dropout=0.0, num_layers=1, resnet_eps=1e-6,
                 resnet_time_scale_shift="default", resnet_act_fn="swish", resnet_groups=32, resnet_pre_norm=True,
                 output_scale_factor=1.0, add_downsample=True, downsample_padding=1):
        super().__init__()
        assert resnet_time_scale_shift == "default"
        self.resnets = nn.ModuleList([_resnet(in_channels if i == 0 else out_channels, out_channels, temb_channels,
                                              resnet_eps, resnet_groups, resnet_act_fn, output_scale_factor, dropout)
                                      for i in range(num_layers)])
        self.downsamplers = (nn.ModuleList([Downsample2D(out_channels, use_conv=True, out_channels=out_channels,
                                                         padding=downsample_padding, name="op")])
                             if add_downsample else None)

    def forward(self, hidden_states, temb=None, scale: float = 1.0):
        output_states = ()
        for resnet in self.resnets:
            hidden_states = resnet(hidden_states, temb)
            output_states = output_states + (hidden_states,)
        if self.downsamplers is not None:
            for d in self.downsamplers:
                hidden_states = d(hidden_states)
            output_states = output_states + (hidden_states,)
        return hidden_states, output_states


class AttnDownBlock2D(nn.Module):
    def __init__(self, in_channels, out_channels, temb_channels, dropout=0.0, num_layers=1, resnet_eps=1e-6,
                 resnet_time_scale_shift="default", resnet_act_fn="swish", resnet_groups=32, resnet_pre_norm=True,
                 attention_head_dim=1, output_scale_factor=1.0, downsample_padding=1, downsample_type="conv"):
        super().__init__()
        assert resnet_time_scale_shift == "default"
        self.downsample_type = downsample_type
        if attention_head_dim is None:
            attention_head_dim = out_channels
        self.resnets = nn.ModuleList([_resnet(in_channels if i == 0 else out_channels, out_channels, temb_channels,
                                              resnet_eps, resnet_groups, resnet_act_fn, output_scale_factor, dropout)
                                      for i in range(num_layers)])
        self.attentions = nn.ModuleList([_attn(out_channels, attention_head_dim, resnet_eps, resnet_groups,
                                               output_scale_factor) for _ in range(num_layers)])
        if downsample_type == "conv":
            self.downsamplers = nn.ModuleList([Downsample2D(out_channels, use_conv=True, out_channels=out_channels,
                                                            padding=downsample_padding, name="op")])
        elif downsample_type is None:
            self.downsamplers = None
        else:
            raise NotImplementedError("oracle: downsample_type 'resnet' is not used by the reference")

    def forward(self, hidden_states, temb=None, upsample_size=None, cross_attention_kwargs=None):
        output_states = ()
        for resnet, attn in zip(self.resnets, self.attentions):
            hidden_states = resnet(hidden_states, temb)
            hidden_states = attn(hidden_states)
            output_states = output_states + (hidden_states,)
        if self.downsamplers is not None:
            for d in self.downsamplers:
                hidden_states = d(hidden_states)
            output_states += (hidden_states,)
        return hidden_states, output_states


class UpBlock2D(nn.Module):
    def __init__(self, in_channels, prev_output_channel, out_channels, temb_channels, resolution_idx=None, dropout=0.0,
                 num_layers=1, resnet_eps=1e-6, resnet_time_scale_shift="default", resnet_act_fn="swish",
                 resnet_groups=32, resnet_pre_norm=True, output_scale_factor=1.0, add_upsample=True):
        super().__init__()
        assert resnet_time_scale_shift == "default"
        resnets = []
        for i in range(num_layers):
            res_skip_channels = in_channels if (i == num_layers - 1) else out_channels
            resnet_in_channels = prev_output_channel if i == 0 else out_channels
            resnets.append(_resnet(resnet_in_channels + res_skip_channels, out_channels, temb_channels, resnet_eps,
                                   resnet_groups, resnet_act_fn, output_scale_factor, dropout))
        self.resnets = nn.ModuleList(resnets)
        self.upsamplers = (nn.ModuleList([Upsample2D(out_channels, use_conv=True, out_channels=out_channels)])
                           if add_upsample else None)

    def forward(self, hidden_states, res_hidden_states_tuple, temb=None, upsample_size=None, scale: float = 1.0):
        for resnet in self.resnets:
            res_hidden_states = res_hidden_states_tuple[-1]
            res_hidden_states_tuple = res_hidden_states_tuple[:-1]
            hidden_states = torch.cat([hidden_states, res_hidden_states], dim=1)
            hidden_states = resnet(hidden_states, temb)
        if self.upsamplers is not None:
            for u in self.upsamplers:
                hidden_states = u(hidden_states, upsample_size)
        return hidden_states


class AttnUpBlock2D(nn.Module):
    def __init__(self, in_channels, prev_output_channel, out_channels, temb_channels, resolution_idx=None, dropout=0.0,
                 num_layers=1, resnet_eps=1e-6, resnet_time_scale_shift="default", resnet_act_fn="swish",
                 resnet_groups=32, resnet_pre_norm=True, attention_head_dim=1, output_scale_factor=1.0,
                 upsample_type="conv"):
        super().__init__()
        assert resnet_time_scale_shift == "default"
        self.upsample_type = upsample_type
        if attention_head_dim is None:
            attention_head_dim = out_channels
        resnets, attentions = [], []
        for i in range(num_layers):
            res_skip_channels = in_channels if (i == num_layers - 1) else out_channels
            resnet_in_channels = prev_output_channel if i == 0 else out_channels
            resnets.append(_resnet(resnet_in_channels + res_skip_channels, out_channels, temb_channels, resnet_eps,
                                   resnet_groups, resnet_act_fn, output_scale_factor, dropout))
            attentions.append(_attn(out_channels, attention_head_dim, resnet_eps, resnet_groups, output_scale_factor))
        self.resnets = nn.ModuleList(resnets)
        self.attentions = nn.ModuleList(attentions)
        if upsample_type == "conv":
            self.upsamplers = nn.ModuleList([Upsample2D(out_channels, use_conv=True, out_channels=out_channels)])
        elif upsample_type is None:
            self.upsamplers = None
        else:
            raise NotImplementedError("oracle: upsample_type 'resnet' is not used by the reference")

    def forward(self, hidden_states, res_hidden_states_tuple, temb=None, upsample_size=None, scale: float = 1.0):
        for resnet, attn in zip(self.resnets, self.attentions):
            res_hidden_states = res_hidden_states_tuple[-1]
            res_hidden_states_tuple = res_hidden_states_tuple[:-1]
            hidden_states = torch.cat([hidden_states, res_hidden_states], dim=1)
            hidden_states = resnet(hidden_states, temb)
            hidden_states = attn(hidden_states)
        if self.upsamplers is not None:
            for u in self.upsamplers:
                hidden_states = u(hidden_states)
        return hidden_states


class UNetMidBlock2D(_MidBase):
    """0.24 signature (`dropout`, `resnet_time_scale_shift`, `attn_groups`, `resnet_pre_norm`) over the restatement in
    `_impl.py` (which the VAE uses with the older positional subset)."""

    def __init__(self, in_channels, temb_channels=None, dropout=0.0, num_layers=1, resnet_eps=1e-6,
                 resnet_time_scale_shift="default", resnet_act_fn="swish", resnet_groups=32, attn_groups=None,
                 resnet_pre_norm=True, add_attention=True, attention_head_dim=1, output_scale_factor=1.0):
        assert resnet_time_scale_shift == "default" and attn_groups in (None, resnet_groups)
        resnet_groups = resnet_groups if resnet_groups is not None else min(in_channels // 4, 32)
        super().__init__(in_channels, temb_channels=temb_channels, num_layers=num_layers, resnet_eps=resnet_eps,
                         resnet_act_fn=resnet_act_fn, resnet_groups=resnet_groups,
                         attention_head_dim=attention_head_dim, output_scale_factor=output_scale_factor,
                         add_attention=add_attention)


def get_down_block(down_block_type, num_layers, in_channels, out_channels, temb_channels, add_downsample, resnet_eps,
                   resnet_act_fn, transformer_layers_per_block=1, num_attention_heads=None, resnet_groups=None,
                   cross_attention_dim=None, downsample_padding=None, dual_cross_attention=False,
                   use_linear_projection=False, only_cross_attention=False, upcast_attention=False,
                   resnet_time_scale_shift="default", attention_type="default", resnet_skip_time_act=False,
                   resnet_out_scale_factor=1.0, cross_attention_norm=None, attention_head_dim=None,
                   downsample_type=None, dropout=0.0):
    if attention_head_dim is None:
        attention_head_dim = num_attention_heads
    down_block_type = down_block_type[7:] if down_block_type.startswith("UNetRes") else down_block_type
    if down_block_type == "DownBlock2D":
        return DownBlock2D(num_layers=num_layers, in_channels=in_channels, out_channels=out_channels,
                           temb_channels=temb_channels, dropout=dropout, add_downsample=add_downsample,
                           resnet_eps=resnet_eps, resnet_act_fn=resnet_act_fn, resnet_groups=resnet_groups,
                           downsample_padding=downsample_padding, resnet_time_scale_shift=resnet_time_scale_shift)
    if down_block_type == "AttnDownBlock2D":
        if add_downsample is False:
            downsample_type = None
        else:
            downsample_type = downsample_type or "conv"
        return AttnDownBlock2D(num_layers=num_layers, in_channels=in_channels, out_channels=out_channels,
                               temb_channels=temb_channels, dropout=dropout, resnet_eps=resnet_eps,
                               resnet_act_fn=resnet_act_fn, resnet_groups=resnet_groups,
                               downsample_padding=downsample_padding, attention_head_dim=attention_head_dim,
                               resnet_time_scale_shift=resnet_time_scale_shift, downsample_type=downsample_type)
    raise ValueError(f"{down_block_type} does not exist.")


def get_up_block(up_block_type, num_layers, in_channels, out_channels, prev_output_channel, temb_channels,
                 add_upsample, resnet_eps, resnet_act_fn, resolution_idx=None, transformer_layers_per_block=1,
                 num_attention_heads=None, resnet_groups=None, cross_attention_dim=None, dual_cross_attention=False,
                 use_linear_projection=False, only_cross_attention=False, upcast_attention=False,
                 resnet_time_scale_shift="default", attention_type="default", resnet_skip_time_act=False,
                 resnet_out_scale_factor=1.0, cross_attention_norm=None, attention_head_dim=None, upsample_type=None,
                 dropout=0.0):
    if attention_head_dim is None:
        attention_head_dim = num_attention_heads
    up_block_type = up_block_type[7:] if up_block_type.startswith("UNetRes") else up_block_type
    if up_block_type == "UpBlock2D":
        return UpBlock2D(num_layers=num_layers, in_channels=in_channels, out_channels=out_channels,
                         prev_output_channel=prev_output_channel, temb_channels=temb_channels,
                         resolution_idx=resolution_idx, dropout=dropout, add_upsample=add_upsample,
                         resnet_eps=resnet_eps, resnet_act_fn=resnet_act_fn, resnet_groups=resnet_groups,
                         resnet_time_scale_shift=resnet_time_scale_shift)
    if up_block_type == "AttnUpBlock2D":
        if add_upsample is False:
            upsample_type = None
        else:
            upsample_type = upsample_type or "conv"
        return AttnUpBlock2D(num_layers=num_layers, in_channels=in_channels, out_channels=out_channels,
                             prev_output_channel=prev_output_channel, temb_channels=temb_channels,
                             resolution_idx=resolution_idx, dropout=dropout, resnet_eps=resnet_eps,
                             resnet_act_fn=resnet_act_fn, resnet_groups=resnet_groups,
                             attention_head_dim=attention_head_dim, resnet_time_scale_shift=resnet_time_scale_shift,
                             upsample_type=upsample_type)
    raise ValueError(f"{up_block_type} does not exist.")
