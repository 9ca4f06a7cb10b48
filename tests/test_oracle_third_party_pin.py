"""CPU: a second, independent pin of oracle leaf semantics (the diffusers shim is "parity unpinned": no diffusers source or
wheel exists in this image).  `transformers` (installed, third party) ships its own implementation of the LDM / taming
autoencoder — the architecture the SD VAE *is* (diffusers' `AutoencoderKL` loads LDM checkpoints through a pure key renaming,
`convert_ldm_vae_checkpoint`) — as `JanusVQVAEEncoder` / `JanusVQVAEDecoder`: ResnetBlock (GroupNorm(32, eps 1e-6) -> swish ->
conv -> GroupNorm -> swish -> conv + (nin_)shortcut), AttnBlock (single head, softmax(Q K^T C^-0.5) V, 1x1 projections,
residual), Downsample (F.pad(0,1,0,1) + stride-2 conv, padding 0), Upsample (nearest x2 + conv), mid block, norm_out -> swish ->
conv_out.  The test copies one set of random weights into both (the LDM <-> diffusers key map) and requires identical outputs
from the shim's `Encoder` / `Decoder`, i.e. from the shim's ResnetBlock2D (temb=None), Downsample2D(padding=0), Upsample2D,
Attention (deprecated attn-block form) and UNetMidBlock2D, and from their composition.

Janus puts attention blocks on the lowest-resolution level, the SD VAE has none there: their output projections are zeroed, which
makes them exact identities (residual + 0).  What this does NOT pin: the time-embedding path of ResnetBlock2D, the transformer
blocks (GEGLU order, norms), TemporalConvLayer, the schedulers — those stay recalled (oracle/shim/diffusers/_impl.py header)."""
import os
import sys

import pytest
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

from oracle.composition import AutoencoderKL, fill_deterministic  # noqa: E402

janus = pytest.importorskip("transformers.models.janus.modeling_janus")
from transformers.models.janus.configuration_janus import JanusVQVAEConfig  # noqa: E402

VAE = dict(block_out_channels=(32, 32, 64, 64), layers_per_block=2, norm_num_groups=32, sample_size=64)


def _copy(dst, src):
    with torch.no_grad():
        dst.weight.copy_(src.weight.reshape(dst.weight.shape))
        dst.bias.copy_(src.bias)


def _copy_resnet(j, d):
    for a in ("norm1", "conv1", "norm2", "conv2"):
        _copy(getattr(j, a), getattr(d, a))
    if d.conv_shortcut is not None:
        _copy(j.nin_shortcut, d.conv_shortcut)
    else:
        assert not hasattr(j, "nin_shortcut")


def _copy_mid(j, d):
    _copy_resnet(j.block_1, d.resnets[0])
    _copy_resnet(j.block_2, d.resnets[1])
    at = d.attentions[0]
    _copy(j.attn_1.norm, at.group_norm)
    for a, b in (("q", at.to_q), ("k", at.to_k), ("v", at.to_v), ("proj_out", at.to_out[0])):
        _copy(getattr(j.attn_1, a), b)                       # Linear [C, C] -> Conv2d 1x1 [C, C, 1, 1]


def _identity_attn(attn_list):
    for a in attn_list:
        torch.nn.init.zeros_(a.proj_out.weight)
        torch.nn.init.zeros_(a.proj_out.bias)


def _models():
    vae = fill_deterministic(AutoencoderKL(**VAE).eval(), seed=3)
    cfg = JanusVQVAEConfig(double_latent=True, latent_channels=4, in_channels=3, out_channels=3, base_channels=32,
                           channel_multiplier=[1, 1, 2, 2], num_res_blocks=2, dropout=0.0)
    return vae, janus.JanusVQVAEEncoder(cfg).eval(), janus.JanusVQVAEDecoder(cfg).eval()


def test_shim_vae_encoder_equals_transformers_ldm_encoder():
    vae, jenc, _ = _models()
    enc = vae.encoder
    _copy(jenc.conv_in, enc.conv_in)
    for i, blk in enumerate(enc.down_blocks):
        for k, r in enumerate(blk.resnets):
            _copy_resnet(jenc.down[i].block[k], r)
        if blk.downsamplers is not None:
            _copy(jenc.down[i].downsample.conv, blk.downsamplers[0].conv)
        else:
            assert not hasattr(jenc.down[i], "downsample")
        _identity_attn(jenc.down[i].attn)
    _copy_mid(jenc.mid, enc.mid_block)
    _copy(jenc.norm_out, enc.conv_norm_out)
    _copy(jenc.conv_out, enc.conv_out)
    x = torch.randn(2, 3, 64, 96, generator=torch.Generator().manual_seed(1)).clamp(-1, 1)
    with torch.no_grad():
        ours = enc(x)
        theirs = jenc(x.clone())
    assert ours.shape == theirs.shape == (2, 8, 8, 12)
    assert torch.allclose(ours, theirs, rtol=1e-5, atol=1e-5), float((ours - theirs).abs().max())


def test_shim_vae_decoder_equals_transformers_ldm_decoder():
    vae, _, jdec = _models()
    dec = vae.decoder
    _copy(jdec.conv_in, dec.conv_in)
    _copy_mid(jdec.mid, dec.mid_block)
    for i, blk in enumerate(dec.up_blocks):                       # Janus builds `up` lowest resolution first, like diffusers
        for k, r in enumerate(blk.resnets):
            _copy_resnet(jdec.up[i].block[k], r)
        if blk.upsamplers is not None:
            _copy(jdec.up[i].upsample.conv, blk.upsamplers[0].conv)
        else:
            assert not hasattr(jdec.up[i], "upsample")
        _identity_attn(jdec.up[i].attn)
    _copy(jdec.norm_out, dec.conv_norm_out)
    _copy(jdec.conv_out, dec.conv_out)
    z = torch.randn(2, 4, 8, 12, generator=torch.Generator().manual_seed(2))
    with torch.no_grad():
        ours = dec(z)
        theirs = jdec(z.clone())
    assert ours.shape == theirs.shape == (2, 3, 64, 96)
    assert torch.allclose(ours, theirs, rtol=1e-5, atol=1e-5), float((ours - theirs).abs().max())


def test_attention_block_is_live_in_the_comparison():
    """Guard against a vacuous pass: with the mid-block attention's output projection zeroed on one side only, the outputs differ."""
    vae, jenc, _ = _models()
    enc = vae.encoder
    x = torch.randn(1, 3, 32, 32, generator=torch.Generator().manual_seed(5))
    with torch.no_grad():
        a = enc(x)
        torch.nn.init.zeros_(enc.mid_block.attentions[0].to_out[0].weight)
        b = enc(x)
    assert (a - b).abs().max() > 1e-4
