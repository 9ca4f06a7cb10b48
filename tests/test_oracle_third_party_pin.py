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


# ---------------------------------------------------------------------------------------------- the reference's own key map
def test_reference_key_map_fixes_resnet_and_head_layer_order():
    """The reference ships one artefact that states the leaf STRUCTURE independently of diffusers: the ModelScope <-> diffusers key
    map of utils/convert_diffusers_to_original_ms_text_to_video.py.  Its index pairs (:22-25 `time_embed.0 / .2` = linear_1 /
    linear_2; :34-37 `out.0 / .2` = conv_norm_out / conv_out; :44-49 `in_layers.0 / .2` = norm1 / conv1, `emb_layers.1` =
    time_emb_proj, `out_layers.0 / .3` = norm2 / conv2, `skip_connection` = conv_shortcut) are the Sequential positions of the
    original (guided-diffusion style) blocks, so the op order between the parametrised layers is fixed: GroupNorm, SiLU, conv |
    SiLU, Linear | GroupNorm, SiLU, Dropout, conv.  Blocks built from exactly those positions and loaded through exactly those
    pairs must equal the shim's ResnetBlock2D (time-embedding path included), TimestepEmbedding and output head."""
    import torch.nn as nn
    from diffusers._impl import ResnetBlock2D, TimestepEmbedding

    res_map = [("in_layers.0", "norm1"), ("in_layers.2", "conv1"), ("out_layers.0", "norm2"), ("out_layers.3", "conv2"),
               ("emb_layers.1", "time_emb_proj"), ("skip_connection", "conv_shortcut")]           # :44-49, verbatim pairs

    class OriginalResBlock(nn.Module):
        def __init__(self, cin, cout, temb, eps):
            super().__init__()
            self.in_layers = nn.Sequential(nn.GroupNorm(32, cin, eps=eps), nn.SiLU(), nn.Conv2d(cin, cout, 3, padding=1))
            self.emb_layers = nn.Sequential(nn.SiLU(), nn.Linear(temb, cout))
            self.out_layers = nn.Sequential(nn.GroupNorm(32, cout, eps=eps), nn.SiLU(), nn.Dropout(0.0),
                                            nn.Conv2d(cout, cout, 3, padding=1))
            self.skip_connection = nn.Conv2d(cin, cout, 1) if cin != cout else nn.Identity()

        def forward(self, x, emb):
            h = self.in_layers(x)
            h = h + self.emb_layers(emb)[:, :, None, None]
            return self.skip_connection(x) + self.out_layers(h)

    g = torch.Generator().manual_seed(0)
    for cin, cout in ((64, 128), (64, 64)):
        shim = fill_deterministic(ResnetBlock2D(in_channels=cin, out_channels=cout, temb_channels=96, eps=1e-5).eval(), seed=cin + cout)
        orig = OriginalResBlock(cin, cout, 96, 1e-5).eval()
        sd = {}
        for k, v in shim.state_dict().items():
            for o_name, hf_name in res_map:
                if k.startswith(hf_name + "."):
                    sd[o_name + k[len(hf_name):]] = v
        assert len(sd) == len(shim.state_dict())
        orig.load_state_dict(sd, strict=True)
        x, emb = torch.randn(2, cin, 8, 8, generator=g), torch.randn(2, 96, generator=g)
        with torch.no_grad():
            assert torch.allclose(shim(x, emb), orig(x, emb), rtol=1e-5, atol=1e-5)
    # time_embed = Sequential(Linear, SiLU, Linear) (:22-25): no activation after linear_2
    te = fill_deterministic(TimestepEmbedding(32, 64, act_fn="silu").eval(), seed=5)
    seq = nn.Sequential(nn.Linear(32, 64), nn.SiLU(), nn.Linear(64, 64)).eval()
    seq.load_state_dict({"0.weight": te.linear_1.weight, "0.bias": te.linear_1.bias, "2.weight": te.linear_2.weight,
                         "2.bias": te.linear_2.bias})
    t = torch.randn(3, 32, generator=g)
    with torch.no_grad():
        assert torch.allclose(te(t), seq(t), rtol=1e-6, atol=1e-6)
    # out = Sequential(GroupNorm, SiLU, conv) (:34-37) is how both the oracle and the mirror end the UNet (conv_norm_out, conv_act, conv_out)
    from oracle.composition import OracleUNet3D
    u = OracleUNet3D(block_out_channels=(32, 32, 32, 32), attention_head_dim=8, cross_attention_dim=16)
    assert isinstance(u.conv_norm_out, nn.GroupNorm) and isinstance(u.conv_act, nn.SiLU) and isinstance(u.conv_out, nn.Conv2d)
