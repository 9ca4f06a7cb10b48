"""Constructor behaviour of the UNet3DConditionModel mirror without a GPU: the three configuration checks of
models/unet_3d_condition_mask.py:118-131 raise ValueError in the same cases as the verbatim reference class (checked where /root/reference exists), the config object
exposes what callers read (train.py:91 `unet.config.in_channels`, models/pipeline.py:107 `unet.config.sample_size`),
`conv_in.weight/bias` are nn.Parameters (train.py:98-101), and unknown block types are rejected
(models/unet_3d_blocks.py:96,172)."""
import os
import sys

import pytest
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))

TINY = dict(sample_size=16, block_out_channels=(64, 64, 64, 64), attention_head_dim=64, cross_attention_dim=32,
            motion_mask=True, motion_strength=True)
BAD = [
    dict(TINY, up_block_types=("UpBlock3D", "CrossAttnUpBlock3D", "CrossAttnUpBlock3D")),
    dict(TINY, block_out_channels=(64, 64, 64)),
    dict(TINY, attention_head_dim=(64, 64)),
]


def _verbatim_reference_class():
    """The reference's own class over the diffusers shim -- only where /root/reference exists (the build container)."""
    if not os.path.isdir("/root/reference/models"):
        return None
    sys.path.insert(0, os.path.join(os.path.dirname(HERE), "oracle", "shim"))
    sys.path.insert(0, "/root/reference")
    from models.unet_3d_condition_mask import UNet3DConditionModel as Ref
    return Ref


@pytest.mark.parametrize("i", range(len(BAD)))
def test_inconsistent_configs_raise_like_the_reference(i):
    from animate_anything_b200.unet_3d_condition_mask import UNet3DConditionModel
    cfg = BAD[i]
    ref = _verbatim_reference_class()
    if ref is not None:
        with pytest.raises(ValueError):
            ref(**cfg)
    with pytest.raises(ValueError):
        UNet3DConditionModel(**cfg)


def test_config_surface_and_parameters():
    from animate_anything_b200.unet_3d_condition_mask import UNet3DConditionModel
    m = UNet3DConditionModel(**TINY)
    assert m.config.in_channels == 4 and m.config.sample_size == 16
    assert isinstance(m.conv_in.weight, torch.nn.Parameter) and isinstance(m.conv_in.bias, torch.nn.Parameter)
    assert m.conv_in2.weight.shape[1] == 5                         # 4 latent channels + 1 mask channel (:140-142)
    assert hasattr(m, "motion_proj") or hasattr(m, "motion_embedding")
    m.requires_grad_(False).eval()
    assert m.dtype == torch.float32
    assert any(n.endswith("attn1.to_q.weight") for n in m.state_dict())


def test_unknown_block_type_rejected():
    from animate_anything_b200.unet_3d_condition_mask import UNet3DConditionModel
    with pytest.raises(ValueError):
        UNet3DConditionModel(**dict(TINY, down_block_types=("NoSuchBlock3D", "CrossAttnDownBlock3D", "CrossAttnDownBlock3D",
                                                            "DownBlock3D")))
