"""Host logic of the VaeImageProcessor mirror (train.py:745 `vae_processor.preprocess(pimg, height, width)`)."""
import numpy as np
import pytest
import torch

from animate_anything_b200.image_processor import VaeImageProcessor


def test_preprocess_pil_matches_formula():
    PIL = pytest.importorskip("PIL.Image")
    rng = np.random.default_rng(0)
    arr = rng.integers(0, 256, size=(70, 90, 3), dtype=np.uint8)
    img = PIL.fromarray(arr)
    vp = VaeImageProcessor()
    out = vp.preprocess(img, 64, 80)
    assert out.shape == (1, 3, 64, 80) and out.dtype == torch.float32
    ref = np.asarray(img.resize((80, 64), resample=PIL.Resampling.LANCZOS)).astype(np.float32) / 255.0
    ref = torch.from_numpy(ref.transpose(2, 0, 1))[None] * 2.0 - 1.0
    assert torch.equal(out, ref)
    # default size: rounded DOWN to a multiple of 8
    assert vp.preprocess(img).shape == (1, 3, 64, 88)
    assert -1.0 <= float(out.min()) and float(out.max()) <= 1.0


def test_preprocess_tensor_numpy_and_roundtrip():
    vp = VaeImageProcessor()
    t = torch.rand(2, 3, 16, 24)
    out = vp.preprocess(t)
    assert torch.allclose(out, 2 * t - 1)
    n = np.random.default_rng(1).random((16, 24, 3)).astype(np.float32)
    out_n = vp.preprocess(n)
    assert out_n.shape == (1, 3, 16, 24) and torch.allclose(out_n[0], torch.from_numpy(n.transpose(2, 0, 1)) * 2 - 1)
    back = vp.postprocess(out, output_type="pt")
    assert torch.allclose(back, t, atol=1e-6)
    lat = torch.randn(1, 4, 8, 8)
    assert vp.preprocess(lat) is lat or torch.equal(vp.preprocess(lat), lat)
    with pytest.raises(ValueError):
        vp.preprocess("not an image")
    with pytest.raises(ValueError):
        VaeImageProcessor(do_convert_rgb=True, do_convert_grayscale=True)
