"""Host-side behaviour of LatentToVideoPipeline that needs no GPU: argument checking raises in the same cases as the
(restated) diffusers TextToVideoSDPipeline.check_inputs the reference inherits (models/pipeline.py:113-115), and
_encode_prompt builds `cat([negative, positive])` (models/pipeline.py:136-144) through a tokenizer / text encoder."""
import os
import sys
import types

import pytest
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle", "shim"))


def _pipe(text_encoder=None, tokenizer=None):
    from animate_anything_b200.pipeline import LatentToVideoPipeline
    unet = types.SimpleNamespace(dtype=torch.float16, config=types.SimpleNamespace(sample_size=32, in_channels=4))
    vae = types.SimpleNamespace(config=types.SimpleNamespace(block_out_channels=(1, 2, 3, 4), scaling_factor=0.18215))
    return LatentToVideoPipeline(vae=vae, text_encoder=text_encoder, tokenizer=tokenizer, unet=unet, scheduler=None)


BAD = [
    dict(prompt="a", height=257, width=256, callback_steps=1),
    dict(prompt="a", height=256, width=256, callback_steps=0),
    dict(prompt="a", height=256, width=256, callback_steps=None),
    dict(prompt="a", height=256, width=256, callback_steps=1, prompt_embeds=torch.zeros(1, 7, 8)),
    dict(prompt=None, height=256, width=256, callback_steps=1),
    dict(prompt=3, height=256, width=256, callback_steps=1),
    dict(prompt="a", height=256, width=256, callback_steps=1, negative_prompt="b",
         negative_prompt_embeds=torch.zeros(1, 7, 8)),
    dict(prompt=None, height=256, width=256, callback_steps=1, prompt_embeds=torch.zeros(1, 7, 8),
         negative_prompt_embeds=torch.zeros(1, 6, 8)),
]
GOOD = [
    dict(prompt="a", height=256, width=512, callback_steps=1),
    dict(prompt=["a", "b"], height=256, width=256, callback_steps=2, negative_prompt="c"),
    dict(prompt=None, height=256, width=256, callback_steps=1, prompt_embeds=torch.zeros(1, 7, 8),
         negative_prompt_embeds=torch.zeros(1, 7, 8)),
]


def test_check_inputs_raises_like_the_diffusers_restatement():
    from diffusers._impl import TextToVideoSDPipeline
    ref = TextToVideoSDPipeline.__new__(TextToVideoSDPipeline)
    ours = _pipe()
    for kw in BAD:
        with pytest.raises(ValueError):
            ref.check_inputs(**kw)
        with pytest.raises(ValueError):
            ours.check_inputs(**kw)
    for kw in GOOD:
        ref.check_inputs(**kw)
        ours.check_inputs(**kw)


def test_call_requires_latents_and_condition_latent():
    """models/pipeline.py:126,161: the reference indexes `latents` / `condition_latent` unconditionally."""
    p = _pipe()
    with pytest.raises(ValueError):
        p(prompt_embeds=torch.zeros(1, 7, 8), negative_prompt_embeds=torch.zeros(1, 7, 8), latents=None,
          condition_latent=None)


class _Tok:
    model_max_length = 5

    def __call__(self, texts, padding, max_length, truncation, return_tensors):
        assert padding == "max_length" and max_length == 5 and truncation and return_tensors == "pt"
        ids = torch.zeros(len(texts), max_length, dtype=torch.long)
        for i, t in enumerate(texts):
            for j, ch in enumerate(t[:max_length]):
                ids[i, j] = ord(ch)
        return types.SimpleNamespace(input_ids=ids)


class _Enc:
    def __call__(self, ids):
        return (ids.float().unsqueeze(-1).repeat(1, 1, 8) / 100.0,)


def test_encode_prompt_orders_negative_then_positive():
    p = _pipe(_Enc(), _Tok())
    e = p._encode_prompt(["ab", "c"], torch.device("cpu"), 1, True, negative_prompt=None)
    assert e.shape == (4, 5, 8) and e.dtype == torch.float16
    assert float(e[:2].abs().max()) == 0.0                      # "" -> all-zero ids: the negative half comes FIRST
    assert float(e[2, 0, 0]) == pytest.approx(ord("a") / 100.0, rel=1e-3)
    e1 = p._encode_prompt("ab", torch.device("cpu"), 1, False)
    assert e1.shape == (1, 5, 8)
    e2 = p._encode_prompt(None, torch.device("cpu"), 1, True, prompt_embeds=torch.ones(1, 5, 8),
                          negative_prompt_embeds=torch.zeros(1, 5, 8))
    assert torch.equal(e2[0], torch.zeros(5, 8, dtype=torch.float16)) and torch.equal(e2[1], torch.ones(5, 8, dtype=torch.float16))
    with pytest.raises(ValueError):
        _pipe()._encode_prompt("needs an encoder", torch.device("cpu"), 1, False)
