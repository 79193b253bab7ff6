"""Host-side logic of the T5 / CLIP prompt encoders (no GPU): the integer relative-position buckets
against transformers' own function (bit-exact), known bucket values, and the `encode_prompt`
contract of the reference (univa/utils/denoiser_prompt_embedding_flux.py:107-144) with stub encoders."""
import pytest
import torch

from gpt_image_edit_b200.text_encoders import (EncoderOutput, SyntheticTokenizer, _encode_prompt_with_t5, encode_prompt,
                                               t5_relative_position_bucket)


def test_relative_position_buckets_known_values():
    b = t5_relative_position_bucket(300)
    assert b.dtype == torch.int64 and b.shape == (300, 300)
    assert b[0, 0] == 0 and b[5, 5] == 0
    # memory to the right of the query: +16; exact buckets below 8, log-spaced up to distance 128, then saturated
    assert b[0, 1] == 17 and b[0, 7] == 23 and b[0, 8] == 24 and b[0, 127] == 31 and b[0, 299] == 31
    assert b[1, 0] == 1 and b[7, 0] == 7 and b[8, 0] == 8 and b[200, 0] == 15
    assert int(b.min()) == 0 and int(b.max()) == 31


@pytest.mark.parametrize("L", [1, 7, 77, 256, 512])
def test_relative_position_buckets_match_transformers(L):
    t5 = pytest.importorskip("transformers.models.t5.modeling_t5")
    rel = torch.arange(L)[None, :] - torch.arange(L)[:, None]
    want = t5.T5Attention._relative_position_bucket(rel, bidirectional=True, num_buckets=32, max_distance=128)
    assert torch.equal(t5_relative_position_bucket(L), want)


class _StubClip:
    dtype, device = torch.float32, torch.device("cpu")

    def __call__(self, ids, output_hidden_states=False):
        h = ids.float()[..., None].repeat(1, 1, 4)
        return EncoderOutput(h, h[:, 0] + 1)


class _StubT5(_StubClip):
    def __call__(self, ids):
        return EncoderOutput(ids.float()[..., None].repeat(1, 1, 6))


def test_encode_prompt_contract():
    toks = [SyntheticTokenizer.clip(), SyntheticTokenizer.t5()]
    emb, pooled = encode_prompt([_StubClip(), _StubT5()], toks, ["make the sky red", "b"], 16, num_images_per_prompt=3)
    assert emb.shape == (6, 16, 6) and pooled.shape == (6, 4)
    # the reference duplicates by repeat(1, n, 1).view(B*n, L, -1): each prompt's copies are adjacent
    assert torch.equal(emb[0], emb[1]) and torch.equal(emb[0], emb[2]) and torch.equal(emb[3], emb[5])
    assert not torch.equal(emb[0], emb[3])
    # ... while the 2-D pooled tensor goes through repeat(1, n, 1) and comes out interleaved (b0, b1, b0, b1, ...):
    # the reference's behaviour, pinned by tests/golden/host_ref.pt
    assert torch.equal(pooled[0], pooled[2]) and torch.equal(pooled[1], pooled[3])
    # an encoder runs only when both it and its tokenizer are given (reference :120, :133)
    emb, pooled = encode_prompt([_StubClip(), _StubT5()], [None, toks[1]], "a", 16)
    assert pooled is None and emb.shape == (1, 16, 6)
    emb, pooled = encode_prompt([_StubClip(), None], toks, "a", 16, device="cpu")
    assert emb is None and pooled.shape == (1, 4)
    # helper-level error behaviour of the reference (:39-42)
    with pytest.raises(ValueError, match="text_input_ids must be provided"):
        _encode_prompt_with_t5(_StubT5(), None, 16, "a")
    ids = torch.arange(16)[None]
    assert _encode_prompt_with_t5(_StubT5(), None, 16, "a", text_input_ids=ids).shape == (1, 16, 6)


def test_synthetic_tokenizer_shapes():
    t = SyntheticTokenizer.clip()("a b c", max_length=77).input_ids
    assert t.shape == (1, 77) and t[0, 0] == 49406 and t[0, 4] == 49407 and int(t.argmax()) == 4
    long = SyntheticTokenizer.t5()(["w " * 600, "x"], max_length=256).input_ids
    assert long.shape == (2, 256) and long[0, -1] == 1 and long[1, 1] == 1 and long[1, 2] == 0
    assert int(long.max()) < 32128 and int(long.min()) >= 0
    assert torch.equal(SyntheticTokenizer.t5()("same text", max_length=8).input_ids,
                       SyntheticTokenizer.t5()("same text", max_length=8).input_ids)


def test_encoder_output_protocol():
    o = EncoderOutput(torch.zeros(1, 2, 3))
    assert o[0].shape == (1, 2, 3) and o.pooler_output is None
    o = EncoderOutput(torch.zeros(1, 2, 3), torch.ones(1, 3))
    assert o[1].shape == (1, 3) and o.last_hidden_state is o[0]
