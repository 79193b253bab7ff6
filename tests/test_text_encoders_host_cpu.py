"""Host orchestration of the prompt encoders (gpt_image_edit_b200/text_encoders.py: B200T5Encoder, B200CLIPTextModel — weight
re-layout into 128-wide head slots, the T5 relative-position bias table, unscaled T5 attention, gated-GELU; CLIP causal
attention, quick-GELU, pooling at the EOS token) on the CPU with TORCH DOUBLES in place of the libb2f kernels and the
CUDA-only constructor guard lifted for the test (both by monkeypatch; the product has no CPU path).  Checker: transformers'
T5EncoderModel / CLIPTextModel on the same weights, plus the state-dict round trip of both layouts."""
from collections import OrderedDict

import pytest
import torch

pytest.importorskip("transformers")
BF = torch.bfloat16


def _rel(a, b):
    return ((a.float() - b.float()).norm() / b.float().norm()).item()


@pytest.fixture()
def cpu_encoders(monkeypatch):
    from gpt_image_edit_b200 import ops
    from gpt_image_edit_b200 import text_encoders as te

    def base_init(self, device="cuda"):
        torch.nn.Module.__init__(self)
        self._dev, self.W = torch.device("cpu"), OrderedDict()

    monkeypatch.setattr(te._Base, "__init__", base_init)

    def linear(x, weight, bias=None, *, epilogue=ops.EPI_BIAS, out=None, resid=None, gate=None):
        y = x.float() @ weight.float().t()
        if bias is not None:
            y = y + bias.float()
        if epilogue == ops.EPI_RESID:
            y = y + resid.float()
        elif epilogue == ops.EPI_QUICK_GELU:
            y = y * torch.sigmoid(1.702 * y)
        elif epilogue != ops.EPI_BIAS:
            raise AssertionError(epilogue)
        return y.to(BF)

    def attention(q, k, v, *, out=None, causal=False, scale=None, bias=None):
        B, S, H, dh = q.shape
        qf, kf, vf = (t.float().permute(0, 2, 1, 3) for t in (q, k, v))
        s = qf @ kf.transpose(-1, -2) * (dh ** -0.5 if scale is None else scale)
        if bias is not None:
            assert bias.shape == (H, S, S)
            s = s + bias.float()
        if causal:
            s = s.masked_fill(torch.ones(S, S, dtype=torch.bool).triu(1), float("-inf"))
        out.copy_((s.softmax(-1) @ vf).permute(0, 2, 1, 3).reshape(B, S, H * dh).to(BF))
        return out

    def rmsnorm(x, weight, *, out=None, eps=1e-6):
        xf = x.float()
        return (weight.float() * (xf * torch.rsqrt(xf.pow(2).mean(-1, keepdim=True) + eps))).to(BF)

    def layernorm(x, weight, bias, *, out=None, eps=1e-5):
        return torch.nn.functional.layer_norm(x.float(), (x.shape[-1],), weight.float(), bias.float(), eps).to(BF)

    def embed(table, ids, pos=None, *, period=0, out=None):
        y = table[ids].float()
        if pos is not None:
            y = y + pos[torch.arange(ids.numel()) % period].float()
        return y.to(BF)

    def geglu(gu, inter, *, out=None):
        return (torch.nn.functional.gelu(gu[:, :inter].float(), approximate="tanh").to(BF).float() * gu[:, inter:].float()).to(BF)

    doubles = dict(linear=linear, attention=attention, rmsnorm=rmsnorm, layernorm=layernorm, embed=embed, geglu=geglu,
                   gather_rows=lambda table, idx, out=None: table[idx].clone())
    for name, fn in doubles.items():
        monkeypatch.setattr(ops, name, fn)
    return te


def test_t5_encoder_orchestration_and_state_dict_round_trip(cpu_encoders):
    from transformers import T5Config, T5EncoderModel

    te = cpu_encoders
    kw = dict(vocab_size=128, d_model=256, d_kv=64, num_heads=4, d_ff=512, num_layers=2)
    torch.manual_seed(0)
    ref = T5EncoderModel(T5Config(**kw, feed_forward_proj="gated-gelu", is_encoder_decoder=False, use_cache=False,
                                  relative_attention_num_buckets=32, relative_attention_max_distance=128)).eval().float()
    with torch.no_grad():
        for p in ref.parameters():
            p.copy_((p + (0.05 * torch.randn_like(p) if p.dim() == 1 else 0)).to(BF).float())
    enc = te.B200T5Encoder(te.T5EncoderConfig(**kw), device="cpu")
    sd = {k: v.detach() for k, v in ref.state_dict().items()}
    enc.load_state_dict(sd)
    back = enc.state_dict()
    for k, v in sd.items():
        assert torch.equal(back[k], v.to(BF)), k                               # both layouts hold the same numbers
    ids = torch.randint(0, 128, (2, 40), generator=torch.Generator().manual_seed(1))
    out = enc(ids)[0]
    with torch.no_grad():
        want = ref(input_ids=ids).last_hidden_state
    assert out.shape == want.shape == (2, 40, 256) and _rel(out, want) < 2e-2
    assert enc.position_bias(40).shape == (4, 40, 40) and enc.position_bias(40) is enc.position_bias(40)   # cached per length
    from gpt_image_edit_b200 import _lib
    with pytest.raises(_lib.B2FError, match="padding masks"):
        enc(ids, attention_mask=torch.tensor([[1] * 39 + [0]] * 2))


def test_clip_text_model_orchestration_pooling_and_round_trip(cpu_encoders):
    from transformers import CLIPTextConfig as HFCfg
    from transformers import CLIPTextModel

    te = cpu_encoders
    kw = dict(vocab_size=200, hidden_size=256, intermediate_size=512, num_hidden_layers=2, num_attention_heads=4,
              max_position_embeddings=77)
    torch.manual_seed(0)
    ref = CLIPTextModel(HFCfg(**kw, hidden_act="quick_gelu", eos_token_id=2, bos_token_id=0, pad_token_id=1)).eval().float()
    with torch.no_grad():
        for p in ref.parameters():
            p.copy_((p + (0.05 * torch.randn_like(p) if p.dim() == 1 else 0)).to(BF).float())
    enc = te.B200CLIPTextModel(te.CLIPTextConfig(**kw, eos_token_id=2), device="cpu")
    sd = {k: v.detach() for k, v in ref.state_dict().items() if "position_ids" not in k}
    enc.load_state_dict(sd)
    back = enc.state_dict()
    for k, v in sd.items():
        assert torch.equal(back[k], v.to(BF)), k
    g = torch.Generator().manual_seed(2)
    ids = torch.randint(3, 190, (2, 77), generator=g)
    ids[0, 20], ids[1, 55] = 199, 199                         # the legacy rule pools at argmax(ids): the highest id marks EOS
    out = enc(ids, output_hidden_states=False)
    with torch.no_grad():
        want = ref(input_ids=ids)
    assert _rel(out[0], want.last_hidden_state) < 2e-2
    assert out.pooler_output.shape == (2, 256) and _rel(out.pooler_output, want.pooler_output) < 2e-2
    assert torch.equal(out.pooler_output[0], out[0][0, 20]) and torch.equal(out.pooler_output[1], out[0][1, 55])
    with pytest.raises(ValueError, match="max_position_embeddings"):
        enc(torch.zeros(1, 78, dtype=torch.long))
