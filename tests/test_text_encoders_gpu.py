"""GPU parity of the T5 / CLIP prompt encoders on libb2f kernels against
transformers' `T5EncoderModel` / `CLIPTextModel` — the classes `encode_prompt` drives in the reference
(univa/utils/denoiser_prompt_embedding_flux.py:44, :91) — with identical random weights.

Tolerance rule (same as the other model tests): the checker runs in fp32 with bf16-rounded weights;
kernel error vs fp32 must stay within 2x the error of the checker's own bf16 run + a small floor."""
import pytest
import torch

pytestmark = pytest.mark.gpu
t5_mod = pytest.importorskip("transformers.models.t5.modeling_t5")
clip_mod = pytest.importorskip("transformers.models.clip.modeling_clip")


def _rel_l2(a, b):
    return ((a.float() - b.float()).norm() / b.float().norm().clamp_min(1e-20)).item()


def _perturb(model):
    with torch.no_grad():
        for p in model.parameters():          # non-trivial norm weights / biases everywhere
            if p.dim() == 1:
                p.add_(0.05 * torch.randn_like(p))
    return model


def _run_both(ref, fn):
    """checker in bf16 and in fp32 (with the same bf16-rounded weights)."""
    ref16 = ref.to("cuda", torch.bfloat16)
    with torch.no_grad():
        o16 = fn(ref16)
    ref32 = ref16.to(torch.float32)
    with torch.no_grad():
        o32 = fn(ref32)
    return o16, o32


def _t5(d_model, heads, d_ff, layers, vocab, seed=0):
    from transformers import T5Config

    from gpt_image_edit_b200.text_encoders import B200T5Encoder, T5EncoderConfig

    torch.manual_seed(seed)
    cfg = T5Config(vocab_size=vocab, d_model=d_model, d_kv=64, num_heads=heads, d_ff=d_ff, num_layers=layers,
                   feed_forward_proj="gated-gelu", dropout_rate=0.0, is_encoder_decoder=False, use_cache=False)
    ref = t5_mod.T5EncoderModel(cfg).eval()
    with torch.no_grad():
        for n, p in ref.named_parameters():
            if p.dim() == 2 and "relative_attention_bias" not in n and "shared" not in n and "embed_tokens" not in n:
                p.normal_(0, 0.03)             # HF's default T5 init is tiny for wide layers; keep activations O(1)
    _perturb(ref)
    mine = B200T5Encoder(T5EncoderConfig(vocab_size=vocab, d_model=d_model, d_kv=64, num_heads=heads, d_ff=d_ff, num_layers=layers))
    mine.load_state_dict({k: v.detach().to(torch.bfloat16) for k, v in ref.state_dict().items()})
    return ref, mine


@pytest.mark.parametrize("shape", [(2, 40), (1, 256), (1, 300)])
def test_t5_toy_matches_transformers(shape):
    ref, mine = _t5(256, 4, 512, 2, 128)
    B, L = shape
    ids = torch.randint(0, 128, (B, L), generator=torch.Generator().manual_seed(5)).cuda()
    o16, o32 = _run_both(ref, lambda m: m(input_ids=ids).last_hidden_state)
    out = mine(ids)[0]
    ek, et = _rel_l2(out, o32), _rel_l2(o16, o32)
    print(f"T5 toy {shape}: kernel-vs-fp32 {ek:.3e}  torch-bf16-vs-fp32 {et:.3e}")
    assert out.shape == o32.shape and out.dtype == torch.bfloat16
    assert ek <= 2.0 * et + 3e-3


def test_t5_xxl_width_one_block():
    """Full T5-XXL widths (d=4096, 64 heads x 64, d_ff=10240, L=256 as the reference pads to), one block."""
    ref, mine = _t5(4096, 64, 10240, 1, 512, seed=1)
    ids = torch.randint(0, 512, (1, 256), generator=torch.Generator().manual_seed(6)).cuda()
    o16, o32 = _run_both(ref, lambda m: m(input_ids=ids).last_hidden_state)
    out = mine(ids).last_hidden_state
    ek, et = _rel_l2(out, o32), _rel_l2(o16, o32)
    print(f"T5-XXL width: kernel-vs-fp32 {ek:.3e}  torch-bf16-vs-fp32 {et:.3e}")
    assert ek <= 2.0 * et + 3e-3


def test_t5_state_dict_round_trip():
    ref, mine = _t5(256, 4, 512, 2, 128)
    sd = mine.state_dict()
    for k, v in ref.state_dict().items():
        assert torch.equal(sd[k].cpu(), v.detach().to(torch.bfloat16)), k


def _clip(hidden, heads, inter, layers, vocab, seed=0):
    from transformers import CLIPTextConfig as HFCfg

    from gpt_image_edit_b200.text_encoders import B200CLIPTextModel, CLIPTextConfig

    torch.manual_seed(seed)
    cfg = HFCfg(vocab_size=vocab, hidden_size=hidden, intermediate_size=inter, num_hidden_layers=layers,
                num_attention_heads=heads, max_position_embeddings=77, hidden_act="quick_gelu", eos_token_id=2,
                attn_implementation="eager")
    ref = clip_mod.CLIPTextModel(cfg).eval()
    with torch.no_grad():
        for p in ref.parameters():
            if p.dim() == 2:
                p.normal_(0, 0.03)
    _perturb(ref)
    mine = B200CLIPTextModel(CLIPTextConfig(vocab_size=vocab, hidden_size=hidden, intermediate_size=inter,
                                            num_hidden_layers=layers, num_attention_heads=heads))
    mine.load_state_dict({k: v.detach().to(torch.bfloat16) for k, v in ref.state_dict().items()})
    return ref, mine


def _clip_ids(B, L, vocab, seed):
    g = torch.Generator().manual_seed(seed)
    ids = torch.randint(1, vocab - 1, (B, L), generator=g)
    for b in range(B):                         # EOS = largest id, at a different position per row, padding after it
        e = 5 + 11 * b
        ids[b, e] = vocab - 1
        ids[b, e + 1:] = vocab - 1 if b % 2 else 0
    return ids


@pytest.mark.parametrize("cfg", [(256, 4, 512, 2, 200), (768, 12, 3072, 12, 49408)], ids=["toy", "clip-l"])
def test_clip_matches_transformers(cfg):
    ref, mine = _clip(*cfg)
    ids = _clip_ids(2, 77, cfg[4], 7).cuda()
    o16, o32 = _run_both(ref, lambda m: (lambda o: (o.last_hidden_state, o.pooler_output))(m(input_ids=ids)))
    out = mine(ids, output_hidden_states=False)
    for name, mine_t, t16, t32 in (("hidden", out.last_hidden_state, o16[0], o32[0]), ("pooled", out.pooler_output, o16[1], o32[1])):
        ek, et = _rel_l2(mine_t, t32), _rel_l2(t16, t32)
        print(f"CLIP {name}: kernel-vs-fp32 {ek:.3e}  torch-bf16-vs-fp32 {et:.3e}")
        assert mine_t.shape == t32.shape
        assert ek <= 2.0 * et + 3e-3
    sd = mine.state_dict()
    for k, v in ref.state_dict().items():
        if "position_ids" in k:
            continue
        assert torch.equal(sd[k].cpu(), v.detach().to(torch.bfloat16).cpu()), k


def test_encoder_kernels_match_eager_chains():
    from gpt_image_edit_b200 import ops

    g = torch.Generator(device="cuda").manual_seed(2)
    r = lambda *s, k=1.0: (torch.randn(*s, device="cuda", generator=g) * k).bfloat16()
    # gated GELU (T5DenseGatedActDense): one fp32 gelu then the product — vs the fp32 evaluation
    gu = r(300, 2 * 1024, k=2.0)
    want = torch.nn.functional.gelu(gu[:, :1024].float(), approximate="tanh").bfloat16().float() * gu[:, 1024:].float()
    assert torch.equal(ops.geglu(gu, 1024), want.bfloat16())
    # LayerNorm with affine parameters
    x, w, b = r(77, 768, k=3.0), (1 + 0.1 * torch.randn(768, device="cuda", generator=g)).bfloat16(), r(768, k=0.1)
    want = torch.nn.functional.layer_norm(x.float(), (768,), w.float(), b.float(), 1e-5)
    got = ops.layernorm(x, w, b, eps=1e-5)
    assert (got.float() - want).abs().max().item() <= 2.0 ** -7 * want.abs().max().item()
    assert _rel_l2(got, want) < 3e-3
    # embedding lookup + position add (bf16 add of two bf16 values: exact one-rounding)
    tok, pos = r(500, 256), r(77, 256)
    ids = torch.randint(0, 500, (2 * 77,), device="cuda", generator=g)
    want = (tok[ids].float() + pos.repeat(2, 1).float()).bfloat16()
    assert torch.equal(ops.embed(tok, ids, pos, period=77), want)
    assert torch.equal(ops.embed(tok, ids), tok[ids])
    # quick-GELU epilogue == the eager bf16 chain of QuickGELUActivation on the bf16 GEMM output
    a, wt, bias = r(130, 256), r(512, 256, k=0.1), r(512, k=0.1)
    y = ops.linear(a, wt, bias)
    want = y * torch.sigmoid(1.702 * y)
    got = ops.linear(a, wt, bias, epilogue=ops.EPI_QUICK_GELU)
    assert (got.float() - want.float()).abs().max().item() <= 2.0 ** -6 * want.float().abs().max().item()


@pytest.mark.parametrize("shape", [(1, 4, 40, 40), (2, 3, 256, 256), (1, 2, 300, 300)])
def test_attention_with_additive_bias(shape):
    from gpt_image_edit_b200 import ops

    B, H, S, _ = shape
    g = torch.Generator(device="cuda").manual_seed(3)
    q, k, v = ((torch.randn(B, S, H, 128, device="cuda", generator=g)).bfloat16() for _ in range(3))
    q = q * 0.3
    bias = (torch.randn(H, S, S, device="cuda", generator=g) * 2).bfloat16()
    out = ops.attention(q, k, v, scale=1.0, bias=bias)
    s = torch.einsum("bqhd,bkhd->bhqk", q.float(), k.float()) + bias.float()[None]
    want = torch.einsum("bhqk,bkhd->bqhd", s.softmax(-1), v.float()).reshape(B, S, H * 128)
    p16 = s.softmax(-1).bfloat16().float()
    t16 = torch.einsum("bhqk,bkhd->bqhd", p16, v.float()).reshape(B, S, H * 128).bfloat16()
    ek, et = _rel_l2(out, want), _rel_l2(t16, want)
    print(f"bias attention {shape}: kernel {ek:.3e} bf16-P reference {et:.3e}")
    assert ek <= 2.0 * et + 2e-3
