"""Host orchestration of the Qwen2.5-VL decoder prefill (gpt_image_edit_b200/qwen2p5vl.py: B200Qwen2p5VL.forward /
_decoder / _rope_tables / padding_spans) run on the CPU with TORCH DOUBLES in place of the libb2f kernels (`ops.*` is
monkeypatched inside this test only; the product has no such path and raises on CPU tensors).  What this checks is the
part that is Python: which rows each launch sees, the views and pitches handed to the attention call, the zeroed rows of
a padded batch, M-RoPE tables, weight fusion order.  Checker: transformers' Qwen2_5_VLModel on the same weights.

Padded batches follow transformers' flash_attention_2 semantics (the backend the reference selects, cli.py:40): a sequence
attends to its own tokens only and a padding row receives a zero attention output, i.e. it only passes through the MLPs."""
import types

import pytest
import torch

hf = pytest.importorskip("transformers.models.qwen2_5_vl.modeling_qwen2_5_vl")
BF = torch.bfloat16


def _doubles():
    from gpt_image_edit_b200 import ops

    def rmsnorm(x, weight, *, out=None, eps=1e-6):
        xf = x.float()
        return (weight.float() * (xf * torch.rsqrt(xf.pow(2).mean(-1, keepdim=True) + eps))).to(BF)

    def linear(x, weight, bias=None, *, epilogue=ops.EPI_BIAS, out=None, resid=None, gate=None):
        y = x.float() @ weight.float().t()
        if bias is not None:
            y = y + bias.float()
        if epilogue == ops.EPI_RESID:
            y = y + resid.float()
        elif epilogue == ops.EPI_GELU_ERF:
            y = torch.nn.functional.gelu(y.to(BF).float())
        elif epilogue != ops.EPI_BIAS:
            raise AssertionError(f"epilogue {epilogue} is not used by the Qwen2.5-VL host code")
        y = y.to(BF)
        if out is not None:
            out.copy_(y)
            return out
        return y

    def rope_half_(x, heads, head_pitch, cos, sin, *, fp32_math):
        # cos / sin hold `rot` columns (the rotary width: the whole head for the decoder, 80 of a 128-wide slot for the ViT);
        # columns beyond it are padding and stay untouched
        rot = cos.shape[-1]
        v = x[:, : heads * head_pitch].float().reshape(x.shape[0], heads, head_pitch)
        a = v[..., :rot]
        r = torch.cat((-a[..., rot // 2:], a[..., : rot // 2]), dim=-1)
        v = torch.cat((a * cos[:, None, :] + r * sin[:, None, :], v[..., rot:]), dim=-1)
        x[:, : heads * head_pitch] = v.reshape(x.shape[0], -1).to(BF)
        return x

    def attention(q, k, v, *, out=None, causal=False, scale=None, bias=None):
        assert bias is None and q.stride(-1) == 1 and q.stride(2) == q.shape[3]          # what ops.attention requires
        B, Sq, H, dh = q.shape
        Skv, Hkv = k.shape[1], k.shape[2]
        qf, kf, vf = (t.float().permute(0, 2, 1, 3) for t in (q, k, v))
        kf, vf = (t.repeat_interleave(H // Hkv, dim=1) for t in (kf, vf))
        s = qf @ kf.transpose(-1, -2) * (scale or dh ** -0.5)
        if causal:
            s = s.masked_fill(torch.ones(Sq, Skv, dtype=torch.bool).triu(Skv - Sq + 1), float("-inf"))
        o = (s.softmax(-1) @ vf).permute(0, 2, 1, 3).reshape(B, Sq, H * dh)
        assert out.shape == o.shape
        out.copy_(o.to(BF))
        return out

    def swiglu(gu, inter, *, out=None):
        return (torch.nn.functional.silu(gu[:, :inter].float()) * gu[:, inter:2 * inter].float()).to(BF)

    def gather_rows(table, idx, *, out=None):
        return table[idx].clone()

    def scatter_rows_(dst, idx, src):
        dst[idx] = src
        return dst

    return dict(rmsnorm=rmsnorm, linear=linear, rope_half_=rope_half_, attention=attention, swiglu=swiglu, gather_rows=gather_rows,
                scatter_rows_=scatter_rows_)


@pytest.fixture()
def engine(monkeypatch):
    from transformers.models.qwen2_5_vl.configuration_qwen2_5_vl import Qwen2_5_VLConfig

    from gpt_image_edit_b200 import ops
    from gpt_image_edit_b200.qwen2p5vl import B200Qwen2p5VL, QwenTextConfig, QwenVisionConfig

    for name, fn in _doubles().items():
        monkeypatch.setattr(ops, name, fn)
    tc = dict(hidden_size=256, num_hidden_layers=2, num_attention_heads=2, num_key_value_heads=1, intermediate_size=512,
              vocab_size=1000, rms_norm_eps=1e-6)
    cfg = Qwen2_5_VLConfig(
        text_config=dict(tc, rope_parameters=dict(rope_type="default", rope_theta=1e6, mrope_section=[16, 24, 24])),
        vision_config=dict(depth=1, hidden_size=256, num_heads=4, intermediate_size=340, out_hidden_size=256,
                           fullatt_block_indexes=[0]),
        image_token_id=900, video_token_id=901, vision_start_token_id=902, vision_end_token_id=903)
    torch.manual_seed(0)
    ref = hf.Qwen2_5_VLModel(cfg).eval().float()
    with torch.no_grad():
        for p in ref.parameters():
            if p.dim() == 1:
                p.add_(0.05 * torch.randn_like(p))
            p.copy_(p.to(BF).float())                        # both sides hold the same bf16-representable weights
    lm = ref.language_model
    me = types.SimpleNamespace(tc=QwenTextConfig(**tc, image_token_id=900, video_token_id=901, vision_start_token_id=902),
                               vc=QwenVisionConfig(), thd=128, _dev=torch.device("cpu"), W={})
    W = me.W
    W["model.embed_tokens"] = lm.embed_tokens.weight.detach().to(BF)
    for i, layer in enumerate(lm.layers):                    # the fused layout B200Qwen2p5VL.load_state_dict builds
        p, a, m = f"model.layers.{i}.", layer.self_attn, layer.mlp
        W[p + "ln1"], W[p + "ln2"] = layer.input_layernorm.weight.detach().to(BF), layer.post_attention_layernorm.weight.detach().to(BF)
        W[p + "qkv.w"] = torch.cat([a.q_proj.weight, a.k_proj.weight, a.v_proj.weight]).detach().to(BF)
        W[p + "qkv.b"] = torch.cat([a.q_proj.bias, a.k_proj.bias, a.v_proj.bias]).detach().to(BF)
        W[p + "o.w"] = a.o_proj.weight.detach().to(BF)
        W[p + "gu.w"] = torch.cat([m.gate_proj.weight, m.up_proj.weight]).detach().to(BF)
        W[p + "down.w"] = m.down_proj.weight.detach().to(BF)
    W["model.norm"] = lm.norm.weight.detach().to(BF)
    for name in ("_rope_tables", "_decoder"):
        setattr(me, name, types.MethodType(getattr(B200Qwen2p5VL, name), me))
    run = lambda ids, mask=None: B200Qwen2p5VL.forward(me, ids, attention_mask=mask)
    return ref, run


def _rel(a, b):
    return ((a.float() - b.float()).norm() / b.float().norm()).item()


def _hf(ref, ids, mask=None):
    from gpt_image_edit_b200.qwen2p5vl import get_rope_index
    pos, _ = get_rope_index(ids, None, mask)
    with torch.no_grad():
        return ref(input_ids=ids, attention_mask=mask, position_ids=pos).last_hidden_state


def test_unpadded_prefill_orchestration_matches_transformers(engine):
    ref, run = engine
    ids = torch.randint(1, 800, (2, 23), generator=torch.Generator().manual_seed(1))
    h = run(ids)
    assert h.shape == (2, 23, 256) and h.dtype == BF
    assert _rel(h, _hf(ref, ids)) < 2e-2
    assert _rel(run(ids, torch.ones_like(ids)), _hf(ref, ids)) < 2e-2              # an all-ones mask is the same path


@pytest.mark.parametrize("side", ["right", "left"])
def test_padded_batch_orchestration(engine, side):
    ref, run = engine
    g = torch.Generator().manual_seed(2)
    rows = [torch.randint(1, 800, (n,), generator=g).tolist() for n in (19, 7, 12)]
    n, PAD = 19, 3
    pad = lambda r, fill: (r + [fill] * (n - len(r))) if side == "right" else ([fill] * (n - len(r)) + r)
    ids = torch.tensor([pad(r, PAD) for r in rows])
    mask = torch.tensor([pad([1] * len(r), 0) for r in rows])
    h = run(ids, mask)
    assert h.shape == (3, n, 256) and torch.isfinite(h.float()).all()
    lm = ref.language_model
    with torch.no_grad():
        x = lm.embed_tokens(torch.tensor([PAD]))
        for layer in lm.layers:                                   # a row whose attention output is zero
            x = x + layer.mlp(layer.post_attention_layernorm(x))
        pad_want = lm.norm(x)[0]
    for b, r in enumerate(rows):
        real = mask[b].bool()
        alone = _hf(ref, torch.tensor([r]))[0]                    # the prompt on its own, no padding
        assert _rel(h[b][real], alone) < 2e-2, (side, b)
        for row in h[b][~real]:
            assert _rel(row, pad_want) < 2e-2, (side, b)
    if side == "right":                                           # transformers given the same mask agrees on the real tokens
        full = _hf(ref, ids, mask)
        for b in range(3):
            assert _rel(h[b][mask[b].bool()], full[b][mask[b].bool()]) < 2e-2


def test_padding_spans_and_refusals():
    from gpt_image_edit_b200 import _lib
    from gpt_image_edit_b200.qwen2p5vl import padding_spans

    assert padding_spans(torch.tensor([[1, 1, 0, 0], [0, 1, 1, 1], [1, 1, 1, 1]])) == [(0, 2), (1, 4), (0, 4)]
    for bad in ([[1, 0, 1, 1]], [[0, 0, 0, 0]]):
        with pytest.raises(_lib.B2FError):
            padding_spans(torch.tensor(bad))
    with pytest.raises(_lib.B2FError):
        padding_spans(torch.ones(4))


def test_vision_tower_and_image_prefill_orchestration(monkeypatch):
    """forward_visual (patch embed, 2-D rope in window order, windowed / full block-diagonal attention launched per run of
    equal segments, 80-wide heads in 128-wide slots, padded MLP width, 2x2 merger, un-windowing) and the scatter of the image
    embeddings into the prompt, on the CPU over the torch doubles, against transformers' Qwen2_5_VLModel with the weights
    loaded through the product's own load_state_dict."""
    from transformers.models.qwen2_5_vl.configuration_qwen2_5_vl import Qwen2_5_VLConfig

    from gpt_image_edit_b200 import ops
    from gpt_image_edit_b200.qwen2p5vl import B200Qwen2p5VL, QwenTextConfig, QwenVisionConfig, _pad8, get_rope_index

    for name, fn in _doubles().items():
        monkeypatch.setattr(ops, name, fn)
    tcd = dict(hidden_size=256, num_hidden_layers=2, num_attention_heads=2, num_key_value_heads=1, intermediate_size=512,
               vocab_size=1000, rms_norm_eps=1e-6)
    vcd = dict(depth=3, hidden_size=320, num_heads=4, intermediate_size=340, out_hidden_size=256, fullatt_block_indexes=[1])
    cfg = Qwen2_5_VLConfig(text_config=dict(tcd, rope_parameters=dict(rope_type="default", rope_theta=1e6, mrope_section=[16, 24, 24])),
                           vision_config=vcd, image_token_id=900, video_token_id=901, vision_start_token_id=902, vision_end_token_id=903)
    torch.manual_seed(0)
    ref = hf.Qwen2_5_VLModel(cfg).eval().float()
    with torch.no_grad():
        for p in ref.parameters():
            if p.dim() == 1:
                p.add_(0.05 * torch.randn_like(p))
            p.copy_(p.to(BF).float())
    tc = QwenTextConfig(**tcd, image_token_id=900, video_token_id=901, vision_start_token_id=902)
    vc = QwenVisionConfig(**{**vcd, "fullatt_block_indexes": (1,)})
    me = types.SimpleNamespace(tc=tc, vc=vc, HP=B200Qwen2p5VL.HP, vhd=320 // 4, vi=_pad8(340), thd=128, _dev=torch.device("cpu"),
                               W=B200Qwen2p5VL.alloc_weights(tc, vc, torch.device("cpu")))
    sd = {k.replace("language_model.", "model."): v.detach() for k, v in ref.state_dict().items()}
    B200Qwen2p5VL.load_state_dict(me, sd)
    for name in ("_rope_tables", "_decoder", "_segment_attention", "forward_visual"):
        setattr(me, name, types.MethodType(getattr(B200Qwen2p5VL, name), me))
    g = torch.Generator().manual_seed(3)
    grid = torch.tensor([[1, 16, 8], [1, 4, 4]])                      # two images: 32 + 4 merged tokens, windows of unequal sizes
    pix = torch.randn(128 + 16, 1176, generator=g).to(BF)
    with torch.no_grad():
        v_ref = ref.visual(pix.float(), grid_thw=grid).pooler_output
    v = me.forward_visual(pix, grid)
    assert v.shape == v_ref.shape == (36, 256)
    assert _rel(v, v_ref) < 2e-2
    ids = torch.tensor([[1, 2, 902] + [900] * 32 + [903, 5, 902] + [900] * 4 + [903] + list(range(10, 22))])
    h = B200Qwen2p5VL.forward(me, ids, pixel_values=pix, image_grid_thw=grid)
    pos, _ = get_rope_index(ids, grid, None, spatial_merge_size=2, image_token_id=900, vision_start_token_id=902)
    with torch.no_grad():
        h_ref = ref(input_ids=ids, pixel_values=pix.float(), image_grid_thw=grid, position_ids=pos).last_hidden_state
    assert h.shape == h_ref.shape and _rel(h, h_ref) < 2e-2
    with pytest.raises(ValueError, match="do not match"):
        B200Qwen2p5VL.forward(me, ids[:, :-8 - 6], pixel_values=pix, image_grid_thw=grid)     # the second image's tokens cut off


def test_kv_cache_decode_orchestration(monkeypatch):
    """generate() — prefill into the KV cache, then one token per step with single-query attention over the cached prefix,
    M-RoPE positions continued at past + delta, lm_head, greedy pick, eos / padding handling — on the CPU over the torch doubles:
    teacher-forced with transformers' greedy reply, the per-step logits must match its scores."""
    from transformers.models.qwen2_5_vl.configuration_qwen2_5_vl import Qwen2_5_VLConfig

    from gpt_image_edit_b200 import ops
    from gpt_image_edit_b200.qwen2p5vl import B200Qwen2p5VL, QwenTextConfig, QwenVisionConfig, _pad8

    for name, fn in _doubles().items():
        monkeypatch.setattr(ops, name, fn)
    tcd = dict(hidden_size=256, num_hidden_layers=2, num_attention_heads=2, num_key_value_heads=1, intermediate_size=512,
               vocab_size=1000, rms_norm_eps=1e-6)
    vcd = dict(depth=1, hidden_size=256, num_heads=4, intermediate_size=340, out_hidden_size=256, fullatt_block_indexes=[0])
    cfg = Qwen2_5_VLConfig(text_config=dict(tcd, rope_parameters=dict(rope_type="default", rope_theta=1e6, mrope_section=[16, 24, 24])),
                           vision_config=vcd, image_token_id=900, video_token_id=901, vision_start_token_id=902, vision_end_token_id=903)
    torch.manual_seed(1)
    ref = hf.Qwen2_5_VLForConditionalGeneration(cfg).eval().float()
    with torch.no_grad():
        for p in ref.parameters():
            if p.dim() == 1:
                p.add_(0.05 * torch.randn_like(p))
        ref.lm_head.weight.normal_(0, 0.2)
        for p in ref.parameters():
            p.copy_(p.to(BF).float())
    tc = QwenTextConfig(**tcd, image_token_id=900, video_token_id=901, vision_start_token_id=902)
    vc = QwenVisionConfig(**{**vcd, "fullatt_block_indexes": (0,)})
    me = types.SimpleNamespace(tc=tc, vc=vc, HP=B200Qwen2p5VL.HP, vhd=64, vi=_pad8(340), thd=128, _dev=torch.device("cpu"),
                               W=B200Qwen2p5VL.alloc_weights(tc, vc, torch.device("cpu")))
    sd = {k.replace("model.language_model.", "model.").replace("model.visual.", "visual."): v.detach() for k, v in ref.state_dict().items()}
    assert B200Qwen2p5VL.load_state_dict(me, sd).missing_keys == []
    for name in ("_rope_tables", "_decoder", "forward", "lm_logits"):
        setattr(me, name, types.MethodType(getattr(B200Qwen2p5VL, name), me))
    ids = torch.tensor([[1, 2, 3] + list(range(10, 30))])
    steps = 8
    with torch.no_grad():
        gen = ref.generate(input_ids=ids, attention_mask=torch.ones_like(ids), max_new_tokens=steps, min_new_tokens=steps, do_sample=False,
                           output_scores=True, return_dict_in_generate=True, repetition_penalty=1.0, eos_token_id=None, pad_token_id=0)
    want_tokens = gen.sequences[:, ids.shape[1]:]
    want_logits = torch.stack(gen.scores, dim=1).float()
    seq, scores = B200Qwen2p5VL.generate(me, ids, forced_tokens=want_tokens, output_scores=True, eos_token_id=(999999,))
    got = torch.stack(scores, dim=1)
    assert torch.equal(seq[:, ids.shape[1]:], want_tokens) and got.shape == want_logits.shape
    assert _rel(got, want_logits) < 2e-2
    top2 = want_logits.topk(2, dim=-1).values
    clear = (top2[..., 0] - top2[..., 1]) > 8 * (got - want_logits).abs().max()
    assert int(clear.sum()) >= 1 and torch.equal(got.argmax(-1)[clear], want_tokens[clear])      # same pick wherever it is not a tie
    # a left-padded batch of two prompts (what the processor produces for generation): each row must reproduce ITS OWN
    # unpadded reply's logits; right-padded batches are refused
    from gpt_image_edit_b200 import _lib
    short_prompt = torch.tensor([[7, 8, 9, 40, 41, 42, 43]])
    with torch.no_grad():
        gen_b = ref.generate(input_ids=short_prompt, attention_mask=torch.ones_like(short_prompt), max_new_tokens=steps,
                             min_new_tokens=steps, do_sample=False, output_scores=True, return_dict_in_generate=True,
                             repetition_penalty=1.0, eos_token_id=None, pad_token_id=0)
    tok_b = gen_b.sequences[:, short_prompt.shape[1]:]
    n, nb = ids.shape[1], short_prompt.shape[1]
    batch = torch.cat([ids, torch.cat([torch.zeros(1, n - nb, dtype=torch.long), short_prompt], dim=1)])
    mask = torch.cat([torch.ones(1, n, dtype=torch.long), torch.cat([torch.zeros(1, n - nb, dtype=torch.long),
                                                                     torch.ones(1, nb, dtype=torch.long)], dim=1)])
    seq2, scores2 = B200Qwen2p5VL.generate(me, batch, attention_mask=mask, forced_tokens=torch.cat([want_tokens, tok_b]),
                                           output_scores=True, eos_token_id=(999999,))
    got2 = torch.stack(scores2, dim=1)                                   # [2, steps, vocab]
    assert _rel(got2[0], want_logits[0]) < 2e-2
    assert _rel(got2[1], torch.stack(gen_b.scores, dim=1).float()[0]) < 2e-2
    with pytest.raises(_lib.B2FError, match="left-padded"):
        B200Qwen2p5VL.generate(me, batch.flip(1), attention_mask=mask.flip(1), max_new_tokens=2)
    # eos: the reply stops at the first eos token and the sequence ends there
    eos_tok = int(want_tokens[0, 1])
    short = B200Qwen2p5VL.generate(me, ids, forced_tokens=None, max_new_tokens=steps, eos_token_id=(eos_tok,))
    if bool(clear[0, :2].all()):
        assert short.shape[1] == ids.shape[1] + 2 and int(short[0, -1]) == eos_tok
