"""Debug: per-stage error of the libb2f Qwen decoder vs transformers (fp32 and bf16) on the toy model."""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[2]))
from transformers.models.qwen2_5_vl import modeling_qwen2_5_vl as hf  # noqa: E402
from transformers.models.qwen2_5_vl.configuration_qwen2_5_vl import Qwen2_5_VLConfig  # noqa: E402

from gpt_image_edit_b200 import ops  # noqa: E402
from gpt_image_edit_b200.qwen2p5vl import B200Qwen2p5VL, QwenTextConfig, QwenVisionConfig, get_rope_index  # noqa: E402

IMG, VSTART = 900, 902
rel = lambda a, b: ((a.float() - b.float()).norm() / b.float().norm()).item()
tc = dict(hidden_size=256, num_hidden_layers=2, num_attention_heads=2, num_key_value_heads=1, intermediate_size=512, vocab_size=1000, rms_norm_eps=1e-6)
vcfg = dict(depth=2, hidden_size=256, num_heads=4, intermediate_size=340, out_hidden_size=256, fullatt_block_indexes=[1])
cfg = Qwen2_5_VLConfig(text_config=dict(tc, rope_parameters=dict(rope_type="default", rope_theta=1e6, mrope_section=[16, 24, 24])),
                       vision_config=vcfg, image_token_id=IMG, video_token_id=901, vision_start_token_id=VSTART, vision_end_token_id=903)
torch.manual_seed(0)
ref = hf.Qwen2_5_VLModel(cfg).eval()
with torch.no_grad():
    for p in ref.parameters():
        if p.dim() == 1:
            p.add_(0.05 * torch.randn_like(p))
sd = {k.replace("language_model.", "model."): v.detach().to(torch.bfloat16) for k, v in ref.state_dict().items()}
mine = B200Qwen2p5VL(QwenTextConfig(**tc, image_token_id=IMG, video_token_id=901, vision_start_token_id=VSTART), QwenVisionConfig(**vcfg))
mine.load_state_dict(sd)
ids = torch.tensor([[1, 2, 3] + list(range(10, 60))]).cuda()      # text only


def run(model):
    caps = {}
    hooks = []
    lm = model.language_model
    for i, layer in enumerate(lm.layers):
        hooks.append(layer.self_attn.register_forward_hook(lambda m, a, o, i=i: caps.__setitem__(f"attn{i}", o[0].detach())))
        hooks.append(layer.mlp.register_forward_hook(lambda m, a, o, i=i: caps.__setitem__(f"mlp{i}", o.detach())))
        hooks.append(layer.register_forward_hook(lambda m, a, o, i=i: caps.__setitem__(f"layer{i}", (o[0] if isinstance(o, tuple) else o).detach())))
        hooks.append(layer.input_layernorm.register_forward_hook(lambda m, a, o, i=i: caps.__setitem__(f"ln1_{i}", o.detach())))
        hooks.append(layer.self_attn.q_proj.register_forward_hook(lambda m, a, o, i=i: caps.__setitem__(f"q{i}", o.detach())))
        hooks.append(layer.self_attn.o_proj.register_forward_hook(lambda m, a, o, i=i: caps.__setitem__(f"oin{i}", a[0].detach())))
    with torch.no_grad():
        caps["out"] = model(input_ids=ids).last_hidden_state
    for h in hooks:
        h.remove()
    return caps


c16 = run(ref.to("cuda", torch.bfloat16))
r32 = ref.to("cuda", torch.float32)
with torch.no_grad():
    for p in r32.parameters():
        p.copy_(p.bfloat16().float())
c32 = run(r32)

# my decoder, stage by stage (mirrors B200Qwen2p5VL._decoder)
W, t = mine.W, mine.tc
B, L = ids.shape
x = ops.gather_rows(W["model.embed_tokens"], ids.reshape(-1).contiguous())
pos, _ = get_rope_index(ids, None, None, spatial_merge_size=2, image_token_id=IMG, vision_start_token_id=VSTART)
cos, sin = mine._rope_tables(pos)
hd, nq, nkv = mine.thd, t.num_attention_heads, t.num_key_value_heads
for i in range(t.num_hidden_layers):
    p = f"model.layers.{i}."
    xn = ops.rmsnorm(x, W[p + "ln1"], eps=t.rms_norm_eps)
    print(f"L{i} ln1   mine {rel(xn, c32[f'ln1_{i}'][0]):.3e}  hf16 {rel(c16[f'ln1_{i}'], c32[f'ln1_{i}']):.3e}")
    qkv = ops.linear(xn, W[p + "qkv.w"], W[p + "qkv.b"])
    print(f"L{i} q     mine {rel(qkv[:, :nq * hd], c32[f'q{i}'][0]):.3e}  hf16 {rel(c16[f'q{i}'], c32[f'q{i}']):.3e}")
    ops.rope_half_(qkv, nq + nkv, hd, cos, sin, fp32_math=False)
    q = qkv[:, : nq * hd].unflatten(1, (nq, hd)).unflatten(0, (B, L))
    k = qkv[:, nq * hd: (nq + nkv) * hd].unflatten(1, (nkv, hd)).unflatten(0, (B, L))
    v = qkv[:, (nq + nkv) * hd:].unflatten(1, (nkv, hd)).unflatten(0, (B, L))
    o = ops.attention(q, k, v, causal=True)
    print(f"L{i} attn-core mine {rel(o[0], c32[f'oin{i}'][0]):.3e}  hf16 {rel(c16[f'oin{i}'], c32[f'oin{i}']):.3e}")
    ao = ops.linear(o.view(B * L, -1), W[p + "o.w"])
    print(f"L{i} attn  mine {rel(ao, c32[f'attn{i}'][0]):.3e}  hf16 {rel(c16[f'attn{i}'], c32[f'attn{i}']):.3e}")
    x = (x.float() + ao.float()).bfloat16()
    xn = ops.rmsnorm(x, W[p + "ln2"], eps=t.rms_norm_eps)
    a = ops.swiglu(ops.linear(xn, W[p + "gu.w"]), t.intermediate_size)
    mo = ops.linear(a, W[p + "down.w"])
    print(f"L{i} mlp   mine {rel(mo, c32[f'mlp{i}'][0]):.3e}  hf16 {rel(c16[f'mlp{i}'], c32[f'mlp{i}']):.3e}")
    x = (x.float() + mo.float()).bfloat16()
    print(f"L{i} out   mine {rel(x, c32[f'layer{i}'][0]):.3e}  hf16 {rel(c16[f'layer{i}'], c32[f'layer{i}']):.3e}")
h = ops.rmsnorm(x, W["model.norm"], eps=t.rms_norm_eps)
print(f"final mine {rel(h, c32['out'][0]):.3e} hf16 {rel(c16['out'], c32['out']):.3e}  full-forward mine {rel(mine(ids), c32['out']):.3e}")
