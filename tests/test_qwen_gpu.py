"""GPU parity of the Qwen2.5-VL prefill on libb2f kernels against transformers' Qwen2_5_VLModel
(the classes the reference subclasses; transformers 5.5 is the copy in this image) with identical
random weights, at toy widths that keep every structural feature: windowed + full ViT attention,
head_dim zero-padding (64 -> 128), padded MLP width (340 -> 344), 2x2 patch merger, masked scatter
of image embeddings, M-RoPE, causal GQA."""
import pytest
import torch

pytestmark = pytest.mark.gpu
hf = pytest.importorskip("transformers.models.qwen2_5_vl.modeling_qwen2_5_vl")

IMG, VSTART = 900, 902


def _rel_l2(a, b):
    return ((a.float() - b.float()).norm() / b.float().norm().clamp_min(1e-20)).item()


def _models():
    from transformers.models.qwen2_5_vl.configuration_qwen2_5_vl import Qwen2_5_VLConfig

    from gpt_image_edit_b200.qwen2p5vl import B200Qwen2p5VL, QwenTextConfig, QwenVisionConfig

    tc = dict(hidden_size=256, num_hidden_layers=2, num_attention_heads=2, num_key_value_heads=1, intermediate_size=512,
              vocab_size=1000)
    vcfg = dict(depth=3, hidden_size=256, num_heads=4, intermediate_size=340, out_hidden_size=256,
                fullatt_block_indexes=[1])
    cfg = Qwen2_5_VLConfig(
        text_config=dict(tc, rope_parameters=dict(rope_type="default", rope_theta=1e6, mrope_section=[16, 24, 24])),
        vision_config=vcfg, image_token_id=IMG, video_token_id=901, vision_start_token_id=VSTART, vision_end_token_id=903)
    torch.manual_seed(0)
    ref = hf.Qwen2_5_VLModel(cfg).eval()
    with torch.no_grad():
        for p in ref.parameters():          # non-trivial norm weights / biases everywhere
            if p.dim() == 1:
                p.add_(0.05 * torch.randn_like(p))
    sd = {}
    for k, v in ref.state_dict().items():
        k = k.replace("language_model.", "model.")     # transformers-5 layout -> the 4.50 names of the checkpoint
        sd[k] = v.detach()
    mine = B200Qwen2p5VL(QwenTextConfig(**tc, image_token_id=IMG, video_token_id=901, vision_start_token_id=VSTART),
                         QwenVisionConfig(**vcfg))
    mine.load_state_dict({k: v.to(torch.bfloat16) for k, v in sd.items()})
    return ref, mine


def test_vision_tower_and_prefill_match_transformers():
    ref, mine = _models()
    g = torch.Generator().manual_seed(3)
    grid = torch.tensor([[1, 16, 8]])                       # 128 patches -> 32 merged tokens, two 4x4 windows
    pix = torch.randn(128, 1176, generator=g)
    ids = torch.tensor([[1, 2, 3, VSTART] + [IMG] * 32 + [903] + list(range(10, 31))])
    ref16 = ref.to("cuda", torch.bfloat16)
    with torch.no_grad():
        v16 = ref16.visual(pix.cuda().bfloat16(), grid_thw=grid.cuda()).pooler_output
        h16 = ref16(input_ids=ids.cuda(), pixel_values=pix.cuda().bfloat16(), image_grid_thw=grid.cuda()).last_hidden_state
    ref32 = ref.to("cuda", torch.float32)
    with torch.no_grad():
        # the checker's fp32 run uses the same bf16-rounded weights
        for p in ref32.parameters():
            p.copy_(p.bfloat16().float())
        v32 = ref32.visual(pix.cuda().bfloat16().float(), grid_thw=grid.cuda()).pooler_output
        h32 = ref32(input_ids=ids.cuda(), pixel_values=pix.cuda().bfloat16().float(), image_grid_thw=grid.cuda()).last_hidden_state
    v = mine.forward_visual(pix.cuda().bfloat16(), grid)
    h = mine(ids.cuda(), pixel_values=pix.cuda().bfloat16(), image_grid_thw=grid)
    ev_k, ev_t = _rel_l2(v, v32), _rel_l2(v16, v32)
    eh_k, eh_t = _rel_l2(h, h32[0:1]), _rel_l2(h16, h32)
    print(f"ViT: kernel-vs-fp32 {ev_k:.3e} torch-bf16-vs-fp32 {ev_t:.3e} | prefill: kernel {eh_k:.3e} torch-bf16 {eh_t:.3e}")
    assert v.shape == v32.shape and h.shape == h32.shape
    assert ev_k <= 2.0 * ev_t + 3e-3
    assert eh_k <= 2.0 * eh_t + 3e-3


def test_llm_kernels_match_eager_chains():
    from gpt_image_edit_b200 import ops

    g = torch.Generator(device="cuda").manual_seed(1)
    x = (torch.randn(300, 3584, device="cuda", generator=g) * 2).bfloat16()
    w = (1 + 0.1 * torch.randn(3584, device="cuda", generator=g)).bfloat16()
    y = ops.rmsnorm(x, w)
    xf = x.float()
    ref = w * (xf * torch.rsqrt(xf.pow(2).mean(-1, keepdim=True) + 1e-6)).to(torch.bfloat16)
    assert (y != ref).float().mean().item() < 2e-3
    gu = torch.randn(200, 2 * 344, device="cuda", generator=g).bfloat16()
    out = ops.swiglu(gu, 344)
    ref = torch.nn.functional.silu(gu[:, :344]) * gu[:, 344:]
    assert (out != ref).float().mean().item() < 2e-3
    table = torch.randn(50, 256, device="cuda", generator=g).bfloat16()
    idx = torch.tensor([3, 3, 49, 0, 17], device="cuda")
    assert torch.equal(ops.gather_rows(table, idx), table[idx])
    dst = torch.zeros(20, 256, device="cuda", dtype=torch.bfloat16)
    ops.scatter_rows_(dst, torch.tensor([5, 1, 19], device="cuda"), table[:3].contiguous())
    assert torch.equal(dst[5], table[0]) and torch.equal(dst[19], table[2]) and dst[0].abs().max() == 0
    # erf-GELU epilogue (patch merger)
    a = torch.randn(130, 128, device="cuda", generator=g).bfloat16()
    wl = (torch.randn(264, 128, device="cuda", generator=g) * 0.1).bfloat16()
    b = torch.randn(264, device="cuda", generator=g).bfloat16()
    o = ops.linear(a, wl, b, epilogue=ops.EPI_GELU_ERF)
    r = torch.nn.functional.gelu((a.float() @ wl.float().t() + b.float()).bfloat16().float())
    assert _rel_l2(o, r) < 6e-3
