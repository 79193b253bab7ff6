"""GPU parity of the HBM-bound fused kernels against the torch-eager op chains they replace
(the chains are written exactly as diffusers executes them)."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def _rel_l2(a, b):
    return ((a.float() - b.float()).norm() / b.float().norm().clamp_min(1e-20)).item()


@pytest.mark.parametrize("B,rows,D", [(1, 33, 256), (2, 300, 3072), (1, 1000, 3584), (1, 8736, 3072)])
def test_ln_modulate_matches_eager_chain(B, rows, D):
    from gpt_image_edit_b200 import ops

    g = torch.Generator(device="cuda").manual_seed(rows)
    x = (torch.randn(B, rows, D, device="cuda", generator=g) * 3 + 0.5).bfloat16()
    mod = torch.randn(B, 2 * D, device="cuda", generator=g).bfloat16()
    scale, shift = mod[:, :D], mod[:, D:]
    out = ops.ln_modulate(x, scale, shift)
    ref_bf16 = F.layer_norm(x, (D,), None, None, 1e-6) * (1 + scale[:, None]) + shift[:, None]
    ref_f32 = F.layer_norm(x.float(), (D,), None, None, 1e-6) * (1 + scale.float()[:, None]) + shift.float()[:, None]
    # identical rounding chain -> nearly bit-identical to torch's bf16 result
    mism = (out != ref_bf16).float().mean().item()
    assert mism < 2e-3, f"{mism:.4%} elements differ from the torch bf16 chain"
    assert _rel_l2(out, ref_f32) <= 1.5 * _rel_l2(ref_bf16, ref_f32) + 1e-4


def test_rmsnorm_rope_matches_eager_chain():
    from gpt_image_edit_b200 import ops
    from oracle import flux_oracle as fo

    B, S_txt, S_img, H = 2, 40, 200, 3
    S = S_txt + S_img
    g = torch.Generator(device="cuda").manual_seed(1)
    qkv = torch.randn(B, S, 3 * H * 128, device="cuda", generator=g).bfloat16()
    wq, wk, wqa, wka = ((1 + 0.1 * torch.randn(128, device="cuda", generator=g)).bfloat16() for _ in range(4))
    ids = torch.zeros(S, 3, device="cuda")
    ids[S_txt:, 0] = (torch.arange(S_img, device="cuda") >= 100).float()
    ids[S_txt:, 1] = (torch.arange(S_img, device="cuda") % 100) // 10
    ids[S_txt:, 2] = torch.arange(S_img, device="cuda") % 10
    cos, sin = ops.rope_tables(ids)
    cos_ref, sin_ref = fo.rope_tables(ids)
    assert torch.allclose(cos, cos_ref, atol=1e-6) and torch.allclose(sin, sin_ref, atol=1e-6)

    ref = qkv.clone()

    def chain(x, w_txt, w_img):  # x [B,S,H*128]
        xh = x.view(B, S, H, 128).transpose(1, 2)
        y = torch.cat([fo.rms_norm(xh[:, :, :S_txt], w_txt), fo.rms_norm(xh[:, :, S_txt:], w_img)], dim=2)
        return fo.apply_rotary_emb(y, cos_ref, sin_ref).transpose(1, 2).reshape(B, S, H * 128)

    ref[:, :, : H * 128] = chain(qkv[:, :, : H * 128], wqa, wq)
    ref[:, :, H * 128 : 2 * H * 128] = chain(qkv[:, :, H * 128 : 2 * H * 128], wka, wk)
    ops.rmsnorm_rope_(qkv, H, wq, wk, cos, sin, wq_added=wqa, wk_added=wka, n_added=S_txt)
    assert torch.equal(qkv[:, :, 2 * H * 128 :], ref[:, :, 2 * H * 128 :])  # V untouched
    mism = (qkv != ref).float().mean().item()
    assert mism < 5e-3, f"{mism:.4%} elements differ"
    assert _rel_l2(qkv, ref) < 2e-3


def test_euler_step_bit_exact():
    from gpt_image_edit_b200 import ops

    g = torch.Generator(device="cuda").manual_seed(2)
    x = torch.randn(2, 4096, 64, device="cuda", generator=g).bfloat16()
    v_all = torch.randn(2, 8192, 64, device="cuda", generator=g).bfloat16()
    v = v_all[:, :4096]
    dt = torch.tensor(-0.0357, device="cuda", dtype=torch.float32)
    ref = (x.float() + dt * v).to(torch.bfloat16)  # 0-dim fp32 * bf16 tensor -> bf16 product (A.5)
    xv = x.clone()
    ops.euler_step_(xv[0], v[0], float(dt))
    ops.euler_step_(xv[1], v[1], float(dt))
    assert torch.equal(xv, ref)


def test_blend_matches_torch_bf16_expression():
    """`old * (1 - f) + image_embeds * f` (reference modeling_univa_qwen2p5vl.py:505) evaluated by torch on bf16 CUDA
    tensors vs b2f_blend_bf16: bit-identical."""
    from gpt_image_edit_b200 import ops

    g = torch.Generator(device="cuda").manual_seed(5)
    old = torch.randn(256, 3584, device="cuda", generator=g).bfloat16()
    emb = torch.randn(256, 3584, device="cuda", generator=g).bfloat16()
    for f in (0.3, 0.5, 0.85):
        want = old * (1 - f) + emb * f
        got = ops.blend(old, emb, 1 - f, f)
        assert torch.equal(want, got), f
