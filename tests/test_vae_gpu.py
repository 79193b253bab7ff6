"""GPU parity of the VAE kernels and of the composed encode/decode against the oracle restatement
of diffusers AutoencoderKL (oracle/vae_oracle.py, pinned against torchtitan's independent AE)."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def _rel_l2(a, b):
    return ((a.float() - b.float()).norm() / b.float().norm().clamp_min(1e-20)).item()


def _lib():
    from gpt_image_edit_b200 import _lib as L
    return L


@pytest.mark.parametrize("N,H,W,Cin,Cout,stride", [
    (1, 8, 16, 64, 128, 1), (2, 24, 40, 128, 128, 1), (1, 33, 50, 64, 256, 1),   # ragged spatial tiles
    (1, 64, 64, 256, 512, 1), (1, 32, 48, 128, 128, 2), (1, 64, 64, 512, 32, 1)])
def test_conv3x3_matches_torch(N, H, W, Cin, Cout, stride):
    L = _lib()
    g = torch.Generator(device="cuda").manual_seed(Cin + Cout + H)
    x = torch.randn(N, Cin, H, W, device="cuda", generator=g).bfloat16()
    w = (torch.randn(Cout, Cin, 3, 3, device="cuda", generator=g) * (1 / (9 * Cin)) ** 0.5).bfloat16()
    b = torch.randn(Cout, device="cuda", generator=g).bfloat16()
    x_nhwc = x.permute(0, 2, 3, 1).contiguous()
    w_ohwi = w.permute(0, 2, 3, 1).contiguous()
    if stride == 1:
        ref = F.conv2d(x.float(), w.float(), b.float(), padding=1)
    else:
        ref = F.conv2d(F.pad(x.float(), (0, 1, 0, 1)), w.float(), b.float(), stride=2)
    Ho, Wo = ref.shape[2:]
    out = torch.empty(N, Ho, Wo, Cout, device="cuda", dtype=torch.bfloat16)
    L.check(L.lib.b2f_conv3x3(L.ptr(x_nhwc), L.ptr(w_ohwi), L.ptr(b), L.ptr(out), None, N, H, W, Cin, Cout, stride, 0,
                              L.stream_ptr()), "conv")
    assert _rel_l2(out.permute(0, 3, 1, 2), ref) < 4e-3
    # residual epilogue (in place) and planar output
    res = torch.randn(N, Ho, Wo, Cout, device="cuda", generator=g).bfloat16()
    out2 = res.clone()
    L.check(L.lib.b2f_conv3x3(L.ptr(x_nhwc), L.ptr(w_ohwi), L.ptr(b), L.ptr(out2), L.ptr(out2), N, H, W, Cin, Cout,
                              stride, 0, L.stream_ptr()), "conv+res")
    ref2 = res.float() + ref.permute(0, 2, 3, 1).bfloat16().float()
    assert _rel_l2(out2, ref2) < 4e-3
    out3 = torch.empty(N, Cout, Ho, Wo, device="cuda", dtype=torch.bfloat16)
    L.check(L.lib.b2f_conv3x3(L.ptr(x_nhwc), L.ptr(w_ohwi), L.ptr(b), L.ptr(out3), None, N, H, W, Cin, Cout, stride, 1,
                              L.stream_ptr()), "conv nchw")
    assert torch.equal(out3, out.permute(0, 3, 1, 2))


@pytest.mark.parametrize("N,P,C,silu", [(1, 48, 32, 1), (2, 1000, 128, 1), (1, 4096, 512, 0), (1, 20000, 256, 1)])
def test_groupnorm_silu_matches_torch_chain(N, P, C, silu):
    L = _lib()
    g = torch.Generator(device="cuda").manual_seed(P)
    x = (torch.randn(N, P, C, device="cuda", generator=g) * 2 + 0.3).bfloat16()
    ga = (1 + 0.1 * torch.randn(C, device="cuda", generator=g)).bfloat16()
    be = (0.1 * torch.randn(C, device="cuda", generator=g)).bfloat16()
    y = torch.empty_like(x)
    stats = torch.empty(64 * N, device="cuda", dtype=torch.float64)
    L.check(L.lib.b2f_groupnorm_silu(L.ptr(x), L.ptr(ga), L.ptr(be), L.ptr(y), L.ptr(stats), N, P, C, 1e-6, silu,
                                     L.stream_ptr()), "gn")
    xc = x.permute(0, 2, 1)  # [N,C,P]
    r16 = F.group_norm(xc, 32, ga, be, eps=1e-6)
    r32 = F.group_norm(xc.float(), 32, ga.float(), be.float(), eps=1e-6)
    if silu:
        r16, r32 = F.silu(r16), F.silu(r32)
    r16, r32 = r16.permute(0, 2, 1), r32.permute(0, 2, 1)
    assert _rel_l2(y, r32) <= 1.5 * _rel_l2(r16, r32) + 1e-4


def test_small_vae_kernels():
    L = _lib()
    g = torch.Generator(device="cuda").manual_seed(3)
    x = torch.randn(2, 5, 7, 64, device="cuda", generator=g).bfloat16()
    up = torch.empty(2, 10, 14, 64, device="cuda", dtype=torch.bfloat16)
    L.check(L.lib.b2f_upsample2x(L.ptr(x), L.ptr(up), 2, 5, 7, 64, L.stream_ptr()), "up")
    ref = F.interpolate(x.permute(0, 3, 1, 2).float(), scale_factor=2.0, mode="nearest").permute(0, 2, 3, 1)
    assert torch.equal(up.float(), ref)
    img = torch.randn(2, 3, 9, 11, device="cuda", generator=g)
    o = torch.empty(2, 9, 11, 64, device="cuda", dtype=torch.bfloat16)
    L.check(L.lib.b2f_nchw_to_nhwc_pad(L.ptr(img), 1, L.ptr(o), 2, 3, 9, 11, 64, L.stream_ptr()), "pad")
    assert torch.equal(o[..., :3], img.permute(0, 2, 3, 1).bfloat16()) and o[..., 3:].abs().max() == 0
    s = (torch.randn(37, 1024, device="cuda", generator=g) * 3).bfloat16()
    ref = torch.softmax(s.float() * 0.25, dim=-1)
    L.check(L.lib.b2f_softmax_rows(L.ptr(s), 1024, 37, 1024, 0.25, L.stream_ptr()), "softmax")
    assert _rel_l2(s, ref) < 4e-3
    m = torch.randn(100, 72, device="cuda", generator=g).bfloat16()
    t = torch.empty(72, 100, device="cuda", dtype=torch.bfloat16)
    L.check(L.lib.b2f_transpose_bf16(L.ptr(m), 72, L.ptr(t), 100, 100, 72, L.stream_ptr()), "transpose")
    assert torch.equal(t, m.t())


def _vae_pair(boc, seed=0):
    from gpt_image_edit_b200.vae import B200AutoencoderKL, VaeConfig
    from oracle import vae_oracle as vo

    ocfg = vo.VaeConfig(block_out_channels=boc)
    sd = vo.make_synthetic_state_dict(ocfg, seed=seed, dtype=torch.bfloat16, device="cuda")
    vae = B200AutoencoderKL(VaeConfig(block_out_channels=boc))
    vae.load_state_dict(sd)
    # state_dict round trip speaks the diffusers layout
    back = vae.state_dict()
    assert all(torch.equal(back[k], sd[k]) for k in sd)
    return vo, ocfg, sd, vae


@pytest.mark.parametrize("boc,H,W", [((64, 128, 256, 256), 64, 96), ((128, 256, 512, 512), 128, 128),
                                     ((128, 256, 512, 512), 1024, 1024)],      # the last one: the FLUX VAE at the C1024 size
                         ids=["toy", "full-width-128", "c1024"])
def test_vae_encode_decode_match_oracle(boc, H, W):
    vo, ocfg, sd, vae = _vae_pair(boc)
    g = torch.Generator(device="cuda").manual_seed(5)
    x = (torch.rand(1, 3, H, W, device="cuda", generator=g) * 2 - 1).bfloat16()
    z = torch.randn(1, 16, H // 8, W // 8, device="cuda", generator=g).bfloat16()
    sd32 = {k: v.float() for k, v in sd.items()}
    mean = vae.encode(x).latent_dist.mode()
    m32, m16 = vo.encode_mode(sd32, ocfg, x.float()), vo.encode_mode(sd, ocfg, x)
    e_k, e_t = _rel_l2(mean, m32), _rel_l2(m16, m32)
    print(f"encode {boc[0]}: kernel-vs-fp32 {e_k:.3e}  torch-bf16-vs-fp32 {e_t:.3e}")
    assert e_k <= 2.0 * e_t + 3e-3
    img = vae.decode(z, return_dict=False)[0]
    i32, i16 = vo.decode(sd32, ocfg, z.float()), vo.decode(sd, ocfg, z)
    e_k, e_t = _rel_l2(img, i32), _rel_l2(i16, i32)
    print(f"decode {boc[0]}: kernel-vs-fp32 {e_k:.3e}  torch-bf16-vs-fp32 {e_t:.3e}")
    assert e_k <= 2.0 * e_t + 3e-3
    assert img.shape == (1, 3, H, W)


def test_vae_is_bit_reproducible_run_to_run():
    """GroupNorm statistics are combined in a fixed order (fp32 shared-memory atomics used to make the encoder, and
    with it the whole edit, differ in the last bf16 bit from run to run)."""
    from gpt_image_edit_b200.vae import B200AutoencoderKL, VaeConfig

    vae = B200AutoencoderKL(VaeConfig(block_out_channels=(64, 128, 256, 256))).randomize_(4)
    g = torch.Generator(device="cuda").manual_seed(0)
    img = (torch.rand(2, 3, 96, 160, device="cuda", generator=g) * 2 - 1).bfloat16()
    z = [vae.encode(img).latent_dist.mode() for _ in range(3)]
    assert torch.equal(z[0], z[1]) and torch.equal(z[0], z[2])
    d = [vae.decode(z[0], return_dict=False)[0] for _ in range(3)]
    assert torch.equal(d[0], d[1]) and torch.equal(d[0], d[2])


def test_uint8_pixels_in_and_out_equal_the_host_side_chain():
    """f3 (image I/O edges): uint8 [N,H,W,3] pixels go straight into encoder.conv_in's feeder kernel and come straight out
    of decoder.conv_out's epilogue; both must be bit-identical to the reference's host-side chain — `/255`, `(x-0.5)/0.5`,
    `.to(bf16)` on the way in (cli.py:99-116), `(x/2+0.5).clamp(0,1)`, `(*255).round().astype(uint8)` on the way out
    (VaeImageProcessor.postprocess, flux_pipeline.py:1130)."""
    from gpt_image_edit_b200.pipeline import VaeImageProcessor
    from gpt_image_edit_b200.vae import B200AutoencoderKL, VaeConfig

    vae = B200AutoencoderKL(VaeConfig(block_out_channels=(64, 128, 256, 256))).randomize_(seed=4)
    g = torch.Generator(device="cuda").manual_seed(2)
    u8 = torch.randint(0, 256, (2, 96, 128, 3), device="cuda", generator=g, dtype=torch.uint8)
    # the reference normalises on the HOST (torch CPU ops: a true division; CUDA's div-by-scalar multiplies by 1/255)
    ref_in = (((u8.cpu().permute(0, 3, 1, 2).float() / 255.0) - 0.5) / 0.5).cuda()
    m_u8 = vae.encode(u8).latent_dist.mean
    m_f32 = vae.encode(ref_in).latent_dist.mean
    m_bf16 = vae.encode(ref_in.bfloat16()).latent_dist.mean
    assert torch.equal(m_u8, m_f32) and torch.equal(m_u8, m_bf16)
    z = torch.randn(2, 16, 12, 16, device="cuda", generator=g).bfloat16()
    img = vae.decode(z, return_dict=False)[0]
    want = [np.asarray(p) for p in VaeImageProcessor.postprocess(img, "pil")]
    got = vae.decode_u8(z)
    assert got.dtype == torch.uint8 and tuple(got.shape) == (2, 96, 128, 3)
    assert all(np.array_equal(w, g_.cpu().numpy()) for w, g_ in zip(want, got))
