"""End-to-end parity of the sampling loop on the GPU: the product pipeline (libb2f transformer, VAE,
scheduler) against the oracle restatement of the reference loop on identical seeds, weights and
initial latents (noise is passed explicitly via `latents=`)."""
import pytest
import torch

pytestmark = pytest.mark.gpu

TOY = dict(num_layers=2, num_single_layers=2, attention_head_dim=128, num_attention_heads=2,
           joint_attention_dim=256, pooled_projection_dim=64)
BOC = (64, 128, 256, 256)


def _rel_l2(a, b):
    return ((a.float() - b.float()).norm() / b.float().norm().clamp_min(1e-20)).item()


def _build():
    from gpt_image_edit_b200.flux_transformer import B200FluxTransformer2DModel, FluxTransformerConfig
    from gpt_image_edit_b200.pipeline import FluxKontextPipeline
    from gpt_image_edit_b200.scheduler import FlowMatchEulerDiscreteScheduler
    from gpt_image_edit_b200.vae import B200AutoencoderKL, VaeConfig
    from oracle import flux_oracle as fo
    from oracle import vae_oracle as vo

    fcfg, vcfg = fo.FluxConfig(**TOY), vo.VaeConfig(block_out_channels=BOC)
    fsd = fo.make_synthetic_state_dict(fcfg, seed=3, dtype=torch.bfloat16, device="cuda")
    vsd = vo.make_synthetic_state_dict(vcfg, seed=4, dtype=torch.bfloat16, device="cuda")
    tr = B200FluxTransformer2DModel(FluxTransformerConfig(**TOY))
    tr.load_state_dict(fsd)
    vae = B200AutoencoderKL(VaeConfig(block_out_channels=BOC))
    vae.load_state_dict(vsd)
    pipe = FluxKontextPipeline(transformer=tr, vae=vae, scheduler=FlowMatchEulerDiscreteScheduler())
    return pipe, (fsd, fcfg, vsd, vcfg)


@pytest.mark.parametrize("B,H,W,steps", [(1, 128, 128, 4), (2, 128, 192, 3)])
def test_pipeline_matches_oracle_loop(B, H, W, steps):
    from oracle import pipeline_oracle as po

    pipe, (fsd, fcfg, vsd, vcfg) = _build()
    g = torch.Generator().manual_seed(7)
    image = (torch.randint(0, 256, (B, 3, H, W), generator=g).float() / 127.5 - 1.0).cuda()
    pe = torch.randn(B, 24, 256, generator=g).bfloat16().cuda()
    pooled = torch.randn(B, 64, generator=g).bfloat16().cuda()
    n_tgt = (H // 16) * (W // 16)
    noise = torch.stack([torch.randn(n_tgt, 64, generator=torch.Generator().manual_seed(42 + i)) for i in range(B)]).bfloat16().cuda()

    lat = pipe(image=image, prompt_embeds=pe, pooled_prompt_embeds=pooled, height=H, width=W, num_inference_steps=steps,
               guidance_scale=3.5, latents=noise.clone(), max_area=H * W, _auto_resize=False, output_type="latent").images
    img = pipe(image=image, prompt_embeds=pe, pooled_prompt_embeds=pooled, height=H, width=W, num_inference_steps=steps,
               guidance_scale=3.5, latents=noise.clone(), max_area=H * W, _auto_resize=False, output_type="pt").images
    assert pipe.scheduler.step_index == steps          # integer step bookkeeping

    kw = dict(height=H, width=W, num_inference_steps=steps, guidance_scale=3.5, max_area=H * W)
    ref_lat = po.sample(fsd, fcfg, vsd, vcfg, image.bfloat16(), pe, pooled, latents=noise.clone(), output="latent", **kw)
    ref_img = po.sample(fsd, fcfg, vsd, vcfg, image.bfloat16(), pe, pooled, latents=noise.clone(), output="image", **kw)
    e_lat = _rel_l2(lat, ref_lat)
    e_img = _rel_l2(img, (ref_img.float() / 2 + 0.5).clamp(0, 1))
    print(f"B={B} {H}x{W} {steps} steps: latents rel-L2 vs torch-bf16 oracle {e_lat:.3e}; decoded image {e_img:.3e}")
    # bf16 tolerance for the whole trajectory: the two bf16 paths differ by re-association only
    assert e_lat < 2e-2
    assert e_img < 3e-2
    assert img.shape == (B, 3, H, W) and torch.isfinite(img).all()


def test_pipeline_pil_output_and_reference_size_rule():
    pipe, _ = _build()
    g = torch.Generator().manual_seed(1)
    image = (torch.rand(1, 3, 128, 128, generator=g) * 2 - 1).cuda()
    pe = torch.randn(1, 24, 256, generator=g).bfloat16().cuda()
    pooled = torch.randn(1, 64, generator=g).bfloat16().cuda()
    out = pipe(image=image, prompt_embeds=pe, pooled_prompt_embeds=pooled, height=128, width=128, num_inference_steps=2,
               max_area=128 * 128, _auto_resize=False, generator=torch.Generator(device="cuda").manual_seed(42)).images
    assert len(out) == 1 and out[0].size == (128, 128)


def test_true_cfg_second_forward_matches_oracle_loop():
    """true_cfg_scale > 1 with negative embeddings (reference flux_pipeline.py:925-957, 1080-1095): a second forward per
    step on the negative prompt, `neg + s * (pos - neg)`.  The oracle loop for this case is bit-identical to the
    reference's own __call__ (tests/golden/pipeline_ref_loop.pt: case_true_cfg)."""
    from oracle import pipeline_oracle as po

    pipe, (fsd, fcfg, vsd, vcfg) = _build()
    g = torch.Generator().manual_seed(13)
    H = W = 128
    image = (torch.randint(0, 256, (1, 3, H, W), generator=g).float() / 127.5 - 1.0).cuda()
    pe = torch.randn(1, 24, 256, generator=g).bfloat16().cuda()
    pooled = torch.randn(1, 64, generator=g).bfloat16().cuda()
    npe = torch.randn(1, 9, 256, generator=g).bfloat16().cuda()          # a negative prompt of another length
    npooled = torch.randn(1, 64, generator=g).bfloat16().cuda()
    noise = torch.randn(1, 64, 64, generator=g).bfloat16().cuda()
    common = dict(height=H, width=W, num_inference_steps=3, guidance_scale=3.5, max_area=H * W)
    lat = pipe(image=image, prompt_embeds=pe, pooled_prompt_embeds=pooled, negative_prompt_embeds=npe,
               negative_pooled_prompt_embeds=npooled, true_cfg_scale=6.0, latents=noise.clone(), _auto_resize=False,
               output_type="latent", **common).images
    ref = po.sample(fsd, fcfg, vsd, vcfg, image.bfloat16(), pe, pooled, latents=noise.clone(), output="latent",
                    true_cfg_scale=6.0, negative_prompt_embeds=npe, negative_pooled=npooled, **common)
    plain = pipe(image=image, prompt_embeds=pe, pooled_prompt_embeds=pooled, latents=noise.clone(), _auto_resize=False,
                 output_type="latent", **common).images
    e = _rel_l2(lat, ref)
    print(f"true-CFG latents rel-L2 vs torch-bf16 oracle {e:.3e}; distance from the plain run {_rel_l2(lat, plain):.3e}")
    assert e < 2e-2
    assert _rel_l2(lat, plain) > 3 * e                                  # the guidance really changed the trajectory
    # scale 1 or no negative prompt: the second forward is skipped (reference :925-928)
    same = pipe(image=image, prompt_embeds=pe, pooled_prompt_embeds=pooled, negative_prompt_embeds=npe,
                negative_pooled_prompt_embeds=npooled, true_cfg_scale=1.0, latents=noise.clone(), _auto_resize=False,
                output_type="latent", **common).images
    assert torch.equal(same, plain)


def test_c1024_full_edit_trajectory_matches_oracle_loop():
    """The whole flagship job — FLUX VAE encode of a 1024x1024 context image, 28 Euler steps of the full 19+38-block
    MMDiT at S = 8736, VAE decode — against the oracle loop (bit-identical to the reference's own __call__, see
    tests/golden/pipeline_ref_loop.pt) run in torch bf16 on the same synthetic weights, inputs and initial noise."""
    from gpt_image_edit_b200.flux_transformer import B200FluxTransformer2DModel, FluxTransformerConfig
    from gpt_image_edit_b200.pipeline import FluxKontextPipeline
    from gpt_image_edit_b200.scheduler import FlowMatchEulerDiscreteScheduler
    from gpt_image_edit_b200.vae import B200AutoencoderKL, VaeConfig
    from oracle import flux_oracle as fo
    from oracle import pipeline_oracle as po
    from oracle import vae_oracle as vo

    fcfg, vcfg = fo.FluxConfig(), vo.VaeConfig()
    fsd = fo.make_synthetic_state_dict(fcfg, seed=3, dtype=torch.bfloat16, device="cuda")
    vsd = vo.make_synthetic_state_dict(vcfg, seed=4, dtype=torch.bfloat16, device="cuda")
    tr = B200FluxTransformer2DModel(FluxTransformerConfig())
    tr.load_state_dict(fsd)
    vae = B200AutoencoderKL(VaeConfig())
    vae.load_state_dict(vsd)
    pipe = FluxKontextPipeline(transformer=tr, vae=vae, scheduler=FlowMatchEulerDiscreteScheduler())
    g = torch.Generator().manual_seed(21)
    H = W = 1024
    image = (torch.randint(0, 256, (1, 3, H, W), generator=g).float() / 127.5 - 1.0).cuda()
    pe = torch.randn(1, 544, 4096, generator=g).bfloat16().cuda()
    pooled = torch.randn(1, 768, generator=g).bfloat16().cuda()
    noise = torch.randn(1, 4096, 64, generator=torch.Generator().manual_seed(42)).bfloat16().cuda()
    kw = dict(height=H, width=W, num_inference_steps=28, guidance_scale=3.5, max_area=H * W)
    lat = pipe(image=image, prompt_embeds=pe, pooled_prompt_embeds=pooled, latents=noise.clone(), _auto_resize=False,
               output_type="latent", **kw).images
    img = pipe(image=image, prompt_embeds=pe, pooled_prompt_embeds=pooled, latents=noise.clone(), _auto_resize=False,
               output_type="pt", **kw).images
    del pipe, tr
    torch.cuda.empty_cache()
    steps_seen = []
    ref_img = po.sample(fsd, fcfg, vsd, vcfg, image.bfloat16(), pe, pooled, latents=noise.clone(), output="image",
                        callback=lambda i, x: steps_seen.append(x.clone()) if i == 27 else None, **kw)
    e_lat = _rel_l2(lat, steps_seen[-1])
    e_img = _rel_l2(img, (ref_img.float() / 2 + 0.5).clamp(0, 1))
    print(f"C1024, 28 steps: final latents rel-L2 vs torch-bf16 oracle loop {e_lat:.3e}; decoded image {e_img:.3e}")
    assert lat.shape == (1, 4096, 64) and img.shape == (1, 3, H, W) and torch.isfinite(img).all()
    assert e_lat < 5e-2 and e_img < 5e-2
