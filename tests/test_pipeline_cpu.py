"""The PRODUCT sampling loop (gpt_image_edit_b200/pipeline.py: FluxKontextPipeline.__call__) against the outputs of the
reference's own `__call__` (tests/golden/pipeline_ref_call.pt, written by tests/golden/make_pipeline_call_golden.py from
/root/reference/univa/utils/flux_pipeline.py).  Both run over the same protocol objects — oracle-backed transformer /
VAE / scheduler adapters on the CPU in fp32 — so every difference would be host logic: size rule, noise draw, latent and
id layout, [target ‖ context] concat, sigma schedule inputs, guidance vector, timestep / 1000, callbacks, interrupt,
VAE affine, postprocess.  The same fp32 torch ops run on both sides, so the results agree to fp32 round-off (bit-equal at
the thread count the fixture was made with; a different intra-op thread count reorders a few sums: <= 1e-5 allowed);
everything discrete — forwards, token counts, timesteps, ids, callback traffic — must be identical."""
import sys
from pathlib import Path

import pytest
import torch

GOLDEN = Path(__file__).parent / "golden"
sys.path.insert(0, str(GOLDEN))


@pytest.fixture(scope="module")
def fx():
    return torch.load(GOLDEN / "pipeline_ref_call.pt", weights_only=False)


def _cases():
    import make_pipeline_call_golden as mk
    return mk.CASES


@pytest.mark.parametrize("name", _cases())
def test_product_call_equals_the_references_call(fx, name):
    import make_pipeline_call_golden as mk

    from gpt_image_edit_b200.pipeline import FluxKontextPipeline

    want = fx[name]
    got = mk.run_case(FluxKontextPipeline, name)
    assert got["n_forwards"] == want["n_forwards"] and got["n_tokens"] == want["n_tokens"]
    assert torch.equal(got["timesteps"], want["timesteps"])
    assert (got["guidance"] is None) == (want["guidance"] is None)
    if want["guidance"] is not None:
        assert torch.equal(got["guidance"], want["guidance"])
    assert torch.equal(got["img_ids"], want["img_ids"]) and torch.equal(got["txt_ids"], want["txt_ids"])
    assert got["callback_log"] == want["callback_log"]
    assert got["num_timesteps"] == want["num_timesteps"] and got["current_timestep"] is None is want["current_timestep"]
    assert got["images"].shape == want["images"].shape and got["images"].dtype == want["images"].dtype
    err = (got["images"] - want["images"]).abs().max().item()
    assert err <= 1e-5 * max(1.0, want["images"].abs().max().item()), (name, err)


def test_fixture_covers_the_behaviours_it_names(fx):
    assert fx["interrupt"]["n_forwards"] == 1 and fx["callback"]["n_forwards"] == 3
    assert fx["callback"]["callback_log"][0]["keys"] == ["latents", "prompt_embeds"]
    assert fx["no_guidance"]["guidance"] is None and torch.all(fx["callback"]["guidance"] == 2.0)
    assert fx["auto_resize"]["n_tokens"][0] == 16 + (1248 // 16) * (832 // 16)         # context at (w, h) = (1248, 832)
    assert fx["text_to_image"]["n_tokens"][0] == 24 and fx["floor_size"]["images"].shape[1] == 24
    assert fx["decode_pt"]["images"].shape == (1, 3, 64, 64) and 0 <= fx["decode_pt"]["images"].min()
    assert not torch.equal(fx["gen_list"]["images"][0], fx["gen_list"]["images"][1])


def test_randn_tensor_and_caller_latents():
    from gpt_image_edit_b200.pipeline import FluxKontextPipeline, randn_tensor

    gens = [torch.Generator().manual_seed(s) for s in (5, 6)]
    want = torch.cat([torch.randn((1, 4, 2, 2), generator=torch.Generator().manual_seed(s)) for s in (5, 6)])
    assert torch.equal(randn_tensor((2, 4, 2, 2), generator=gens), want)
    one = randn_tensor((2, 4, 2, 2), generator=[torch.Generator().manual_seed(5)])         # a list of one is that one
    assert torch.equal(one, torch.randn((2, 4, 2, 2), generator=torch.Generator().manual_seed(5)))
    pipe = FluxKontextPipeline(transformer=None)
    with pytest.raises(ValueError):
        pipe.prepare_latents(None, 3, 16, 64, 64, torch.float32, "cpu", generator=gens)
    mine = torch.randn(1, 16, 64)
    keep = mine.clone()
    latents, _, ids, _ = pipe.prepare_latents(None, 1, 16, 64, 64, torch.float32, "cpu", latents=mine)
    latents.mul_(0)                                       # the Euler kernel works in place on the pipeline's own copy
    assert torch.equal(mine, keep) and ids.shape == (16, 3)


def test_image_processor_accepts_what_the_reference_pipeline_accepts():
    """PIL / array / tensor / list inputs of `pipe(image=...)` (flux_pipeline.py:959-973 hands them to VaeImageProcessor)."""
    import numpy as np
    from PIL import Image

    from gpt_image_edit_b200.pipeline import VaeImageProcessor

    ip = VaeImageProcessor(vae_scale_factor=16)
    rng = np.random.default_rng(0)
    u8 = rng.integers(0, 256, (70, 100, 3), dtype=np.uint8)
    pil = Image.fromarray(u8)
    assert ip.get_default_height_width(pil) == (64, 96) and ip.get_default_height_width([pil, pil]) == (64, 96)
    x = ip.preprocess(ip.resize(pil, 64, 96), 64, 96)
    assert x.shape == (1, 3, 64, 96) and x.dtype == torch.float32 and -1 <= x.min() and x.max() <= 1 and x.min() < 0
    same = ip.preprocess(Image.fromarray(u8[:64, :96]))                      # no resize needed: exact pixels, (u/255)*2-1
    assert torch.equal(same[0], torch.from_numpy(u8[:64, :96].astype(np.float32) / 255.0).permute(2, 0, 1) * 2.0 - 1.0)
    t01 = torch.rand(2, 3, 32, 48)
    assert torch.equal(ip.preprocess(t01), 2.0 * t01 - 1.0)
    t11 = t01 * 2 - 1
    assert ip.preprocess(t11) is t11 or torch.equal(ip.preprocess(t11), t11)           # already [-1, 1]: passes through
    assert torch.equal(ip.resize(t01, 64, 96), torch.nn.functional.interpolate(t01, size=(64, 96)))
    arr = rng.random((32, 48, 3), dtype=np.float32)
    assert torch.equal(ip.preprocess(arr)[0], torch.from_numpy(arr).permute(2, 0, 1) * 2.0 - 1.0)
    assert ip.preprocess([t01[0], t01[1]]).shape == (2, 3, 32, 48)
    img = torch.rand(2, 3, 8, 8) * 2 - 1
    pt = ip.postprocess(img, "pt")
    assert torch.equal(pt, (img * 0.5 + 0.5).clamp(0, 1))
    assert np.array_equal(ip.postprocess(img, "np"), pt.permute(0, 2, 3, 1).numpy())
    pils = ip.postprocess(img, "pil")
    assert np.array_equal(np.array(pils[1]), (pt[1].permute(1, 2, 0).numpy() * 255).round().astype("uint8"))
    assert ip.postprocess(img, "latent") is img
    with pytest.raises(ValueError):
        ip.postprocess(img, "jpeg")
