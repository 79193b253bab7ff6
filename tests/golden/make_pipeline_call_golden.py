"""Generates tests/golden/pipeline_ref_call.pt: the outputs of THE REFERENCE'S OWN `FluxKontextPipeline.__call__`
(/root/reference/univa/utils/flux_pipeline.py, executed from its file as make_pipeline_ref_golden.py does) over the
argument surface the product pipeline (gpt_image_edit_b200/pipeline.py) re-hosts: noise drawn from one generator or
a list of generators, step callbacks that read and replace `latents` / `prompt_embeds`, `interrupt`, a transformer
without the guidance embedder, caller-given sigmas, default and non-multiple-of-16 sizes, `_auto_resize` onto the
preferred Kontext resolutions, decode + postprocess, text-to-image (no context image), a context given as latents and
one context image shared by a batch of prompts.

tests/test_pipeline_cpu.py runs the PRODUCT pipeline on the same protocol objects (oracle-backed transformer / VAE /
scheduler adapters, CPU, fp32) and requires the same tensors, so the product's host loop is pinned to the reference's
statements without a GPU.

Third-party behaviour the reference file reaches and that is not on disk is restated in the stand-ins below and
marked [dep-spec] (diffusers 0.32.2): `randn_tensor`, and `VaeImageProcessor.{get_default_height_width, resize,
preprocess, postprocess}` for tensor inputs.

Run here (needs /root/reference):  python tests/golden/make_pipeline_call_golden.py
"""
from __future__ import annotations

import sys
import types
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(Path(__file__).resolve().parent))
import make_pipeline_ref_golden as base  # noqa: E402
from oracle import flux_oracle as fo  # noqa: E402
from oracle import vae_oracle as vo  # noqa: E402

TOY_FLUX, TOY_VAE = base.TOY_FLUX, base.TOY_VAE


# ------------------------------------------------------------------------------------------ [dep-spec] stand-ins
def randn_tensor(shape, generator=None, device=None, dtype=None, layout=None):
    """[dep-spec] diffusers.utils.torch_utils.randn_tensor: CPU generators draw on the CPU; a list of generators draws
    one batch item each (a list of one is that one)."""
    rand_device = device
    batch_size = shape[0]
    device = torch.device(device or "cpu")
    if generator is not None:
        gtype = generator.device.type if not isinstance(generator, list) else generator[0].device.type
        if gtype != device.type and gtype == "cpu":
            rand_device = "cpu"
        elif gtype != device.type and gtype == "cuda":
            raise ValueError(f"Cannot generate a {device} tensor from a generator of type {gtype}.")
    if isinstance(generator, list) and len(generator) == 1:
        generator = generator[0]
    if isinstance(generator, list):
        shape = (1,) + tuple(shape[1:])
        latents = [torch.randn(shape, generator=generator[i], device=rand_device, dtype=dtype) for i in range(batch_size)]
        return torch.cat(latents, dim=0).to(device)
    return torch.randn(tuple(shape), generator=generator, device=rand_device, dtype=dtype).to(device)


class VaeImageProcessorSpec:
    """[dep-spec] diffusers.image_processor.VaeImageProcessor(vae_scale_factor=16) on torch tensors."""

    def __init__(self, vae_scale_factor=8, **kw):
        self.vae_scale_factor = vae_scale_factor

    def get_default_height_width(self, image, height=None, width=None):
        height = image.shape[2] if height is None else height
        width = image.shape[3] if width is None else width
        return tuple(x - x % self.vae_scale_factor for x in (height, width))

    @staticmethod
    def resize(image, height, width):
        return torch.nn.functional.interpolate(image, size=(height, width))

    def preprocess(self, image, height=None, width=None):
        if image.ndim == 3:
            image = image.unsqueeze(0)
        height, width = self.get_default_height_width(image, height, width)
        image = self.resize(image, height, width)
        if image.min() < 0:          # diffusers warns and leaves [-1, 1] inputs alone
            return image
        return 2.0 * image - 1.0

    @staticmethod
    def postprocess(image, output_type="pil"):
        assert output_type == "pt"
        return (image * 0.5 + 0.5).clamp(0, 1)


class OracleVaeFull(base.OracleVae):
    def decode(self, z, return_dict=True):
        img = vo.decode(self.sd, self.cfg, z)
        return (img,) if not return_dict else types.SimpleNamespace(sample=img)


def load_reference():
    ref = base.load_reference_pipeline_module()          # installs the diffusers stand-ins, executes the reference file
    ref.randn_tensor = randn_tensor                      # the names the reference file imported
    ref.VaeImageProcessor = VaeImageProcessorSpec
    return ref


# ------------------------------------------------------------------------------------------ the cases
def case_inputs(name):
    """Everything a case needs, from seeds only: (call kwargs without callbacks, flux config overrides)."""
    g = torch.Generator().manual_seed(sum(map(ord, name)))
    jd, pd = TOY_FLUX["joint_attention_dim"], TOY_FLUX["pooled_projection_dim"]

    def embeds(B, L=12):
        return dict(prompt_embeds=torch.randn(B, L, jd, generator=g), pooled_prompt_embeds=torch.randn(B, pd, generator=g))

    def img(B, H, W, lo=-1.0):
        return torch.rand(B, 3, H, W, generator=g) * (1.0 - lo) + lo

    common = dict(guidance_scale=3.5, output_type="latent", _auto_resize=False)
    flux_over = {}
    if name == "gen_noise":
        kw = dict(image=img(2, 64, 64), height=64, width=64, num_inference_steps=2, max_area=64 * 64, **embeds(2),
                  generator=("seed", 7))
    elif name == "gen_list":
        kw = dict(image=img(2, 64, 64), height=64, width=64, num_inference_steps=2, max_area=64 * 64, **embeds(2),
                  generator=("seeds", [1, 2]))
    elif name in ("callback", "interrupt"):
        kw = dict(image=img(1, 64, 64), height=64, width=64, num_inference_steps=3, max_area=64 * 64, **embeds(1),
                  latents=torch.randn(1, 16, 64, generator=g), guidance_scale=2.0)
    elif name == "no_guidance":
        flux_over = dict(guidance_embeds=False)
        kw = dict(image=img(1, 64, 64), height=64, width=64, num_inference_steps=2, max_area=64 * 64, **embeds(1),
                  latents=torch.randn(1, 16, 64, generator=g))
    elif name == "sigmas":
        kw = dict(image=img(1, 64, 64), height=64, width=64, num_inference_steps=3, sigmas=[1.0, 0.6, 0.3], max_area=64 * 64,
                  **embeds(1), latents=torch.randn(1, 16, 64, generator=g))
    elif name == "default_size":
        kw = dict(image=img(1, 64, 64), num_inference_steps=2, max_area=64 * 64, **embeds(1), generator=("seed", 3))
    elif name == "floor_size":
        kw = dict(image=img(1, 64, 96), height=70, width=100, num_inference_steps=2, max_area=70 * 100, **embeds(1),
                  generator=("seed", 4))
    elif name == "decode_pt":
        kw = dict(image=img(1, 64, 64), height=64, width=64, num_inference_steps=1, max_area=64 * 64, **embeds(1),
                  latents=torch.randn(1, 16, 64, generator=g), output_type="pt")
    elif name == "auto_resize":
        # a [0, 1] image off the preferred list: nearest-resized to (w, h) = (1248, 832), then normalised
        kw = dict(image=img(1, 32, 48, lo=0.0), height=64, width=64, num_inference_steps=1, max_area=64 * 64, **embeds(1),
                  latents=torch.randn(1, 16, 64, generator=g), _auto_resize=True)
    elif name == "text_to_image":
        kw = dict(image=None, height=64, width=96, num_inference_steps=2, max_area=64 * 96, **embeds(2), generator=("seed", 5))
    elif name == "latent_context":
        kw = dict(image=torch.randn(1, 16, 8, 12, generator=g), height=64, width=64, num_inference_steps=2, max_area=64 * 64,
                  **embeds(1), latents=torch.randn(1, 16, 64, generator=g))
    elif name == "shared_context":
        kw = dict(image=img(1, 64, 64), height=64, width=64, num_inference_steps=2, max_area=64 * 64, **embeds(2),
                  latents=torch.randn(2, 16, 64, generator=g))
    else:
        raise KeyError(name)
    return {**common, **kw}, flux_over


CASES = ["gen_noise", "gen_list", "callback", "interrupt", "no_guidance", "sigmas", "default_size", "floor_size", "decode_pt",
         "auto_resize", "text_to_image", "latent_context", "shared_context"]


def materialise(kw):
    kw = dict(kw)
    gen = kw.get("generator")
    if isinstance(gen, tuple):
        kw["generator"] = torch.Generator().manual_seed(gen[1]) if gen[0] == "seed" else \
            [torch.Generator().manual_seed(s) for s in gen[1]]
    for k, v in kw.items():
        if isinstance(v, torch.Tensor):
            kw[k] = v.clone()
    return kw


def make_callbacks(name, log):
    """Step callbacks of the `callback` / `interrupt` cases (the same functions drive the reference and the product)."""
    if name == "callback":
        def cb(pipe, i, t, kwargs):
            log.append(dict(i=i, t=float(t), keys=sorted(kwargs), num_timesteps=pipe.num_timesteps,
                            guidance_scale=pipe.guidance_scale, current=float(pipe.current_timestep),
                            interrupt=pipe.interrupt))
            out = {}
            if i == 0:
                out["latents"] = kwargs["latents"] * 0.5
                out["prompt_embeds"] = kwargs["prompt_embeds"] * 1.1
            return out
        return dict(callback_on_step_end=cb, callback_on_step_end_tensor_inputs=["latents", "prompt_embeds"])
    if name == "interrupt":
        def cb(pipe, i, t, kwargs):
            log.append(dict(i=i, keys=sorted(kwargs)))
            pipe._interrupt = True
            return {}
        return dict(callback_on_step_end=cb)
    return {}


def build_components(flux_over):
    fcfg, vcfg = fo.FluxConfig(**{**TOY_FLUX, **flux_over}), vo.VaeConfig(**TOY_VAE)
    fsd = fo.make_synthetic_state_dict(fcfg, seed=3, dtype=torch.float32)
    vsd = vo.make_synthetic_state_dict(vcfg, seed=4, dtype=torch.float32)
    return base.OracleTransformer(fsd, fcfg), OracleVaeFull(vsd, vcfg), base.OracleScheduler()


def run_case(pipeline_cls, name):
    kw, flux_over = case_inputs(name)
    tr, vae, sched = build_components(flux_over)
    tr.device = torch.device("cpu")
    pipe = pipeline_cls(scheduler=sched, vae=vae, text_encoder=None, tokenizer=None, text_encoder_2=None, tokenizer_2=None,
                        transformer=tr)
    log = []
    out = pipe(**materialise(kw), **make_callbacks(name, log)).images
    return dict(images=out, n_forwards=len(tr.calls), n_tokens=[c["n_tokens"] for c in tr.calls],
                timesteps=torch.stack([c["timestep"] for c in tr.calls]),
                guidance=None if tr.calls[0]["guidance"] is None else tr.calls[0]["guidance"],
                img_ids=tr.calls[0]["img_ids"], txt_ids=tr.calls[0]["txt_ids"], callback_log=log,
                num_timesteps=pipe.num_timesteps, current_timestep=pipe.current_timestep)


if __name__ == "__main__":
    if not base.REF_FILE.exists():
        raise SystemExit("needs /root/reference (run in the build container)")
    ref = load_reference()
    fx = {}
    for name in CASES:
        fx[name] = run_case(ref.FluxKontextPipeline, name)
        print(f"{name:16s} forwards {fx[name]['n_forwards']}  tokens {fx[name]['n_tokens'][0]}  "
              f"out {tuple(fx[name]['images'].shape)}  |out| {fx[name]['images'].abs().mean():.4f}")
    torch.save(fx, Path(__file__).with_name("pipeline_ref_call.pt"))
    print("wrote pipeline_ref_call.pt")
