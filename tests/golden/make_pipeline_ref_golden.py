"""Generates tests/golden/pipeline_ref_loop.pt by EXECUTING THE REFERENCE'S OWN PIPELINE FILE
(/root/reference/univa/utils/flux_pipeline.py: `FluxKontextPipeline.__call__`, `prepare_latents`,
`_pack_latents`, `_unpack_latents`, `_prepare_latent_image_ids`, `_encode_vae_image`, `calculate_shift`,
`retrieve_timesteps`, `encode_prompt`, `check_inputs` — the code SURVEY.md §8 rows a1 / a8 name) on a toy model.

The file imports `diffusers`, which is not installed here, so the THIRD-PARTY names it imports are replaced by
minimal stand-ins before the file is loaded (nothing in the reference file itself is altered):
  * `DiffusionPipeline`            register_modules / _execution_device / progress_bar / maybe_free_model_hooks
  * `VaeImageProcessor`            identity on tensors that are already [-1,1] at the target size (what diffusers
                                   does for such tensors); `postprocess` is never reached (output_type="latent")
  * `FlowMatchEulerDiscreteScheduler`, `AutoencoderKL`, `FluxTransformer2DModel`  type names only: the objects
                                   handed to the pipeline are thin adapters over oracle/ (scheduler restated from
                                   SURVEY.md A.5, transformer = oracle.flux_oracle.flux_forward, VAE = oracle.vae_oracle)
  * loader mixins, logging, lora helpers, randn_tensor, FluxPipelineOutput      trivial
So the fixture pins the oracle's LOOP (size rule, latent/ids layout, [target ‖ context] concat and slice,
guidance vector, timestep / 1000, sigma schedule inputs, scheduler call protocol, VAE affine) against the
reference's own statements; the transformer / VAE / scheduler arithmetic inside the adapters is pinned elsewhere
(torchtitan cross-check, tests/golden/*_titan.pt).

Run here (needs /root/reference):  python tests/golden/make_pipeline_ref_golden.py
"""
from __future__ import annotations

import contextlib
import importlib.util
import sys
import types
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
from oracle import flux_oracle as fo  # noqa: E402
from oracle import pipeline_oracle as po  # noqa: E402
from oracle import vae_oracle as vo  # noqa: E402

REF_FILE = Path("/root/reference/univa/utils/flux_pipeline.py")
TOY_FLUX = dict(num_layers=1, num_single_layers=2, attention_head_dim=128, num_attention_heads=2, joint_attention_dim=64,
                pooled_projection_dim=32)
TOY_VAE = dict(block_out_channels=(32, 32, 64, 64), layers_per_block=1)


# ------------------------------------------------------------------------------------------ third-party stand-ins
def install_diffusers_standins():
    def mod(name):
        m = types.ModuleType(name)
        sys.modules[name] = m
        return m

    root = mod("diffusers")
    root.__path__ = []
    ip = mod("diffusers.image_processor")
    ip.PipelineImageInput = object

    class VaeImageProcessor:
        def __init__(self, vae_scale_factor=8, **kw):
            self.vae_scale_factor = vae_scale_factor

        @staticmethod
        def get_default_height_width(image, height=None, width=None):
            return image.shape[-2], image.shape[-1]

        @staticmethod
        def resize(image, height, width):
            assert tuple(image.shape[-2:]) == (height, width), "stand-in: inputs are already at the target size"
            return image

        @staticmethod
        def preprocess(image, height=None, width=None):
            assert float(image.min()) < 0, "stand-in: tensors already normalised to [-1, 1] pass through"
            return image

        def postprocess(self, *a, **k):
            raise AssertionError("postprocess is not part of this fixture (output_type='latent')")

    ip.VaeImageProcessor = VaeImageProcessor
    ld = mod("diffusers.loaders")
    for n in ("FluxIPAdapterMixin", "FluxLoraLoaderMixin", "FromSingleFileMixin", "TextualInversionLoaderMixin"):
        setattr(ld, n, type(n, (), {}))
    mod("diffusers.models").__path__ = []
    mod("diffusers.models.autoencoders").AutoencoderKL = type("AutoencoderKL", (), {})
    mod("diffusers.models.transformers").FluxTransformer2DModel = type("FluxTransformer2DModel", (), {})
    mod("diffusers.schedulers").FlowMatchEulerDiscreteScheduler = type("FlowMatchEulerDiscreteScheduler", (), {})
    ut = mod("diffusers.utils")
    ut.__path__ = []
    ut.USE_PEFT_BACKEND = False
    ut.is_torch_xla_available = lambda: False
    ut.replace_example_docstring = lambda doc: (lambda f: f)
    ut.scale_lora_layers = ut.unscale_lora_layers = lambda *a, **k: None
    import logging as pylog
    ut.logging = types.SimpleNamespace(get_logger=lambda name: pylog.getLogger(name))

    def randn_tensor(shape, generator=None, device=None, dtype=None, layout=None):
        return torch.randn(shape, generator=generator, dtype=dtype).to(device)

    mod("diffusers.utils.torch_utils").randn_tensor = randn_tensor
    mod("diffusers.pipelines").__path__ = []

    class DiffusionPipeline:
        def register_modules(self, **kw):
            for k, v in kw.items():
                setattr(self, k, v)

        @property
        def _execution_device(self):
            return torch.device("cpu")

        @contextlib.contextmanager
        def progress_bar(self, total=None):
            yield types.SimpleNamespace(update=lambda *a: None)

        def maybe_free_model_hooks(self):
            pass

    mod("diffusers.pipelines.pipeline_utils").DiffusionPipeline = DiffusionPipeline
    mod("diffusers.pipelines.flux").__path__ = []

    class FluxPipelineOutput:
        def __init__(self, images):
            self.images = images

    mod("diffusers.pipelines.flux.pipeline_output").FluxPipelineOutput = FluxPipelineOutput


def load_reference_pipeline_module():
    install_diffusers_standins()
    spec = importlib.util.spec_from_file_location("ref_flux_pipeline", REF_FILE)   # by path: `univa` here is this repo's shim
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


# ------------------------------------------------------------------------------------------ adapters over oracle/
class _Cfg(dict):
    __getattr__ = dict.__getitem__


class OracleTransformer:
    def __init__(self, sd, cfg):
        self.sd, self.cfg = sd, cfg
        self.config = _Cfg(in_channels=cfg.in_channels, guidance_embeds=cfg.guidance_embeds)
        self.dtype = next(iter(sd.values())).dtype
        self.calls = []

    def __call__(self, hidden_states, timestep, guidance, pooled_projections, encoder_hidden_states, txt_ids, img_ids,
                 joint_attention_kwargs=None, return_dict=False):
        self.calls.append(dict(timestep=timestep.clone(), guidance=None if guidance is None else guidance.clone(),
                               n_tokens=hidden_states.shape[1], img_ids=img_ids.clone(), txt_ids=txt_ids.clone()))
        return (fo.flux_forward(self.sd, self.cfg, hidden_states, encoder_hidden_states, pooled_projections, timestep,
                                img_ids, txt_ids, guidance=guidance),)


class OracleVae:
    def __init__(self, sd, cfg):
        self.sd, self.cfg = sd, cfg
        self.config = _Cfg(block_out_channels=cfg.block_out_channels, latent_channels=cfg.latent_channels,
                           scaling_factor=cfg.scaling_factor, shift_factor=cfg.shift_factor)
        self.dtype = next(iter(sd.values())).dtype

    def encode(self, x):
        mode = vo.encode_mode(self.sd, self.cfg, x)
        return types.SimpleNamespace(latent_dist=types.SimpleNamespace(mode=lambda: mode, sample=lambda g=None: mode))


class OracleScheduler(po.EulerSchedulerOracle):
    """call protocol of diffusers' scheduler as the reference uses it (retrieve_timesteps / step(..., return_dict=False))"""

    def set_timesteps(self, num_inference_steps=None, device=None, sigmas=None, mu=None, **kw):
        super().set_timesteps(sigmas, mu, device=device)

    def step(self, model_output, timestep, sample, return_dict=False):
        return (super().step(model_output, timestep, sample),)

    @property
    def config(self):
        return self._config

    @config.setter
    def config(self, v):
        self._config = v


def run(dtype=torch.float32, H=64, W=96, steps=3, B=1, seed=5, true_cfg_scale=1.0):
    ref = load_reference_pipeline_module()
    fcfg, vcfg = fo.FluxConfig(**TOY_FLUX), vo.VaeConfig(**TOY_VAE)
    fsd = fo.make_synthetic_state_dict(fcfg, seed=3, dtype=dtype)
    vsd = vo.make_synthetic_state_dict(vcfg, seed=4, dtype=dtype)
    g = torch.Generator().manual_seed(seed)
    image = (torch.rand(B, 3, H, W, generator=g) * 2 - 1).to(dtype)
    pe = torch.randn(B, 12, TOY_FLUX["joint_attention_dim"], generator=g).to(dtype)
    pooled = torch.randn(B, TOY_FLUX["pooled_projection_dim"], generator=g).to(dtype)
    noise = torch.randn(B, (H // 16) * (W // 16), 64, generator=g).to(dtype)
    cfg_kw, cfg_kw_oracle = {}, {}
    if true_cfg_scale > 1:      # true classifier-free guidance: negative prompt of a different length
        npe = torch.randn(B, 7, TOY_FLUX["joint_attention_dim"], generator=g).to(dtype)
        npool = torch.randn(B, TOY_FLUX["pooled_projection_dim"], generator=g).to(dtype)
        cfg_kw = dict(true_cfg_scale=true_cfg_scale, negative_prompt_embeds=npe, negative_pooled_prompt_embeds=npool)
        cfg_kw_oracle = dict(true_cfg_scale=true_cfg_scale, negative_prompt_embeds=npe, negative_pooled=npool)
    tr = OracleTransformer(fsd, fcfg)
    pipe = ref.FluxKontextPipeline(scheduler=OracleScheduler(), vae=OracleVae(vsd, vcfg), text_encoder=None, tokenizer=None,
                                   text_encoder_2=None, tokenizer_2=None, transformer=tr)
    out = pipe(image=image, prompt_embeds=pe, pooled_prompt_embeds=pooled, height=H, width=W, num_inference_steps=steps,
               guidance_scale=3.5, latents=noise.clone(), output_type="latent", max_area=H * W, _auto_resize=False, **cfg_kw)
    latents = out.images
    mine = po.sample(fsd, fcfg, vsd, vcfg, image, pe, pooled, height=H, width=W, num_inference_steps=steps, guidance_scale=3.5,
                     latents=noise.clone(), output="latent", max_area=H * W, **cfg_kw_oracle)
    return dict(args=dict(H=H, W=W, steps=steps, B=B, seed=seed, dtype=str(dtype), true_cfg_scale=true_cfg_scale),
                n_forwards=len(tr.calls), latents=latents, oracle_latents=mine,
                timesteps=torch.stack([c["timestep"] for c in tr.calls]), guidance=tr.calls[0]["guidance"],
                n_tokens=tr.calls[0]["n_tokens"], img_ids=tr.calls[0]["img_ids"], txt_ids=tr.calls[0]["txt_ids"],
                helpers=dict(
                    calculate_shift=[float(ref.calculate_shift(n)) for n in (256, 1024, 4096)],
                    pack=ref.FluxKontextPipeline._pack_latents(torch.arange(2 * 16 * 4 * 6.0).view(2, 16, 4, 6), 2, 16, 4, 6),
                    ids=ref.FluxKontextPipeline._prepare_latent_image_ids(1, 3, 2, "cpu", torch.float32),
                    preferred=list(ref.PREFERRED_KONTEXT_RESOLUTIONS)))


if __name__ == "__main__":
    if not REF_FILE.exists():
        raise SystemExit("needs /root/reference (run in the build container)")
    fx = {"case_64x96": run(), "case_batch2_128x64": run(H=128, W=64, steps=2, B=2, seed=9),
          "case_true_cfg": run(H=64, W=64, steps=2, seed=11, true_cfg_scale=2.5)}
    for k, v in fx.items():
        d = (v["latents"] - v["oracle_latents"]).abs().max().item()
        print(k, "max |reference loop - oracle loop| =", d, " tokens per forward:", v["n_tokens"])
        assert d < 1e-5
    torch.save(fx, Path(__file__).with_name("pipeline_ref_loop.pt"))
    print("wrote pipeline_ref_loop.pt")
