"""Generates tests/golden/tower_ref.pt by EXECUTING THE REFERENCE'S OWN
/root/reference/univa/models/modeling_univa_denoise_tower.py (`UnivaDenoiseTower.__init__`, `_init_denoise_projector`,
`forward` — SURVEY.md §8 row a12 and the MLP2 half of a10) and configuration_univa_denoise_tower.py.

`diffusers` is not installed, so the names that file imports from it are stand-ins: `FluxTransformer2DModel.from_config`
returns a RECORDING denoiser (it stores every keyword it is called with and returns `(2 * hidden_states,)`), which is
exactly what is needed to pin the glue: which tensors are concatenated in which order, what `txt_ids` looks like,
which keyword arguments are dropped, forwarded or popped.  The projector (`nn.Sequential(Linear, SiLU, Linear)`) is the
reference's real torch module.  Both reference files are loaded BY PATH (this repo has its own `univa` package).

Run here (needs /root/reference):  python tests/golden/make_tower_ref_golden.py
"""
from __future__ import annotations

import importlib.util
import sys
import types
from pathlib import Path

import torch

REF = Path("/root/reference/univa/models")


class RecordingDenoiser(torch.nn.Module):
    def __init__(self):
        super().__init__()
        self.calls = []

    def forward(self, **kw):
        self.calls.append(kw)
        return (kw["hidden_states"] * 2,)


def load_reference_tower():
    def mod(name, **attrs):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        sys.modules[name] = m
        return m

    saved = {k: v for k, v in sys.modules.items() if k == "diffusers" or k.startswith(("diffusers.", "univa"))}
    for k in saved:
        del sys.modules[k]
    flux_cls = type("FluxTransformer2DModel", (), {"from_config": staticmethod(lambda cfg: RecordingDenoiser())})
    d = mod("diffusers", FluxTransformer2DModel=flux_cls, SD3Transformer2DModel=type("SD3Transformer2DModel", (), {}))
    d.__path__ = []
    mod("diffusers.utils", is_torch_version=lambda *a: True)
    mod("diffusers.models").__path__ = []
    mod("diffusers.models.modeling_outputs", Transformer2DModelOutput=type("Transformer2DModelOutput", (), {}))
    try:
        # the reference's own config class, registered under the module name its tower file imports
        mod("univa").__path__ = []
        mod("univa.models").__path__ = []
        spec = importlib.util.spec_from_file_location("univa.models.configuration_univa_denoise_tower",
                                                      REF / "configuration_univa_denoise_tower.py")
        cfg_mod = importlib.util.module_from_spec(spec)
        sys.modules[spec.name] = cfg_mod
        spec.loader.exec_module(cfg_mod)
        spec = importlib.util.spec_from_file_location("ref_tower", REF / "modeling_univa_denoise_tower.py")
        tower_mod = importlib.util.module_from_spec(spec)
        sys.modules["ref_tower"] = tower_mod      # transformers' PreTrainedModel.__init__ looks its class module up
        spec.loader.exec_module(tower_mod)
    finally:
        for k in [k for k in sys.modules if k == "diffusers" or k.startswith(("diffusers.", "univa"))]:
            del sys.modules[k]
        sys.modules.update(saved)
    return cfg_mod, tower_mod


def describe(call: dict) -> dict:
    return {k: (v.clone() if torch.is_tensor(v) else v) for k, v in call.items()}


def main():
    cfg_mod, tower_mod = load_reference_tower()
    cfg = cfg_mod.UnivaDenoiseTowerConfig(denoiser_type="flux", denoise_projector_type="mlp2x_gelu", input_hidden_size=64,
                                          output_hidden_size=32, denoiser_config={"num_layers": 1})
    torch.manual_seed(0)
    tower = tower_mod.UnivaDenoiseTower(cfg).eval()
    g = torch.Generator().manual_seed(1)
    hs = torch.randn(2, 10, 64, generator=g)
    vlm = torch.randn(2, 5, 32, generator=g)
    t5 = torch.randn(2, 3, 32, generator=g)
    pooled = torch.randn(2, 8, generator=g)
    t = torch.tensor([0.25, 0.75])
    img_ids = torch.zeros(10, 3)
    guidance = torch.full((2,), 3.5)
    cases = {}
    with torch.no_grad():
        out = tower(hs, t, vlm, pooled, prefix_prompt_embeds=t5, img_ids=img_ids, guidance=guidance,
                    joint_attention_kwargs={"attention_mask": torch.ones(2, 18)}, enc_attention_mask=torch.ones(2, 5))
        cases["vlm_plus_prefix"] = dict(out=out, call=describe(tower.denoiser.calls[-1]))
        out = tower(hs, t, vlm, pooled, img_ids=img_ids, guidance=guidance)
        cases["vlm_only"] = dict(out=out, call=describe(tower.denoiser.calls[-1]))
        out = tower(hs, t, None, pooled, prefix_prompt_embeds=t5, img_ids=img_ids, guidance=guidance)
        cases["prefix_only"] = dict(out=out, call=describe(tower.denoiser.calls[-1]))
        x = torch.randn(3, 7, 64, generator=g)
        proj_sd = {k: v.clone() for k, v in tower.denoise_projector.state_dict().items()}
        y32 = tower.denoise_projector(x)
        y16 = tower.denoise_projector.to(torch.bfloat16)(x.bfloat16())
    fx = dict(inputs=dict(hs=hs, vlm=vlm, t5=t5, pooled=pooled, t=t, img_ids=img_ids, guidance=guidance), cases=cases,
              projector=dict(state_dict=proj_sd, x=x, y_fp32=y32, y_bf16=y16,
                             structure=[type(m).__name__ for m in tower.denoise_projector],
                             keys=sorted(k for k in tower.state_dict() if k.startswith("denoise_projector"))))
    torch.save(fx, Path(__file__).with_name("tower_ref.pt"))
    for k, c in cases.items():
        print(k, "forwarded keywords:", sorted(c["call"]), "enc tokens:", c["call"]["encoder_hidden_states"].shape[1])
    print("projector:", fx["projector"]["structure"], fx["projector"]["keys"])


if __name__ == "__main__":
    if not REF.exists():
        raise SystemExit("needs /root/reference (run in the build container)")
    main()
