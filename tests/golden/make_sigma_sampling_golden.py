"""Generates tests/golden/sigma_sampling_ref.pt by executing the reference's OWN statements for the flow-matching noising
step of the training loop (train_denoiser.py:935-995: noise, timestep / sigma sampling in both the discrete and the
continuous (resolution-shifted logit-normal) branch, `noisy_model_input`) and its `get_sigmas` helper (:779-788), extracted
with `ast` and run on the CPU with a seeded global generator.  Third-party pieces that are absent here are restated from their
published definitions and marked: diffusers' `compute_density_for_timestep_sampling` and the `timesteps` / `sigmas` tables a
`FlowMatchEulerDiscreteScheduler` holds after construction.  Run here (needs /root/reference):
    python tests/golden/make_sigma_sampling_golden.py"""
import ast
import math
from pathlib import Path
from types import SimpleNamespace

import torch

REF = Path("/root/reference/train_denoiser.py")


def compute_density_for_timestep_sampling(weighting_scheme, batch_size, logit_mean=None, logit_std=None, mode_scale=None):
    """diffusers.training_utils (restated)"""
    if weighting_scheme == "logit_normal":
        u = torch.normal(mean=logit_mean, std=logit_std, size=(batch_size,), device="cpu")
        u = torch.nn.functional.sigmoid(u)
    elif weighting_scheme == "mode":
        u = torch.rand(size=(batch_size,), device="cpu")
        u = 1 - u - mode_scale * (torch.cos(math.pi * u / 2) ** 2 - 1 + u)
    else:
        u = torch.rand(size=(batch_size,), device="cpu")
    return u


def scheduler_tables(cfg):
    """what FlowMatchEulerDiscreteScheduler.__init__ leaves in .timesteps / .sigmas (diffusers 0.32, restated)"""
    n = cfg["num_train_timesteps"]
    timesteps = torch.linspace(1, n, n).flip(0)
    sigmas = timesteps / n
    if not cfg["use_dynamic_shifting"]:
        sigmas = cfg["shift"] * sigmas / (1 + (cfg["shift"] - 1) * sigmas)
    return sigmas * n, sigmas


def statements(tree, lo, hi):
    """the statements of the block that holds line `lo`, up to line `hi`"""
    for node in ast.walk(tree):
        for field in ("body", "orelse"):
            body = getattr(node, field, None)
            if isinstance(body, list) and any(isinstance(st, ast.stmt) and st.lineno == lo for st in body):
                return [st for st in body if lo <= st.lineno <= hi]
    raise SystemExit(f"no statement starts at line {lo}")


def main():
    tree = ast.parse(REF.read_text())
    get_sigmas = next(n for n in ast.walk(tree) if isinstance(n, ast.FunctionDef) and n.name == "get_sigmas")
    stmts = statements(tree, 935, 995)
    assert len(stmts) >= 4, [s.lineno for s in stmts]
    code = compile(ast.Module(body=[get_sigmas] + stmts, type_ignores=[]), str(REF), "exec")
    flux_cfg = dict(num_train_timesteps=1000, shift=3.0, use_dynamic_shifting=True, base_shift=0.5, max_shift=1.15,
                    base_image_seq_len=256, max_image_seq_len=4096)
    cases = []
    for name, discrete, scheme, dyn, shape, seed in [
        ("continuous_64x64", False, "logit_normal", True, (2, 2, 64, 64), 11),
        ("continuous_32x48", False, "logit_normal", True, (3, 2, 32, 48), 12),
        ("discrete_logit_normal", True, "logit_normal", True, (4, 2, 8, 8), 13),
        ("discrete_mode_static_shift", True, "mode", False, (4, 2, 8, 8), 14),
        ("discrete_uniform", True, "null", True, (5, 2, 8, 8), 15),
    ]:
        cfg = dict(flux_cfg, use_dynamic_shifting=dyn)
        ts, sg = scheduler_tables(cfg)
        x = torch.randn(shape, generator=torch.Generator().manual_seed(seed + 100))
        ns = dict(torch=torch, math=math, compute_density_for_timestep_sampling=compute_density_for_timestep_sampling,
                  accelerator=SimpleNamespace(device=torch.device("cpu")),
                  noise_scheduler_copy=SimpleNamespace(config=SimpleNamespace(**cfg), timesteps=ts, sigmas=sg),
                  args=SimpleNamespace(training_config=SimpleNamespace(discrete_timestep=discrete, weighting_scheme=scheme,
                                                                       logit_mean=0.0, logit_std=1.0, mode_scale=1.29)),
                  model_input=x)
        torch.manual_seed(seed)
        exec(code, ns)
        cases.append(dict(name=name, seed=seed, discrete=discrete, scheme=scheme, sched=cfg, model_input=x, noise=ns["noise"],
                          sigmas=ns["sigmas"].clone(), timesteps=ns["timesteps"].clone(), noisy=ns["noisy_model_input"].clone()))
        print(name, ns["sigmas"].flatten().tolist())
    torch.save(dict(cases=cases, lines=[s.lineno for s in stmts]), Path(__file__).with_name("sigma_sampling_ref.pt"))


if __name__ == "__main__":
    main()
