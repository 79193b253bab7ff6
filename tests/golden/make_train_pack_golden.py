"""Generates tests/golden/train_pack_ref.pt by executing the reference's OWN statements that build the transformer's token
inputs in the training loop (train_denoiser.py:998-1056: `prepare_latents` on the context image with the noised target passed
as `latents`, manual packing of the target, `[target ‖ context]` concatenation of tokens and position ids; the no-context
fallback) with `flux_pipeline` = the reference's own `FluxKontextPipeline` loaded from its file (third-party diffusers names
replaced by the stand-ins of make_pipeline_ref_golden.py) over a stub VAE whose `encode` is a fixed arithmetic function
(StubVae below, imported by the test as well).  Run here (needs /root/reference):  python tests/golden/make_train_pack_golden.py"""
import ast
import sys
import types
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).parent))
from make_pipeline_ref_golden import OracleScheduler, load_reference_pipeline_module  # noqa: E402

REF = Path("/root/reference/train_denoiser.py")


class StubVae:
    """8x spatial reduction to 16 channels by fixed arithmetic: enough to tell every latent element apart"""
    dtype = torch.float32

    def __init__(self):
        self.config = types.SimpleNamespace(block_out_channels=(1, 1, 1, 1), latent_channels=16, scaling_factor=0.3611,
                                            shift_factor=0.1159)

    def encode(self, x):
        x = x.float()
        z = torch.nn.functional.avg_pool2d(x, 8)                                  # [B, 3, h, w]
        ch = torch.arange(16, dtype=torch.float32).view(1, 16, 1, 1)
        z = z[:, :1] * (1 + 0.1 * ch) + z[:, 1:2] * 0.01 * ch + z[:, 2:3]
        dist = types.SimpleNamespace(mode=lambda: z, sample=lambda generator=None: z)
        return types.SimpleNamespace(latent_dist=dist)


def statements(tree, lo, hi):
    for node in ast.walk(tree):
        for field in ("body", "orelse"):
            body = getattr(node, field, None)
            if isinstance(body, list) and any(isinstance(st, ast.stmt) and st.lineno == lo for st in body):
                return [st for st in body if lo <= st.lineno <= hi]
    raise SystemExit(f"no statement starts at line {lo}")


def main():
    ref = load_reference_pipeline_module()
    pipe = ref.FluxKontextPipeline(scheduler=OracleScheduler(), vae=StubVae(), text_encoder=None, tokenizer=None,
                                   text_encoder_2=None, tokenizer_2=None, transformer=types.SimpleNamespace(
                                       config=types.SimpleNamespace(in_channels=64), dtype=torch.float32))
    tree = ast.parse(REF.read_text())
    stmts = statements(tree, 998, 1056)
    code = compile(ast.Module(body=stmts, type_ignores=[]), str(REF), "exec")
    g = torch.Generator().manual_seed(77)
    cases = []
    for name, B, hw, cond_hw in [("context_same_size", 2, (8, 12), (64, 96)), ("context_other_size", 1, (8, 8), (96, 64)),
                                 ("no_context", 2, (6, 10), None)]:
        h, w = hw
        model_input = torch.randn(B, 16, h, w, generator=g)
        noisy = torch.randn(B, 16, h, w, generator=g)
        cond = None if cond_hw is None else torch.rand(B, 3, *cond_hw, generator=g) * 2 - 1
        ns = dict(torch=torch, flux_pipeline=pipe, FluxKontextPipeline=ref.FluxKontextPipeline, condition_pixel_values=cond,
                  model_input=model_input, noisy_model_input=noisy, vae_scale_factor=8, weight_dtype=torch.float32,
                  accelerator=types.SimpleNamespace(device=torch.device("cpu")))
        exec(code, ns)
        cases.append(dict(name=name, noisy=noisy, cond=cond, tokens=ns["packed_noisy_model_input"].clone(),
                          ids=ns["latent_image_ids"].clone()))
        print(name, tuple(ns["packed_noisy_model_input"].shape), tuple(ns["latent_image_ids"].shape))
    torch.save(dict(cases=cases, lines=[s.lineno for s in stmts]), Path(__file__).with_name("train_pack_ref.pt"))


if __name__ == "__main__":
    main()
