"""Generates tests/golden/check_inputs_ref.pt by executing the SOURCE of the reference's own
`FluxKontextPipeline.check_inputs` (univa/utils/flux_pipeline.py:490-560; extracted with `ast`, run with a stand-in
`self`) on a grid of argument combinations: which combinations are rejected, and with what message.
Run here (needs /root/reference):  python tests/golden/make_check_inputs_golden.py"""
import ast
import logging
from pathlib import Path
from types import SimpleNamespace

import torch

REF = Path("/root/reference/univa/utils/flux_pipeline.py")


def cases():
    E, P = "EMB", "POOLED"          # placeholders: the function only tests `is None`
    base = dict(prompt=None, prompt_2=None, height=1024, width=1024, negative_prompt=None, negative_prompt_2=None,
                prompt_embeds=None, negative_prompt_embeds=None, pooled_prompt_embeds=None, negative_pooled_prompt_embeds=None,
                callback_on_step_end_tensor_inputs=["latents"], max_sequence_length=512)
    deltas = [dict(prompt="a cat"), dict(prompt=["a", "b"]), dict(prompt_embeds=E, pooled_prompt_embeds=P), dict(),
              dict(prompt="a", prompt_embeds=E, pooled_prompt_embeds=P), dict(prompt_2="b", prompt_embeds=E, pooled_prompt_embeds=P),
              dict(prompt=3), dict(prompt="a", prompt_2=4.5), dict(prompt_embeds=E),
              dict(prompt="a", negative_prompt="n", negative_prompt_embeds=E, negative_pooled_prompt_embeds=P),
              dict(prompt="a", negative_prompt_2="n", negative_prompt_embeds=E, negative_pooled_prompt_embeds=P),
              dict(prompt="a", negative_prompt_embeds=E), dict(prompt="a", max_sequence_length=513),
              dict(prompt="a", callback_on_step_end_tensor_inputs=["latents", "bogus"]),
              dict(prompt="a", callback_on_step_end_tensor_inputs=["prompt_embeds"]), dict(prompt="a", height=1000, width=1000),
              dict(prompt="a", negative_prompt="n"), dict(prompt_embeds=E, pooled_prompt_embeds=P, negative_prompt_embeds=E,
                                                          negative_pooled_prompt_embeds=P)]
    return [dict(base, **d) for d in deltas]


def main():
    tree = ast.parse(REF.read_text())
    cls = next(n for n in tree.body if isinstance(n, ast.ClassDef) and n.name == "FluxKontextPipeline")
    fn = next(n for n in cls.body if isinstance(n, ast.FunctionDef) and n.name == "check_inputs")
    ns = {"logger": logging.getLogger("ref")}
    exec(compile(ast.Module(body=[fn], type_ignores=[]), str(REF), "exec"), ns)
    self = SimpleNamespace(vae_scale_factor=8, _callback_tensor_inputs=["latents", "prompt_embeds"])
    out = []
    for kw in cases():
        try:
            ns["check_inputs"](self, **kw)
            out.append(dict(kwargs=kw, raised=False, message=""))
        except ValueError as e:
            out.append(dict(kwargs=kw, raised=True, message=str(e)))
    torch.save(out, Path(__file__).with_name("check_inputs_ref.pt"))
    print(sum(o["raised"] for o in out), "of", len(out), "combinations rejected")


if __name__ == "__main__":
    main()
