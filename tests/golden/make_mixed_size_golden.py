"""Generates tests/golden/mixed_size_ref.pt by executing the reference's OWN statements for mixed-size training batches:
`pad_x_and_mask` (train_denoiser.py:158-183) and the loss section of the training loop (:1104-1165: target, loss weighting,
area-mask weights as a list / tensor, weight_mask, the normalisation), extracted with `ast` and run on seeded stand-ins for
the tensors the loop holds at that point.  diffusers.training_utils.compute_loss_weighting_for_sd3 (third party, absent here)
is restated from its published definition.  Run here (needs /root/reference):  python tests/golden/make_mixed_size_golden.py"""
import ast
import math
from pathlib import Path
from types import SimpleNamespace

import torch
import torch.nn.functional as F

REF = Path("/root/reference/train_denoiser.py")


def compute_loss_weighting_for_sd3(weighting_scheme, sigmas=None):
    if weighting_scheme == "sigma_sqrt":
        return (sigmas ** -2.0).float()
    if weighting_scheme == "cosmap":
        return 2 / (math.pi * (1 - 2 * sigmas + 2 * sigmas ** 2))
    return torch.ones_like(sigmas)


def loop_statements(tree, lo, hi):
    """the statements of the training loop body whose first line lies in [lo, hi], in order (one nesting level)"""
    best = []
    for node in ast.walk(tree):
        for field in ("body", "orelse"):
            body = getattr(node, field, None)
            if not isinstance(body, list):
                continue
            sel = [st for st in body if isinstance(st, ast.stmt) and lo <= st.lineno <= hi]
            if len(sel) > len(best):
                best = sel
    return [st for st in best if not isinstance(st, ast.FunctionDef)]


def main():
    tree = ast.parse(REF.read_text())
    pad_fn = next(n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name == "pad_x_and_mask")
    stmts = loop_statements(tree, 1104, 1165)
    assert len(stmts) >= 6, [s.lineno for s in stmts]
    code = compile(ast.Module(body=[pad_fn] + stmts, type_ignores=[]), str(REF), "exec")
    g = torch.Generator().manual_seed(1234)
    cases = []
    for name, sizes, scheme, sigw, mwt, area in [
        ("mixed", [(12, 20), (16, 16), (8, 24)], "logit_normal", False, None, None),
        ("mixed_area_list", [(12, 20), (16, 16)], "cosmap", False, "log", "list"),
        ("mixed_sigmas", [(8, 8), (16, 8)], "logit_normal", True, None, None),
        ("same_size_area", [(8, 8), (8, 8)], "sigma_sqrt", False, "log", "tensor"),
        ("same_size_plain", [(8, 12), (8, 12)], "logit_normal", False, None, None),
    ]:
        B, C = len(sizes), 16
        mixed = len(set(sizes)) > 1
        unpad = [torch.randn(1, C, h, w, generator=g) for h, w in sizes]
        ns = {"torch": torch, "F": F, "compute_loss_weighting_for_sd3": compute_loss_weighting_for_sd3}
        exec(compile(ast.Module(body=[pad_fn], type_ignores=[]), str(REF), "exec"), ns)
        if mixed:
            model_input, mask = ns["pad_x_and_mask"](unpad, [torch.ones_like(x) for x in unpad])
            weight_mask = mask.detach().clone()
        else:
            model_input, mask, weight_mask = torch.cat(unpad, 0), None, None
        H, W = model_input.shape[-2:]
        noise = torch.randn(model_input.shape, generator=g)
        model_pred = torch.randn(model_input.shape, generator=g)
        sigmas = torch.rand(B, 1, 1, 1, generator=g) * 0.8 + 0.1
        if area == "list":      # per-sample weights at image resolution (4x the latent here), resized by the loop itself
            area_w = [torch.rand(1, 1, 4 * h, 4 * w, generator=g) + 0.5 for h, w in sizes]
        elif area == "tensor":
            area_w = torch.rand(B, 1, 4 * H, 4 * W, generator=g) + 0.5
        else:
            area_w = None
        ns.update(dict(
            args=SimpleNamespace(training_config=SimpleNamespace(weighting_scheme=scheme, sigmas_as_weight=sigw,
                                                                 mask_weight_type=mwt),
                                 dataset_config=SimpleNamespace(batch_size=B)),
            accelerator=SimpleNamespace(device=torch.device("cpu")), noise=noise, model_input=model_input,
            model_pred=model_pred.clone(), sigmas=sigmas, area_mask_weights=area_w,
            unpad_model_input=unpad if mixed else None, weight_mask=weight_mask))
        exec(code, ns)
        cases.append(dict(name=name, sizes=sizes, scheme=scheme, sigmas_as_weight=sigw, mask_weight_type=mwt,
                          unpad=unpad, model_input=model_input, mask=mask, noise=noise, model_pred=model_pred, sigmas=sigmas,
                          area_weights=area_w, target=ns["target"], weighting=ns["weighting"].float().expand(B, 1, H, W).clone()
                          if ns["weighting"].shape[-1] == 1 else ns["weighting"].float(), loss=ns["loss"].clone()))
        print(name, "loss", float(ns["loss"]))
    torch.save(dict(cases=cases, lines=[s.lineno for s in stmts]), Path(__file__).with_name("mixed_size_ref.pt"))


if __name__ == "__main__":
    main()
