"""Generates tests/golden/host_ref.pt by importing (by path) the reference's own pure-Python host modules, which need
no third-party stand-ins:
  /root/reference/univa/utils/anyres_util.py                      pick_ratio / compute_size / dynamic_resize on a size grid
  /root/reference/univa/utils/denoiser_prompt_embedding_flux.py   encode_prompt & helpers with stub encoders / tokenizers
and by executing the SOURCE of single functions of files whose module-level imports are not satisfiable here
(extracted with `ast`, run in a namespace that holds only what the function uses):
  /root/reference/univa/serve/cli.py          update_size, prepare_condition_images (on PNG files written by this script)
  /root/reference/train_denoiser.py           get_trainable_params, check_param_is_in_components (the parameter-name
                                              contract of SURVEY.md §8b: which diffusers key names are un-frozen)
Run here (needs /root/reference):  python tests/golden/make_host_ref_golden.py
"""
from __future__ import annotations

import importlib.util
from pathlib import Path
from types import SimpleNamespace

import torch

REF = Path("/root/reference/univa/utils")


def load(name):
    spec = importlib.util.spec_from_file_location("ref_" + name, REF / f"{name}.py")
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


class StubTok:
    """tokenizer protocol the reference uses: call -> object with .input_ids [B, max_length]"""

    def __init__(self, base):
        self.base = base

    def __call__(self, prompt, padding=None, max_length=None, truncation=None, return_length=None,
                 return_overflowing_tokens=None, return_tensors=None):
        rows = [[self.base + (len(p) * 7 + j) % 50 for j in range(max_length)] for p in prompt]
        return SimpleNamespace(input_ids=torch.tensor(rows))


class StubT5:
    dtype, device = torch.float32, torch.device("cpu")

    def __call__(self, ids):
        return (ids.float()[..., None] * torch.arange(1, 5).float(),)


class StubClip(StubT5):
    def __call__(self, ids, output_hidden_states=False):
        return SimpleNamespace(pooler_output=ids.float()[:, :3] * 0.5)


def extract_functions(path: Path, names, namespace: dict) -> dict:
    import ast
    src = path.read_text()
    tree = ast.parse(src)
    for node in tree.body:
        if isinstance(node, ast.FunctionDef) and node.name in names:
            exec(compile(ast.Module(body=[node], type_ignores=[]), str(path), "exec"), namespace)
    return namespace


def cli_and_training_host_functions(any_, tmp: Path) -> dict:
    import numpy as np
    from PIL import Image

    ns = extract_functions(Path("/root/reference/univa/serve/cli.py"), {"update_size", "prepare_condition_images"},
                           dict(Image=Image, np=np, torch=torch, dynamic_resize=any_.dynamic_resize))
    rng = np.random.default_rng(0)
    files = []
    for i, (h, w) in enumerate([(30, 42), (64, 64), (72, 128)]):      # small files, same aspect classes as 300x420 / 720x1280
        f = tmp / f"im{i}.png"
        Image.fromarray(rng.integers(0, 256, size=(h, w, 3), dtype=np.uint8)).save(f)
        files.append(str(f))
    sizes = {}
    for anchor in (1024 * 1024, 256 * 256):
        sizes[("none", anchor)] = ns["update_size"](None, None, "any_11ratio", anchor)
        for i, f in enumerate(files):
            sizes[(i, anchor)] = ns["update_size"](f, None, "any_11ratio", anchor)
        sizes[("0+2", anchor)] = ns["update_size"](files[0], files[2], "any_11ratio", anchor)
    cond = ns["prepare_condition_images"](files[:1], "cpu")
    images = [np.asarray(Image.open(f).convert("RGB")) for f in files]
    from typing import List
    tns = extract_functions(Path("/root/reference/train_denoiser.py"), {"get_trainable_params", "check_param_is_in_components"},
                            dict(List=List))
    comps = dict(default=tns["get_trainable_params"](),
                 both_branches=tns["get_trainable_params"](only_img_branch=False),
                 some_layers=tns["get_trainable_params"](layers_to_train=[0, 18, 19, 56]))
    probe = ["denoise_tower.denoiser.transformer_blocks.3.attn.to_q.weight", "denoise_tower.denoiser.transformer_blocks.3.ff.net.0.proj.weight",
             "denoise_tower.denoiser.single_transformer_blocks.37.norm.linear.bias", "denoise_tower.denoiser.x_embedder.weight",
             "denoise_tower.denoiser.transformer_blocks.30.attn.to_q.weight", "denoise_tower.denoise_projector.0.weight"]
    checks = {mode: [tns["check_param_is_in_components"](n, c) for n in probe] for mode, c in comps.items()}
    return dict(images=images, update_size=sizes, condition=cond, components=comps, probe=probe, probe_result=checks)


def main():
    any_ = load("anyres_util")
    sizes = [(h, w) for h in (256, 300, 512, 720, 768, 1024, 1365, 2048) for w in (256, 400, 512, 1024, 1280, 1500, 2048)]
    anyres = {}
    for mode in ("any_11ratio", "any_17ratio"):
        for (h, w) in sizes:
            rw, rh = any_.pick_ratio(h, w, anyres=mode)
            anyres[(mode, h, w)] = dict(ratio=(rw, rh),
                                        size16=any_.compute_size(rw, rh, stride=16, anchor_pixels=1024 * 1024),
                                        size28=any_.compute_size(rw, rh, stride=28, min_pixels=448 * 448, max_pixels=448 * 448),
                                        dyn=any_.dynamic_resize(h, w, mode, anchor_pixels=1024 * 1024),
                                        dyn512=any_.dynamic_resize(h, w, mode, anchor_pixels=512 * 512))
    pe = load("denoiser_prompt_embedding_flux")
    toks, encs = [StubTok(100), StubTok(500)], [StubClip(), StubT5()]
    prompts = ["turn the sky red", "b"]
    out = {}
    e, p = pe.encode_prompt(encs, toks, prompts, 16, device="cpu", num_images_per_prompt=3)
    out["both_n3"] = dict(embeds=e, pooled=p)
    e, p = pe.encode_prompt(encs, toks, "single", 8, device="cpu", num_images_per_prompt=1)
    out["single"] = dict(embeds=e, pooled=p)
    e, p = pe.encode_prompt(encs, [None, toks[1]], "single", 8, device="cpu")
    out["no_clip_tokenizer"] = dict(embeds=e, pooled=p)
    e, p = pe.encode_prompt([encs[0], None], toks, "single", 8, device="cpu")
    out["no_t5_encoder"] = dict(embeds=e, pooled=p)
    out["tokenize_prompt"] = pe.tokenize_prompt(toks[1], ["x y"], 6)
    try:
        pe._encode_prompt_with_t5(encs[1], None, 8, "p")
        out["error_no_ids"] = None
    except ValueError as ex:
        out["error_no_ids"] = str(ex)
    import tempfile
    with tempfile.TemporaryDirectory() as td:
        host = cli_and_training_host_functions(any_, Path(td))
    print("update_size:", host["update_size"], "| trainable components:", {k: len(v) for k, v in host["components"].items()})
    torch.save(dict(anyres=anyres, encode_prompt=out, host=host), Path(__file__).with_name("host_ref.pt"))
    print(len(anyres), "anyres entries;", {k: (None if v is None else "ok") for k, v in out.items() if k.startswith("error")})


if __name__ == "__main__":
    if not REF.exists():
        raise SystemExit("needs /root/reference (run in the build container)")
    main()
