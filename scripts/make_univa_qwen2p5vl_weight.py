#!/usr/bin/env python
"""Assemble a Univa checkpoint directory from a Qwen2.5-VL checkpoint and a FLUX checkpoint — the job of the reference's
scripts/make_univa_qwen2p5vl_weight.py:35-76 — without instantiating either model: tensors are streamed shard by shard
from safetensors to safetensors on the CPU (the reference builds both models in fp32 in host memory, ~150 GB).

    python scripts/make_univa_qwen2p5vl_weight.py --origin_qwenvl_ckpt_path Qwen2.5-VL-7B-Instruct \
        --origin_flux_ckpt_path FLUX.1-Kontext-dev --save_path UniWorld [--dtype bfloat16]

What the reference's script produces, and this one too:
  * every tensor of the Qwen2.5-VL checkpoint under its own name (`visual.*`, `model.*`, `lm_head.weight`:
    `model.load_state_dict(qwenvl.state_dict(), strict=False)` with only `denoise_tower.*` allowed to be missing, :57-59);
  * the FLUX transformer's tensors under `denoise_tower.denoiser.` (`model.denoise_tower.denoiser = flux`, :71);
  * a freshly initialised MLP2 `denoise_tower.denoise_projector.{0,2}.{weight,bias}`: Linear(hidden, 3*4096) . SiLU .
    Linear(3*4096, 4096) (modeling_univa_denoise_tower.py:31-47; trained in stage 1).  Initialisation: N(0, initializer_range)
    weights and zero biases, seeded (the reference relies on `_from_config`'s module initialisation);
  * config.json = the Qwen2.5-VL config with `model_type: univa_qwen2p5vl`, the Univa architecture name and a
    `denoise_tower` section {denoiser_type: flux, denoise_projector_type: mlp2x_gelu, input_hidden_size: hidden_size,
    output_hidden_size: 4096, denoiser_config: <FLUX transformer/config.json>} (:39-46);
  * the processor / tokenizer files next to it (`processor.save_pretrained(save_path)`, :76).
The directory is what gpt_image_edit_b200.checkpoint.load_univa_checkpoint and the reference's `from_pretrained` read.
The reference saves fp32; `--dtype` chooses (default: keep each tensor's dtype).
"""
from __future__ import annotations

import argparse
import json
import shutil
from pathlib import Path

import sys

import torch
from safetensors import safe_open

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from gpt_image_edit_b200.checkpoint import ShardWriter  # noqa: E402

PROCESSOR_FILES = ("tokenizer.json", "tokenizer_config.json", "vocab.json", "merges.txt", "added_tokens.json",
                   "special_tokens_map.json", "preprocessor_config.json", "processor_config.json", "chat_template.json",
                   "chat_template.jinja", "generation_config.json")
DTYPES = {"float32": torch.float32, "bfloat16": torch.bfloat16, "float16": torch.float16}


def iter_tensors(directory: Path):
    files = sorted(Path(directory).glob("*.safetensors"))
    if not files:
        raise FileNotFoundError(f"no *.safetensors under {directory}")
    for f in files:
        with safe_open(str(f), framework="pt", device="cpu") as sf:
            for k in sf.keys():
                yield k, sf.get_tensor(k)


def univa_config(qwen_cfg: dict, flux_cfg: dict) -> dict:
    cfg = dict(qwen_cfg)
    cfg["model_type"] = "univa_qwen2p5vl"
    cfg["architectures"] = ["UnivaQwen2p5VLForConditionalGeneration"]
    hidden = cfg.get("hidden_size", (cfg.get("text_config") or {}).get("hidden_size"))
    if hidden is None:
        raise KeyError("the Qwen2.5-VL config.json has no hidden_size")
    cfg["denoise_tower"] = {"model_type": "univa_denoise_tower", "denoiser_type": "flux", "denoise_projector_type": "mlp2x_gelu",
                            "input_hidden_size": hidden, "output_hidden_size": 4096, "denoiser_config": dict(flux_cfg)}
    return cfg


def assemble(qwen_dir, flux_dir, save_dir, dtype=None, seed: int = 0, max_shard_bytes: int = 5 << 30, log=print) -> dict:
    qwen_dir, flux_dir, save_dir = Path(qwen_dir), Path(flux_dir), Path(save_dir)
    save_dir.mkdir(parents=True, exist_ok=True)
    qcfg = json.loads((qwen_dir / "config.json").read_text())
    fcfg = json.loads((flux_dir / "transformer" / "config.json").read_text())
    cfg = univa_config(qcfg, fcfg)
    cast = (lambda t: t.to(dtype)) if dtype is not None else (lambda t: t)
    w = ShardWriter(save_dir, max_shard_bytes)
    n_q = n_f = 0
    for k, t in iter_tensors(qwen_dir):
        if k.startswith("denoise_tower."):
            raise KeyError(f"{qwen_dir} already holds {k}: not a plain Qwen2.5-VL checkpoint")
        w.add(k, cast(t))
        n_q += 1
    for k, t in iter_tensors(flux_dir / "transformer"):
        w.add("denoise_tower.denoiser." + k, cast(t))
        n_f += 1
    hidden, out = cfg["denoise_tower"]["input_hidden_size"], cfg["denoise_tower"]["output_hidden_size"]
    std = float(qcfg.get("initializer_range", (qcfg.get("text_config") or {}).get("initializer_range", 0.02)))
    g = torch.Generator().manual_seed(seed)
    pd = dtype or torch.float32
    w.add("denoise_tower.denoise_projector.0.weight", (torch.randn(3 * out, hidden, generator=g) * std).to(pd))
    w.add("denoise_tower.denoise_projector.0.bias", torch.zeros(3 * out, dtype=pd))
    w.add("denoise_tower.denoise_projector.2.weight", (torch.randn(out, 3 * out, generator=g) * std).to(pd))
    w.add("denoise_tower.denoise_projector.2.bias", torch.zeros(out, dtype=pd))
    weight_map = w.close()
    if dtype is not None:
        cfg["torch_dtype"] = str(dtype).replace("torch.", "")
    (save_dir / "config.json").write_text(json.dumps(cfg, indent=2))
    copied = []
    for name in PROCESSOR_FILES:
        if (qwen_dir / name).exists():
            shutil.copy2(qwen_dir / name, save_dir / name)
            copied.append(name)
    log(f"{save_dir}: {n_q} Qwen2.5-VL tensors, {n_f} FLUX tensors under denoise_tower.denoiser., MLP2 {hidden} -> {3 * out} -> {out} "
        f"(N(0, {std}), seed {seed}); {len(set(weight_map.values()))} shard(s), {w.total / 2 ** 30:.2f} GiB; processor files: {copied}")
    return {"tensors": len(weight_map), "shards": sorted(set(weight_map.values())), "processor_files": copied}


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--origin_flux_ckpt_path", type=str, required=True, help="FLUX checkpoint directory (holds transformer/)")
    ap.add_argument("--origin_qwenvl_ckpt_path", type=str, required=True, help="Qwen2.5-VL checkpoint directory")
    ap.add_argument("--save_path", type=str, required=True)
    ap.add_argument("--dtype", choices=sorted(DTYPES), default=None, help="cast every tensor (the reference saves float32)")
    ap.add_argument("--seed", type=int, default=0, help="seed of the MLP2 initialisation")
    a = ap.parse_args()
    assemble(a.origin_qwenvl_ckpt_path, a.origin_flux_ckpt_path, a.save_path, DTYPES.get(a.dtype), a.seed)
