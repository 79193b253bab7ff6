"""Evaluation-run settings: the reference's schema (univa/eval/configuration_eval.py:4-55), every field with its name
and default, so a yaml written for the reference loads here unchanged and a key outside the schema is rejected as
OmegaConf's structured merge rejects it.  Only the GEdit sampling driver is built in this repo; the other benchmarks'
fields are carried, not read."""
from __future__ import annotations

from dataclasses import dataclass, fields


@dataclass
class EvalConfig:
    pretrained_lvlm_name_or_path: str = ""
    pretrained_denoiser_name_or_path: str = ""
    pretrained_siglip_name_or_path: str = ""

    ocr_enhancer: bool = False
    joint_with_t5: bool = False
    only_use_t5: bool = False

    seed: int = 42
    allow_tf32: bool = False

    output_dir: str = "./output"

    num_images_per_prompt: int = 1
    num_inference_steps: int = 32
    guidance_scale: float = 3.5
    num_samples_per_prompt: int = 1
    height: int = 1024
    width: int = 1024
    min_pixels: int = 448 * 448
    max_pixels: int = 448 * 448
    anyres: str = "any_11ratio"
    padding_side: str = "right"

    local_rank: int = 0
    world_size: int = 1

    genai_prompt_path: str = "univa/eval/genai/eval_prompts/genai527/genai_image.json"
    n_samples: int = 4
    geneval_prompt_path: str = "univa/eval/geneval/evaluation_metadata.jsonl"
    resized_height: int = 1024
    resized_width: int = 1024
    dpgbench_prompt_path: str = "univa/eval/dpgbench/dpgbench_prompts.json"
    wise_prompt_path: str = "univa/eval/wise/data"
    imgedit_prompt_path: str = "univa/eval/imgedit/basic_edit.json"
    imgedit_image_dir: str = "/mnt/data/lb/Remake/imgedit_bench_eval_images"
    gedit_prompt_path: str = "univa/eval/gedit/gedit_edit.json"
    gedit_image_dir: str = "/mnt/data/lb/Remake/gedit_bench_eval_images"

    # additions (no checkpoints / tokenizer files exist offline)
    synthetic: bool = False
    small: bool = False

    @classmethod
    def from_mapping(cls, m: dict) -> "EvalConfig":
        known = {f.name: f for f in fields(cls)}
        unknown = sorted(set(m) - set(known))
        if unknown:
            raise KeyError(f"keys outside the EvalConfig schema: {unknown}")
        out = cls()
        for k, v in m.items():
            tp = known[k].type
            if tp in ("int", "float") and isinstance(v, str):
                v = float(v) if tp == "float" else int(v)         # PyYAML reads `1e-8`-style numbers as strings
            if tp == "float" and isinstance(v, int) and not isinstance(v, bool):
                v = float(v)
            setattr(out, k, v)
        return out
