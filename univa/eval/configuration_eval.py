"""Evaluation-run settings with the reference's field names and defaults
(univa/eval/configuration_eval.py:4-55; only the fields the GEdit sampling driver reads)."""
from __future__ import annotations

from dataclasses import dataclass, fields


@dataclass
class EvalConfig:
    pretrained_lvlm_name_or_path: str = ""
    pretrained_denoiser_name_or_path: str = ""
    joint_with_t5: bool = False
    only_use_t5: bool = False
    seed: int = 42
    output_dir: str = "./output"
    num_images_per_prompt: int = 1
    num_inference_steps: int = 32
    guidance_scale: float = 3.5
    height: int = 1024
    width: int = 1024
    min_pixels: int = 448 * 448
    max_pixels: int = 448 * 448
    local_rank: int = 0
    world_size: int = 1
    gedit_prompt_path: str = "univa/eval/gedit/gedit_edit.json"
    gedit_image_dir: str = ""
    # additions (no checkpoints / tokenizer files exist offline)
    synthetic: bool = False
    small: bool = False

    @classmethod
    def from_mapping(cls, m: dict) -> "EvalConfig":
        known = {f.name for f in fields(cls)}
        return cls(**{k: v for k, v in m.items() if k in known})   # other benchmarks' keys are ignored
