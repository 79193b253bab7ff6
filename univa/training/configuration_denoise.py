"""Configuration schema of `train_denoiser.py` — the field names and defaults of the reference's dataclasses
(univa/training/configuration_denoise.py:6-154), so the reference's yaml files load unchanged, plus a small
yaml -> dataclass loader standing in for `OmegaConf.merge(OmegaConf.structured(...), OmegaConf.load(path))`
(train_denoiser.py:1626-1631; omegaconf is not available offline): unknown keys are errors, scalars are coerced to
the annotated type (yaml reads `1e-8` as a string).

Additions, all optional and all off by default: `model_config.synthetic` / `model_config.small` (seeded random weights
at the real or a reduced architecture: no checkpoints exist offline) and `dataset_config.dataset_type: "synthetic"`
(synthetic (source image, instruction, target image) triples, BASELINE.json configs[3]).
"""
from __future__ import annotations

import dataclasses
import typing
from dataclasses import dataclass, field
from typing import List, Optional


@dataclass
class TrainingConfig:
    seed: int = 42
    wandb_project: str = "univa-denoiser"
    wandb_name: str = "default_config"
    output_dir: str = "./output"
    logging_dir: str = "./logs"
    gradient_accumulation_steps: int = 1
    learning_rate: float = 1e-5
    adam_beta1: float = 0.9
    adam_beta2: float = 0.999
    adam_epsilon: float = 1e-8
    adam_weight_decay: float = 1e-2
    mixed_precision: str = "bf16"
    report_to: str = "wandb"
    gradient_checkpointing: bool = False
    num_train_epochs: int = 1
    max_train_steps: Optional[int] = None
    lr_scheduler: str = "constant"
    lr_warmup_steps: int = 0
    lr_num_cycles: int = 1
    lr_power: float = 1.0
    resume_from_checkpoint: Optional[str] = None
    weighting_scheme: Optional[str] = "logit_normal"   # ["sigma_sqrt", "logit_normal", "mode", "cosmap", "null"]
    logit_mean: float = 0.0
    logit_std: float = 1.0
    mode_scale: float = 1.29
    max_grad_norm: float = 1.0
    checkpointing_steps: int = 100
    checkpoints_total_limit: Optional[int] = 500
    drop_condition_rate: float = 0.0
    drop_t5_rate: float = 1.0
    validation_steps: int = 100
    num_validation_images: int = 1
    noise_reference_images: bool = False
    mask_weight_type: Optional[str] = None             # ['log', 'exp']
    sigmas_as_weight: bool = False
    discrete_timestep: bool = True
    optimizer: str = "adamw"                           # ['adamw', 'prodigy']
    prodigy_use_bias_correction: bool = True
    prodigy_safeguard_warmup: bool = True
    prodigy_decouple: bool = True
    prodigy_beta3: Optional[float] = None
    prodigy_d_coef: float = 1.0
    profile_out_dir: Optional[str] = None
    ema_deepspeed_config_file: Optional[str] = None
    ema_update_freq: int = 1
    ema_decay: float = 0.99


@dataclass
class DatasetConfig:
    dataset_type: str = "synthetic"
    data_txt: str = ""
    batch_size: int = 16
    num_workers: int = 4
    height: int = 512
    width: int = 512
    min_pixels: int = 448 * 448
    max_pixels: int = 448 * 448
    anyres: str = "any_1ratio"
    ocr_enhancer: bool = False
    random_data: bool = False
    padding_side: str = "right"
    validation_t2i_prompt: Optional[str] = None
    validation_it2i_prompt: Optional[str] = None
    validation_image_path: Optional[str] = None
    pin_memory: bool = True
    validation_iit2i_prompt: Optional[str] = None
    validation_iit2i_path: Optional[List[str]] = None
    validation_REFiit2i_prompt: Optional[str] = None
    validation_REFiit2i_path: Optional[List[str]] = None
    validation_cannyt2i_prompt: Optional[str] = None
    validation_cannyt2i_path: Optional[str] = None
    validation_poset2i_prompt: Optional[str] = None
    validation_poset2i_path: Optional[str] = None
    validation_it2pose_prompt: Optional[str] = None
    validation_it2pose_path: Optional[str] = None
    validation_it2canny_prompt: Optional[str] = None
    validation_it2canny_path: Optional[str] = None
    validation_NIKEit2i_prompt: Optional[str] = None
    validation_NIKEit2i_path: Optional[str] = None
    validation_TRANSFERit2i_prompt: Optional[str] = None
    validation_TRANSFERit2i_path: Optional[str] = None
    validation_EXTRACTit2i_prompt: Optional[str] = None
    validation_EXTRACTit2i_path: Optional[str] = None
    validation_TRYONit2i_prompt: Optional[str] = None
    validation_TRYONit2i_path: Optional[str] = None
    validation_REPLACEit2i_prompt: Optional[str] = None
    validation_REPLACEit2i_path: Optional[str] = None
    validation_DETit2i_prompt: Optional[str] = None
    validation_DETit2i_path: Optional[str] = None
    validation_SEGit2i_prompt: Optional[str] = None
    validation_SEGit2i_path: Optional[str] = None
    # addition: length of the synthetic dataset (dataset_type: "synthetic")
    synthetic_len: int = 1 << 30
    # addition: target sizes [[H, W], ...] the synthetic samples cycle through (mixed-size batches, batch_size != 1)
    synthetic_target_sizes: Optional[list] = None


@dataclass
class ModelConfig:
    pretrained_lvlm_name_or_path: str = ""
    pretrained_denoiser_name_or_path: str = ""
    guidance_scale: float = 1.0
    tune_mlp1_only: bool = False
    pretrained_mlp1_path: Optional[str] = None
    with_tune_mlp2: bool = False
    only_tune_mlp2: bool = False
    pretrained_mlp2_path: Optional[str] = None
    only_tune_image_branch: bool = True
    flux_train_layer_idx: Optional[list] = None
    joint_ref_feature: bool = False
    joint_ref_feature_as_condition: bool = False
    only_use_t5: bool = False
    vlm_residual_image_factor: float = 0.0
    vae_fp32: bool = True
    compile_flux: bool = False
    compile_qwen2p5vl: bool = False
    ema_pretrained_lvlm_name_or_path: Optional[str] = None
    # additions (offline): seeded random weights, optionally with a few layers only
    synthetic: bool = False
    small: bool = False


@dataclass
class UnivaTrainingDenoiseConfig:
    training_config: TrainingConfig = field(default_factory=TrainingConfig)
    dataset_config: DatasetConfig = field(default_factory=DatasetConfig)
    model_config: ModelConfig = field(default_factory=ModelConfig)


def _coerce(value, tp, where):
    origin = typing.get_origin(tp)
    if origin is typing.Union:                          # Optional[X]
        args = [a for a in typing.get_args(tp) if a is not type(None)]
        if value is None:
            return None
        return _coerce(value, args[0], where)
    if value is None:
        return None
    if tp is bool:
        if isinstance(value, bool):
            return value
        if isinstance(value, str) and value.lower() in ("true", "false"):
            return value.lower() == "true"
        raise TypeError(f"{where}: expected a bool, got {value!r}")
    if tp is int:
        if isinstance(value, bool) or (isinstance(value, float) and value != int(value)):
            raise TypeError(f"{where}: expected an int, got {value!r}")
        return int(value)
    if tp is float:
        return float(value)
    if tp is str:
        return str(value)
    if tp is list or origin is list:
        if not isinstance(value, (list, tuple)):
            raise TypeError(f"{where}: expected a list, got {value!r}")
        return list(value)
    return value


def _build(cls, mapping, where):
    mapping = dict(mapping or {})
    hints = typing.get_type_hints(cls)
    names = {f.name for f in dataclasses.fields(cls)}
    unknown = sorted(set(mapping) - names)
    if unknown:
        raise KeyError(f"{where}: unknown keys {unknown} (the schema is the reference's configuration_denoise.py)")
    return cls(**{k: _coerce(v, hints[k], f"{where}.{k}") for k, v in mapping.items()})


def from_mapping(raw: dict) -> UnivaTrainingDenoiseConfig:
    raw = dict(raw or {})
    unknown = sorted(set(raw) - {"training_config", "dataset_config", "model_config"})
    if unknown:
        raise KeyError(f"unknown top-level sections {unknown}")
    return UnivaTrainingDenoiseConfig(training_config=_build(TrainingConfig, raw.get("training_config"), "training_config"),
                                      dataset_config=_build(DatasetConfig, raw.get("dataset_config"), "dataset_config"),
                                      model_config=_build(ModelConfig, raw.get("model_config"), "model_config"))


def load_config(path) -> UnivaTrainingDenoiseConfig:
    import yaml

    with open(path) as f:
        return from_mapping(yaml.safe_load(f) or {})
