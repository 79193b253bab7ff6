"""safetensors -> device loaders with diffusers / HF key names (SURVEY.md §8f-4, §5 "Checkpoint").

Directory layouts read (those the reference's `from_pretrained` calls consume, cli.py:37-41, 64-68):
  flux_path/transformer/*.safetensors        diffusers FluxTransformer2DModel keys
  flux_path/vae/*.safetensors                diffusers AutoencoderKL keys
  flux_path/scheduler/scheduler_config.json  FlowMatchEulerDiscreteScheduler config
  flux_path/text_encoder/*.safetensors       transformers CLIPTextModel keys (+ config.json), tokenizer/ beside it
  flux_path/text_encoder_2/*.safetensors     transformers T5EncoderModel keys (+ config.json), tokenizer_2/ beside it
  model_path/*.safetensors                   Univa checkpoint: visual.*, model.*, denoise_tower.denoiser.*,
                                             denoise_tower.denoise_projector.{0,2}.*  (train_denoiser.py:112-115)
  model_path/task_head_final.pt              cli.py:49
No real checkpoint exists offline; tests round-trip synthetic ones written with `save_state_dict`.
"""
from __future__ import annotations

import json
from pathlib import Path

import torch
from safetensors import safe_open
from safetensors.torch import save_file

from .scheduler import FlowMatchEulerDiscreteScheduler


def iter_safetensors(directory, prefix: str = ""):
    """Yields (key-without-prefix, tensor) over every *.safetensors shard in `directory`."""
    d = Path(directory)
    files = sorted(d.glob("*.safetensors"))
    if not files:
        raise FileNotFoundError(f"no *.safetensors under {d}")
    for f in files:
        with safe_open(str(f), framework="pt", device="cpu") as sf:
            for k in sf.keys():
                if k.startswith(prefix):
                    yield k[len(prefix):], sf.get_tensor(k)


def load_state_dict_from_dir(directory, prefix: str = "") -> dict:
    return dict(iter_safetensors(directory, prefix))


def save_state_dict(sd: dict, directory, filename: str = "diffusion_pytorch_model.safetensors", prefix: str = ""):
    Path(directory).mkdir(parents=True, exist_ok=True)
    save_file({prefix + k: v.detach().cpu().contiguous() for k, v in sd.items()}, str(Path(directory) / filename))


class ShardWriter:
    """Writes tensors to `model.safetensors`, or `model-0000i-of-0000n.safetensors` + `model.safetensors.index.json` when more
    than `max_bytes` accumulate (the layout `save_pretrained` produces and `from_pretrained` / iter_safetensors read)."""

    def __init__(self, out, max_bytes: int = 5 << 30):
        self.out, self.max_bytes, self.cur, self.cur_bytes, self.shards, self.total = Path(out), max_bytes, {}, 0, [], 0
        self.out.mkdir(parents=True, exist_ok=True)

    def add(self, name: str, t: torch.Tensor):
        n = t.numel() * t.element_size()
        if self.cur and self.cur_bytes + n > self.max_bytes:
            self.flush()
        self.cur[name] = t.detach().to("cpu", copy=True).contiguous()       # an owned copy: safetensors refuses shared storage
        self.cur_bytes += n
        self.total += n

    def flush(self):
        if self.cur:
            tmp = self.out / f"model-tmp-{len(self.shards):05d}.safetensors"
            save_file(self.cur, str(tmp), metadata={"format": "pt"})
            self.shards.append((tmp, list(self.cur)))
            self.cur, self.cur_bytes = {}, 0

    def close(self) -> dict:
        self.flush()
        n = len(self.shards)
        weight_map = {}
        for i, (tmp, names) in enumerate(self.shards):
            final = "model.safetensors" if n == 1 else f"model-{i + 1:05d}-of-{n:05d}.safetensors"
            tmp.rename(self.out / final)
            weight_map.update({k: final for k in names})
        if n > 1:
            (self.out / "model.safetensors.index.json").write_text(
                json.dumps({"metadata": {"total_size": self.total}, "weight_map": weight_map}, indent=2))
        return weight_map


CHECKPOINT_SIDE_FILES = ("config.json", "generation_config.json", "tokenizer.json", "tokenizer_config.json", "vocab.json",
                         "merges.txt", "added_tokens.json", "special_tokens_map.json", "preprocessor_config.json",
                         "processor_config.json", "chat_template.json", "chat_template.jinja", "task_head_final.pt")


def rewrite_checkpoint(src, dst, updates: dict, max_shard_bytes: int = 5 << 30) -> dict:
    """A copy of the checkpoint directory `src` at `dst` in which the tensors named in `updates` carry their new values
    (cast to the stored dtype): what `save_pretrained(checkpoint-N/univa)` + `processor.save_pretrained` leave after a
    training step in the reference (train_denoiser.py:489-498) — a directory `from_pretrained` / load_univa_checkpoint reads
    — produced by streaming the source shards instead of serialising a second copy of the model.  Every update must name a
    tensor of the source.  Returns {name: shard file}."""
    src, dst = Path(src), Path(dst)
    w = ShardWriter(dst, max_shard_bytes)
    seen = set()
    for k, t in iter_safetensors(src):
        if k in updates:
            u = updates[k]
            if tuple(u.shape) != tuple(t.shape):
                raise ValueError(f"{k}: the update has shape {tuple(u.shape)}, the checkpoint {tuple(t.shape)}")
            t = u.detach().to("cpu", t.dtype)
            seen.add(k)
        w.add(k, t)
    missing = sorted(set(updates) - seen)
    if missing:
        raise KeyError(f"{src} has no tensors named {missing[:4]} ({len(missing)} updates without a home)")
    weight_map = w.close()
    for name in CHECKPOINT_SIDE_FILES:
        if (src / name).exists():
            (dst / name).write_bytes((src / name).read_bytes())
    return weight_map


def load_pipeline_components(flux_path, device="cuda"):
    """-> (vae, scheduler) for FluxKontextPipeline.from_pretrained."""
    from .vae import B200AutoencoderKL, VaeConfig

    root = Path(flux_path)
    cfg = {}
    if (root / "vae" / "config.json").exists():
        raw = json.loads((root / "vae" / "config.json").read_text())
        cfg = {k: raw[k] for k in ("block_out_channels", "layers_per_block", "latent_channels", "scaling_factor",
                                   "shift_factor", "in_channels", "out_channels") if k in raw}
        if "block_out_channels" in cfg:
            cfg["block_out_channels"] = tuple(cfg["block_out_channels"])
    vae = B200AutoencoderKL(VaeConfig(**cfg), device=device)
    vae.load_state_dict(load_state_dict_from_dir(root / "vae"))
    scfg = {}
    if (root / "scheduler" / "scheduler_config.json").exists():
        raw = json.loads((root / "scheduler" / "scheduler_config.json").read_text())
        scfg = {k: raw[k] for k in ("num_train_timesteps", "shift", "use_dynamic_shifting", "base_shift", "max_shift",
                                    "base_image_seq_len", "max_image_seq_len") if k in raw}
    return vae, FlowMatchEulerDiscreteScheduler(**scfg)


def _tokenizer(directory, cls_name):
    """Local-files tokenizer (transformers) when the directory exists; None otherwise (no network here)."""
    if not Path(directory).is_dir():
        return None
    import transformers

    return getattr(transformers, cls_name).from_pretrained(str(directory), local_files_only=True)


def load_text_encoders(flux_path, device="cuda"):
    """-> (clip, clip_tokenizer, t5, t5_tokenizer) from the FLUX.1 directory layout the reference's
    `FluxKontextPipeline.from_pretrained` reads (cli.py:64-76); a missing sub-directory yields None."""
    from .text_encoders import B200CLIPTextModel, B200T5Encoder, CLIPTextConfig, T5EncoderConfig

    root = Path(flux_path)
    clip = t5 = None
    if (root / "text_encoder").is_dir():
        raw = json.loads((root / "text_encoder" / "config.json").read_text()) if (root / "text_encoder" / "config.json").exists() else {}
        keys = ("vocab_size", "hidden_size", "intermediate_size", "num_hidden_layers", "num_attention_heads",
                "max_position_embeddings", "layer_norm_eps", "eos_token_id")
        clip = B200CLIPTextModel(CLIPTextConfig(**{k: raw[k] for k in keys if k in raw}), device=device)
        clip.load_state_dict(load_state_dict_from_dir(root / "text_encoder"))
    if (root / "text_encoder_2").is_dir():
        raw = json.loads((root / "text_encoder_2" / "config.json").read_text()) if (root / "text_encoder_2" / "config.json").exists() else {}
        keys = ("vocab_size", "d_model", "d_kv", "num_heads", "d_ff", "num_layers", "relative_attention_num_buckets",
                "relative_attention_max_distance", "layer_norm_epsilon")
        t5 = B200T5Encoder(T5EncoderConfig(**{k: raw[k] for k in keys if k in raw}), device=device)
        t5.load_state_dict(load_state_dict_from_dir(root / "text_encoder_2"))
    return clip, _tokenizer(root / "tokenizer", "CLIPTokenizer"), t5, _tokenizer(root / "tokenizer_2", "T5TokenizerFast")


def load_flux_transformer(directory, device="cuda", prefix: str = ""):
    from .flux_transformer import B200FluxTransformer2DModel, FluxTransformerConfig

    d = Path(directory)
    cfg = {}
    if (d / "config.json").exists():
        raw = json.loads((d / "config.json").read_text())
        cfg = {k: raw[k] for k in ("in_channels", "num_layers", "num_single_layers", "attention_head_dim",
                                   "num_attention_heads", "joint_attention_dim", "pooled_projection_dim",
                                   "guidance_embeds", "axes_dims_rope") if k in raw}
    m = B200FluxTransformer2DModel(FluxTransformerConfig(**cfg), device=device)
    m.load_state_dict(load_state_dict_from_dir(d, prefix))
    return m


PROCESSOR_FILES = ("tokenizer.json", "tokenizer_config.json", "vocab.json", "preprocessor_config.json", "processor_config.json")


def load_processor(model_path, min_pixels=448 * 448, max_pixels=448 * 448):
    """`AutoProcessor.from_pretrained(model_path, min_pixels=, max_pixels=)` (reference cli.py:51-55): tokenizer, chat
    template and Qwen2-VL image processor from the files next to the checkpoint.  Raises when they are missing — without
    them the instruction text cannot reach the VLM."""
    root = Path(model_path)
    if not any((root / f).exists() for f in PROCESSOR_FILES):
        raise FileNotFoundError(f"{root}: no tokenizer / processor files ({', '.join(PROCESSOR_FILES)}); the chat template and "
                                "tokenizer of the Univa checkpoint are required (reference cli.py:51-55)")
    from transformers import AutoProcessor

    return AutoProcessor.from_pretrained(str(root), min_pixels=min_pixels, max_pixels=max_pixels)


TEXT_KEYS = ("hidden_size", "num_hidden_layers", "num_attention_heads", "num_key_value_heads", "intermediate_size",
             "vocab_size", "rms_norm_eps", "rope_theta")
VISION_KEYS = ("depth", "hidden_size", "num_heads", "intermediate_size", "patch_size", "temporal_patch_size",
               "spatial_merge_size", "window_size", "fullatt_block_indexes", "out_hidden_size", "tokens_per_second")


def univa_config_kwargs(raw: dict) -> dict:
    """Keyword arguments of UnivaQwen2p5VLConfig from a checkpoint's config.json.  The reference's config class extends
    transformers-4.50's Qwen2_5_VLConfig (configuration_univa_qwen2p5vl.py:7-52), whose json keeps the language-model
    fields at the TOP level and M-RoPE under `rope_scaling`; transformers-5 files nest them under `text_config` /
    `rope_parameters`.  Both are read; a field that is absent keeps the Qwen2.5-VL-7B default."""
    nested = dict(raw.get("text_config") or {})
    text = {k: nested.get(k, raw.get(k)) for k in TEXT_KEYS if nested.get(k, raw.get(k)) is not None}
    rope = nested.get("rope_parameters") or nested.get("rope_scaling") or raw.get("rope_parameters") or raw.get("rope_scaling") or {}
    if rope.get("mrope_section") is not None:
        text["mrope_section"] = tuple(rope["mrope_section"])
    if rope.get("rope_theta") is not None and "rope_theta" not in text:
        text["rope_theta"] = rope["rope_theta"]
    vraw = dict(raw.get("vision_config") or {})
    vision = {k: vraw[k] for k in VISION_KEYS if k in vraw}
    if "in_chans" in vraw or "in_channels" in vraw:
        vision["in_channels"] = vraw.get("in_channels", vraw.get("in_chans"))
    if "fullatt_block_indexes" in vision:
        vision["fullatt_block_indexes"] = tuple(vision["fullatt_block_indexes"])
    kw = dict(denoise_tower=raw.get("denoise_tower"), text_config=text, vision_config=vision)
    for k in ("image_token_id", "video_token_id", "vision_start_token_id", "shortcut_image_embeds", "shortcut_projector_type"):
        if raw.get(k) is not None:
            kw[k] = raw[k]
    return kw


def univa_state_dict(model) -> dict:
    """Every tensor of a Univa model under its checkpoint name (the keys `save_pretrained` writes in the reference)."""
    sd = dict(model.lvlm.state_dict())
    sd.update({"denoise_tower.denoiser." + k: v for k, v in model.denoise_tower.denoiser.state_dict().items()})
    proj = getattr(model.denoise_tower, "denoise_projector", None)
    if proj is not None:
        sd.update({"denoise_tower.denoise_projector." + k: v for k, v in proj.state_dict().items()})
    return sd


def save_univa_model(model, save_directory, max_shard_bytes: int = 5 << 30) -> dict:
    """safetensors shards + config.json of a Univa model (`UnivaQwen2p5VLForConditionalGeneration.save_pretrained`)."""
    out = Path(save_directory)
    w = ShardWriter(out, max_shard_bytes)
    for k, v in univa_state_dict(model).items():
        w.add(k, v)
    weight_map = w.close()
    (out / "config.json").write_text(json.dumps(model.config.to_dict(), indent=2))
    return weight_map


def load_univa_model(model_path, device="cuda"):
    """UnivaQwen2p5VLForConditionalGeneration from a checkpoint directory (config.json + safetensors shards)."""
    return _load_univa_model(Path(model_path), device)[0]


def load_univa_checkpoint(model_path, device="cuda", min_pixels=448 * 448, max_pixels=448 * 448, task_head: bool = True):
    """-> (UnivaQwen2p5VLForConditionalGeneration, task_head, processor) from a Univa checkpoint directory
    (reference cli.py:30-56).  `task_head=False` is the eval drivers' load (gedit/step1_gen_samples.py:45-56: model and
    processor only; no `task_head_final.pt` is read) and returns None in its place."""
    from univa.serve.cli import TaskHead

    root = Path(model_path)
    processor = load_processor(root, min_pixels, max_pixels)
    model, cfg = _load_univa_model(root, device)
    if not task_head:
        return model, None, processor
    head = TaskHead(cfg.hidden_size, device=device)
    th = root / "task_head_final.pt"
    if not th.exists():      # the reference's torch.load (cli.py:49) fails here; a zero head would answer every turn with text
        raise FileNotFoundError(f"{th}: the generate / understand router of the Univa checkpoint is missing")
    t = torch.load(th, map_location="cpu")
    head.w0.copy_(t["0.weight"]); head.b0.copy_(t["0.bias"])
    head.w3[:2].copy_(t["3.weight"]); head.b3[:2].copy_(t["3.bias"])
    return model, head, processor


def _load_univa_model(root: Path, device):
    from univa.models.qwen2p5vl.modeling_univa_qwen2p5vl import UnivaQwen2p5VLConfig, UnivaQwen2p5VLForConditionalGeneration

    raw = json.loads((root / "config.json").read_text()) if (root / "config.json").exists() else {}
    cfg = UnivaQwen2p5VLConfig(**univa_config_kwargs(raw))
    model = UnivaQwen2p5VLForConditionalGeneration(cfg, device=device)
    sd = load_state_dict_from_dir(root)
    lvlm_sd = {k: v for k, v in sd.items() if k.startswith(("visual.", "model.", "lm_head."))}
    if "lm_head.weight" not in lvlm_sd and raw.get("tie_word_embeddings", False) and "model.embed_tokens.weight" in lvlm_sd:
        lvlm_sd["lm_head.weight"] = lvlm_sd["model.embed_tokens.weight"]
    res = model.lvlm.load_state_dict(lvlm_sd)
    missing = list(getattr(res, "missing_keys", []) or [])
    if missing:      # e.g. lm_head.weight: the text-reply branch would emit argmax = 0 at every step
        raise KeyError(f"{root}: the checkpoint lacks tensors of the Qwen2.5-VL tower: {missing[:4]}")
    model.denoise_tower.denoiser.load_state_dict({k[len("denoise_tower.denoiser."):]: v for k, v in sd.items()
                                                  if k.startswith("denoise_tower.denoiser.")})
    model.denoise_tower.denoise_projector.load_state_dict({k[len("denoise_tower.denoise_projector."):]: v for k, v in sd.items()
                                                           if k.startswith("denoise_tower.denoise_projector.")})
    return model, cfg
