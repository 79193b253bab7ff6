"""`UnivaQwen2p5VLForConditionalGeneration` — the conditioning model surface of the reference
(univa/models/qwen2p5vl/modeling_univa_qwen2p5vl.py:32-47, 325-536, 604-621) over libb2f:

    model(input_ids, pixel_values=, attention_mask=, image_grid_thw=, output_type="denoise_embeds")
        -> [B, L, 4096]     Qwen2.5-VL prefill (ViT + decoder + final norm) -> denoise_projector (MLP2)
    model.denoise_tower.denoiser            the FLUX module handed to the pipeline (cli.py:126-128)
    model.denoise_tower.denoise_projector   MLP2
    model.forward_denoiser_context()        pass-through mode used by FSDP / log_validation (:352-355)

    model(..., output_type="lvlm", output_hidden_states=True)
        -> namespace(logits [B, L, vocab], hidden_states=(last,))   what cli.py:199-206 reads for the task head
    model.generate(input_ids, ..., max_new_tokens=128)              greedy KV-cache decode (cli.py:256-267)

One prefill serves both of cli.py's forward calls: `hidden_states[-1]` of the "lvlm" call is exactly the
pre-MLP2 tensor `prefill_hidden()` returns.
"""
from __future__ import annotations

from contextlib import contextmanager

import torch

from gpt_image_edit_b200._lib import B2FError
from gpt_image_edit_b200.qwen2p5vl import B200Qwen2p5VL, QwenTextConfig, QwenVisionConfig, get_rope_index

from ..configuration_univa_denoise_tower import UnivaDenoiseTowerConfig
from ..modeling_univa_denoise_tower import UnivaDenoiseTower


class UnivaQwen2p5VLConfig:
    model_type = "univa_qwen2p5vl"

    def __init__(self, denoise_tower: dict | UnivaDenoiseTowerConfig | None = None, text_config: dict | None = None,
                 vision_config: dict | None = None, image_token_id: int = 151655, video_token_id: int = 151656,
                 vision_start_token_id: int = 151652, shortcut_projector_type=None, shortcut_image_embeds=False, **kw):
        self.text_config = QwenTextConfig(image_token_id=image_token_id, video_token_id=video_token_id,
                                          vision_start_token_id=vision_start_token_id, **(text_config or {}))
        self.vision_config = QwenVisionConfig(**(vision_config or {}))
        if not isinstance(denoise_tower, UnivaDenoiseTowerConfig):
            dt = dict(denoise_tower or {})
            dt.setdefault("input_hidden_size", self.text_config.hidden_size)
            denoise_tower = UnivaDenoiseTowerConfig(**dt)
        self.denoise_tower = denoise_tower
        self.image_token_id, self.video_token_id = image_token_id, video_token_id
        self.vision_start_token_id = vision_start_token_id
        self.hidden_size = self.text_config.hidden_size
        if shortcut_image_embeds:
            # the reference's shortcut branch (:421-441, 507-520) cannot run there either (`len(num_blocks)` of an int) and no
            # shipped configuration enables it
            raise B2FError("shortcut_image_embeds=True is not built")
        self.shortcut_projector_type, self.shortcut_image_embeds = None, False       # :29-31: no projector without the flag

    def to_dict(self) -> dict:
        """config.json content in the layout the reference's config class writes (transformers 4.50: language-model fields at
        the top level, M-RoPE under `rope_scaling`, `vision_config` and `denoise_tower` sub-dicts)."""
        tc, vc, dt = self.text_config, self.vision_config, self.denoise_tower
        d = {"model_type": self.model_type, "architectures": ["UnivaQwen2p5VLForConditionalGeneration"], "torch_dtype": "bfloat16"}
        for k in ("hidden_size", "num_hidden_layers", "num_attention_heads", "num_key_value_heads", "intermediate_size", "vocab_size",
                  "rms_norm_eps", "rope_theta"):
            d[k] = getattr(tc, k)
        d["rope_scaling"] = {"type": "mrope", "mrope_section": list(tc.mrope_section)}
        d.update(image_token_id=self.image_token_id, video_token_id=self.video_token_id,
                 vision_start_token_id=self.vision_start_token_id, shortcut_image_embeds=False)
        d["vision_config"] = {k: (list(v) if isinstance(v, tuple) else v) for k, v in vars(vc).items()}
        d["vision_config"]["in_chans"] = vc.in_channels
        d["denoise_tower"] = {"model_type": dt.model_type, "denoiser_type": dt.denoiser_type,
                              "denoise_projector_type": dt.denoise_projector_type, "input_hidden_size": dt.input_hidden_size,
                              "output_hidden_size": dt.output_hidden_size,
                              "denoiser_config": {k: (list(v) if isinstance(v, tuple) else v) for k, v in dict(dt.denoiser_config).items()}}
        return d


class UnivaQwen2p5VLForConditionalGeneration(torch.nn.Module):
    config_class = UnivaQwen2p5VLConfig

    def __init__(self, config: UnivaQwen2p5VLConfig, device="cuda"):
        super().__init__()
        self.config = config
        self.lvlm = B200Qwen2p5VL(config.text_config, config.vision_config, device=device)
        self.denoise_tower = UnivaDenoiseTower(config.denoise_tower, device=device)
        self.forward_denoiser = False
        self.rope_deltas = None

    @property
    def dtype(self):
        return torch.bfloat16

    @property
    def device(self):
        return self.lvlm.device

    @classmethod
    def from_pretrained(cls, pretrained_model_name_or_path, torch_dtype=torch.bfloat16, attn_implementation=None, device="cuda",
                        **kwargs):
        """The reference's loading call (cli.py:37-41, gedit/step1_gen_samples.py:47-50, train_denoiser.py:1633):
        `from_pretrained(path, torch_dtype=torch.bfloat16, attn_implementation="flash_attention_2")`.  The engine computes in
        bf16 and has one attention implementation (its padding semantics are flash_attention_2's), so `torch_dtype` must be
        bf16 (or None) and `attn_implementation` is accepted for the signature's sake."""
        if torch_dtype not in (None, torch.bfloat16):
            raise B2FError(f"torch_dtype={torch_dtype}: the libb2f engine computes in bf16")
        from gpt_image_edit_b200.checkpoint import load_univa_model
        return load_univa_model(pretrained_model_name_or_path, device=device)

    def save_pretrained(self, save_directory, max_shard_size: int = 5 << 30, **kwargs):
        """`model.save_pretrained(dir)` of the reference's save hook (train_denoiser.py:492-494): safetensors shards with the
        checkpoint's key names (`visual.*`, `model.*`, `lm_head.weight`, `denoise_tower.denoiser.*`,
        `denoise_tower.denoise_projector.*`) and config.json."""
        from gpt_image_edit_b200.checkpoint import save_univa_model
        return save_univa_model(self, save_directory, max_shard_size)

    def get_rope_index(self, input_ids, image_grid_thw=None, video_grid_thw=None, second_per_grid_ts=None,
                       attention_mask=None):
        if video_grid_thw is not None:
            raise B2FError("video inputs are outside the image-editing hot path")
        c = self.config
        return get_rope_index(input_ids, image_grid_thw, attention_mask, spatial_merge_size=c.vision_config.spatial_merge_size,
                              image_token_id=c.image_token_id, vision_start_token_id=c.vision_start_token_id)

    def forward_visual(self, pixel_values, grid_thw):
        return self.lvlm.forward_visual(pixel_values, grid_thw)

    @torch.no_grad()
    def prefill_hidden(self, input_ids, pixel_values=None, attention_mask=None, image_grid_thw=None):
        """Last hidden state after the final norm, [B, L, hidden]."""
        return self.lvlm(input_ids, pixel_values=pixel_values, attention_mask=attention_mask, image_grid_thw=image_grid_thw)

    @torch.no_grad()
    def forward(self, input_ids=None, attention_mask=None, pixel_values=None, image_grid_thw=None,
                output_type: str = "lvlm", denoiser_kwargs=None, only_use_t5: bool = False,
                vlm_residual_image_factor: float = 0.0, **kwargs):
        if not only_use_t5 and self.forward_denoiser:           # reference :352-355
            return self.denoise_tower.denoiser(**kwargs)
        if output_type == "lvlm":
            # the understanding branch's forward: logits + the last hidden state (after the final norm; earlier layers'
            # states are not kept — the reference only reads hidden_states[-1], cli.py:201)
            from types import SimpleNamespace
            hidden = self.prefill_hidden(input_ids, pixel_values, attention_mask, image_grid_thw)
            B, L, _ = hidden.shape
            return SimpleNamespace(logits=self.lvlm.lm_logits(hidden).view(B, L, -1), hidden_states=(hidden,))
        if not output_type.startswith("denoise"):
            raise ValueError(f"Unknown output_type: {output_type}.")
        outputs = None
        if not only_use_t5:
            if vlm_residual_image_factor > 0.0 and pixel_values is not None:
                # reference :504-506: the decoder's outputs at the image-token positions are blended with the ViT features
                from gpt_image_edit_b200 import ops
                hidden, image_embeds = self.lvlm(input_ids, pixel_values=pixel_values, attention_mask=attention_mask,
                                                 image_grid_thw=image_grid_thw, return_image_embeds=True)
                flat = hidden.view(-1, hidden.shape[-1])
                where = (input_ids.to(flat.device).reshape(-1) == self.config.image_token_id).nonzero().squeeze(1).contiguous()
                old = ops.gather_rows(flat, where)
                f = float(vlm_residual_image_factor)
                ops.scatter_rows_(flat, where, ops.blend(old, image_embeds.contiguous(), 1.0 - f, f))
            else:
                hidden = self.prefill_hidden(input_ids, pixel_values, attention_mask, image_grid_thw)
            outputs = self.denoise_tower.denoise_projector(hidden)
        if output_type == "denoise_embeds":
            return outputs
        if output_type == "denoise_model_pred":
            kw = dict(denoiser_kwargs or {})
            assert len(kw) > 0, "denoiser_kwargs should not be empty when output_type is denoise_model_pred"
            kw["enc_attention_mask"] = attention_mask
            return self.denoise_tower(encoder_hidden_states=outputs, **kw)
        raise ValueError(f"Unknown output_type: {output_type}.")

    @torch.no_grad()
    def generate(self, input_ids=None, attention_mask=None, pixel_values=None, image_grid_thw=None, max_new_tokens: int = 128,
                 **kw):
        """ids [B, L + new], prompt included (transformers' `generate` contract; reference cli.py:258)."""
        return self.lvlm.generate(input_ids, pixel_values=pixel_values, attention_mask=attention_mask,
                                  image_grid_thw=image_grid_thw, max_new_tokens=max_new_tokens, **kw)

    @contextmanager
    def forward_denoiser_context(self):                          # reference :604-621
        """Pass-through mode: the model answers as its denoiser (`model.config` becomes the denoiser's config, so a
        pipeline built with `transformer=model` inside the context reads `.config.in_channels` etc.); yields the model."""
        prev, prev_cfg = self.forward_denoiser, self.config
        self.forward_denoiser = True
        self.config = self.denoise_tower.denoiser.config
        try:
            yield self
        finally:
            self.forward_denoiser = prev
            self.config = prev_cfg
