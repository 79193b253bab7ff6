"""Config of the denoise tower (reference univa/models/configuration_univa_denoise_tower.py:7-31):
same field names; `denoiser_config` is the diffusers FluxTransformer2DModel config dict (or a JSON path)."""
from __future__ import annotations

import json
from types import SimpleNamespace


class UnivaDenoiseTowerConfig(SimpleNamespace):
    model_type = "univa_denoise_tower"

    def __init__(self, denoiser_type: str = "flux", denoise_projector_type: str = "mlp2x_gelu",
                 input_hidden_size: int = 1152, output_hidden_size: int = 4096, denoiser_config=None, **kw):
        if isinstance(denoiser_config, str):
            with open(denoiser_config) as f:
                denoiser_config = json.load(f)
        super().__init__(denoiser_type=denoiser_type, denoise_projector_type=denoise_projector_type,
                         input_hidden_size=input_hidden_size, output_hidden_size=output_hidden_size,
                         denoiser_config=denoiser_config or {}, **kw)
