"""`univa.utils.flux_pipeline` — same import path as the reference's vendored FluxKontextPipeline
(reference univa/utils/flux_pipeline.py); the implementation lives in gpt_image_edit_b200.pipeline."""
from gpt_image_edit_b200.pipeline import (  # noqa: F401
    PREFERRED_KONTEXT_RESOLUTIONS,
    FluxKontextPipeline,
    FluxPipeline,
    FluxPipelineOutput,
    VaeImageProcessor,
    calculate_shift,
    retrieve_timesteps,
)
