"""Import-path shim: `univa.utils.denoiser_prompt_embedding_flux` of the reference, served by the
libb2f-backed implementation (gpt_image_edit_b200/text_encoders.py)."""
from gpt_image_edit_b200.text_encoders import (  # noqa: F401
    _encode_prompt_with_clip,
    _encode_prompt_with_t5,
    encode_prompt,
    tokenize_prompt,
)
