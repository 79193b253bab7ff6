"""Model registry of the reference (univa/models/__init__.py:1-9).  Only the Qwen2.5-VL variant is on the hot path and built
(every shipped configuration selects `qwen2p5vl`, scripts/denoiser/*.yaml); the legacy Qwen2 / Qwen2-VL variants are out of
scope (component 8) and absent from the registry."""


def __getattr__(name):          # lazily: importing the package must not load libb2f
    if name == "UnivaQwen2p5VLForConditionalGeneration":
        from .qwen2p5vl.modeling_univa_qwen2p5vl import UnivaQwen2p5VLForConditionalGeneration
        return UnivaQwen2p5VLForConditionalGeneration
    if name == "MODEL_TYPE":
        from .qwen2p5vl.modeling_univa_qwen2p5vl import UnivaQwen2p5VLForConditionalGeneration
        return {"qwen2p5vl": UnivaQwen2p5VLForConditionalGeneration}
    raise AttributeError(name)
