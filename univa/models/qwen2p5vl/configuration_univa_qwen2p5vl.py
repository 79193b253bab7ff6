"""Import path of the reference's config class (univa/models/qwen2p5vl/configuration_univa_qwen2p5vl.py:7-52); the class lives
next to the model in this repo."""
from .modeling_univa_qwen2p5vl import UnivaQwen2p5VLConfig  # noqa: F401
