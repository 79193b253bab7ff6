"""Special-token names of the reference (univa/utils/constant.py:1-18) for the model family that is built."""
SPACIAL_TOKEN = {
    "qwen2p5vl": {
        "image_token": "<|image_pad|>",
        "image_begin_token": "<|vision_start|>",
        "image_end_token": "<|vision_end|>",
    },
}
GENERATE_TOKEN = "<gen_image>"
