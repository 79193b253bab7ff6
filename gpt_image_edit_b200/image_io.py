"""Host-side image preparation shared by the CLI and the benchmarks (no GPU compute here).

`qwen_pixel_values` restates the patch layout of transformers' Qwen2VLImageProcessor (what
`AutoProcessor(min_pixels=max_pixels=448*448)` produces in reference cli.py:33-34, 190-197) for an
already-resized RGB image: CLIP mean/std normalisation, the frame repeated to temporal_patch_size,
patches of 14x14 emitted in 2x2-merge order, each row = (C, T, 14, 14) flattened = 1176 values.
"""
from __future__ import annotations

import numpy as np
import torch

OPENAI_CLIP_MEAN = (0.48145466, 0.4578275, 0.40821073)
OPENAI_CLIP_STD = (0.26862954, 0.26130258, 0.27577711)


def qwen_pixel_values(image_u8: np.ndarray, patch_size: int = 14, temporal_patch_size: int = 2, merge_size: int = 2):
    """image_u8: [H, W, 3] uint8 with H, W multiples of patch_size*merge_size.
    -> (pixel_values [gh*gw, 3*T*ps*ps] float32, grid_thw [[1, gh, gw]])."""
    H, W, _ = image_u8.shape
    ps, m, T = patch_size, merge_size, temporal_patch_size
    assert H % (ps * m) == 0 and W % (ps * m) == 0, "resize to a multiple of 28 first"
    x = image_u8.astype(np.float32) / 255.0
    x = (x - np.array(OPENAI_CLIP_MEAN, dtype=np.float32)) / np.array(OPENAI_CLIP_STD, dtype=np.float32)
    x = np.transpose(x, (2, 0, 1))[None]                       # [1, C, H, W]
    x = np.repeat(x, T, axis=0)                                # [T, C, H, W]  (a still image fills the temporal patch)
    gh, gw = H // ps, W // ps
    x = x.reshape(1, T, 3, gh // m, m, ps, gw // m, m, ps)
    x = x.transpose(0, 3, 6, 4, 7, 2, 1, 5, 8)                 # grid_t, gh/m, gw/m, m, m, C, T, ps, ps
    flat = x.reshape(gh * gw, 3 * T * ps * ps)
    return torch.from_numpy(np.ascontiguousarray(flat)), torch.tensor([[1, gh, gw]])


def image_to_condition_tensor(image_u8: np.ndarray) -> torch.Tensor:
    """[H,W,3] uint8 -> [1,3,H,W] float32 in [-1,1] (reference cli.py:99-116)."""
    t = torch.from_numpy(image_u8.astype(np.float32) / 255.0).permute(2, 0, 1)
    return ((t - 0.5) / 0.5)[None]


def resize_u8(image_u8: np.ndarray, height: int, width: int) -> np.ndarray:
    from PIL import Image

    return np.asarray(Image.fromarray(image_u8).resize((width, height), Image.BICUBIC))


# ------------------------------------------------------------------------------------------------
# `qwen_vl_utils.process_vision_info` (un-vendored dependency of the reference, requirements.txt:34, unpinned;
# algorithm as published in qwen-vl-utils 0.0.8-0.0.11 `vision_process.py`), images only.  Reference call sites:
# univa/serve/cli.py:189, univa/eval/gedit/step1_gen_samples.py:143.
IMAGE_FACTOR = 28
MIN_PIXELS = 4 * 28 * 28
MAX_PIXELS = 16384 * 28 * 28
MAX_RATIO = 200


def smart_resize(height: int, width: int, factor: int = IMAGE_FACTOR, min_pixels: int = MIN_PIXELS,
                 max_pixels: int = MAX_PIXELS):
    """(h, w): both multiples of `factor`, area within [min_pixels, max_pixels], aspect ratio kept as closely as the
    grid allows (the rule `AutoProcessor(min_pixels=max_pixels=448*448)` applies, reference cli.py:51-55)."""
    import math

    if max(height, width) / min(height, width) > MAX_RATIO:
        raise ValueError(f"absolute aspect ratio must be smaller than {MAX_RATIO}, got {max(height, width) / min(height, width)}")
    h_bar = max(factor, round(height / factor) * factor)
    w_bar = max(factor, round(width / factor) * factor)
    if h_bar * w_bar > max_pixels:
        beta = math.sqrt((height * width) / max_pixels)
        h_bar = max(factor, math.floor(height / beta / factor) * factor)
        w_bar = max(factor, math.floor(width / beta / factor) * factor)
    elif h_bar * w_bar < min_pixels:
        beta = math.sqrt(min_pixels / (height * width))
        h_bar = math.ceil(height * beta / factor) * factor
        w_bar = math.ceil(width * beta / factor) * factor
    return h_bar, w_bar


def fetch_image(ele: dict, size_factor: int = IMAGE_FACTOR):
    """One {"type": "image", "image": path | file:// | http(s):// | data:image;base64, | PIL.Image, [min_pixels,
    max_pixels | resized_height, resized_width]} content item -> RGB PIL image resized as qwen_vl_utils.fetch_image does
    (PIL's default bicubic)."""
    from PIL import Image

    image = ele.get("image", ele.get("image_url"))
    obj = None
    if isinstance(image, Image.Image):
        obj = image
    elif isinstance(image, str) and image.startswith(("http://", "https://")):
        import requests                                   # as qwen_vl_utils does; needs a network
        from io import BytesIO
        obj = Image.open(BytesIO(requests.get(image, stream=True).content))
    elif isinstance(image, str) and image.startswith("file://"):
        obj = Image.open(image[7:])
    elif isinstance(image, str) and image.startswith("data:image"):
        if "base64," in image:
            import base64
            from io import BytesIO
            obj = Image.open(BytesIO(base64.b64decode(image.split("base64,", 1)[1])))
    else:
        obj = Image.open(image)
    if obj is None:
        raise ValueError(f"Unrecognized image input, support local path, http url, base64 and PIL.Image, got {image}")
    obj = obj.convert("RGB")
    if "resized_height" in ele and "resized_width" in ele:
        rh, rw = smart_resize(ele["resized_height"], ele["resized_width"], factor=size_factor)
    else:
        w, h = obj.size
        rh, rw = smart_resize(h, w, factor=size_factor, min_pixels=ele.get("min_pixels", MIN_PIXELS),
                              max_pixels=ele.get("max_pixels", MAX_PIXELS))
    return obj.resize((rw, rh))


def process_vision_info(conversation: list):
    """(image_inputs | None, video_inputs = None) for a list of chat messages, in message order."""
    images = []
    for message in conversation:
        content = message.get("content")
        if not isinstance(content, list):
            continue
        for ele in content:
            if "image" in ele or "image_url" in ele or ele.get("type") in ("image", "image_url"):
                images.append(fetch_image(ele))
            elif "video" in ele or ele.get("type") == "video":
                raise ValueError("video inputs are not supported (the reference's serving path is image-only)")
    return (images or None), None
