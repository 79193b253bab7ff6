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
