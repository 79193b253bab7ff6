"""Synthetic (source image, instruction, target image) triples with the batch keys the reference's training loop reads
(train_denoiser.py:829-858): BASELINE.json configs[3] names synthetic triples, and no dataset exists offline.  The
reference's real dataset code (univa/dataset/*) is out of scope.

Per sample (seeded by its index, so every rank / run sees the same stream):
  generated_image   [3, H, W] fp32 in [-1, 1]   the edit target
  ref_pixel_values  [3, H, W] fp32 in [-1, 1]   the source image the edit starts from (FLUX-Kontext context)
  input_ids / attention_mask / pixel_values / image_grid_thw   the VLM prompt: a 448x448 view of the source (1024 patch
                    rows -> 256 image tokens) + instruction tokens in the canonical layout of SURVEY.md section 8d
  weights           [1, h, w] fp32: the area-mask weights of `mask_weight_type: log` (a random edited box weighted by
                    log2(area ratio) + 1 as univa/utils/get_mask.py does for its masks; 1 outside)
"""
from __future__ import annotations

import math

import torch


class SyntheticEditDataset(torch.utils.data.Dataset):
    def __init__(self, height: int = 512, width: int = 512, length: int = 1 << 30, seed: int = 0, n_text: int = 22,
                 vocab: int = 152064):
        self.h, self.w, self.length, self.seed, self.n_text, self.vocab = height, width, length, seed, n_text, vocab

    def __len__(self):
        return self.length

    def __getitem__(self, i: int) -> dict:
        from univa.serve.cli import synthetic_chat_tokens
        g = torch.Generator().manual_seed(self.seed * 1_000_003 + i)
        # everything is derived from uint8 pixels with integer arithmetic, so a sample is bit-identical whatever the
        # host's thread count (a float resize rounds differently under OMP_NUM_THREADS=1, which torchrun sets)
        src_u8 = torch.randint(0, 256, (3, self.h, self.w), generator=g, dtype=torch.uint8)
        src = src_u8.float() / 127.5 - 1.0
        # the target is the source with a rectangular region re-drawn: a real edit changes part of the picture
        tgt = src.clone()
        bh, bw = max(16, self.h // 4), max(16, self.w // 4)
        y0 = int(torch.randint(0, self.h - bh + 1, (1,), generator=g))
        x0 = int(torch.randint(0, self.w - bw + 1, (1,), generator=g))
        tgt[:, y0:y0 + bh, x0:x0 + bw] = torch.randint(0, 256, (3, bh, bw), generator=g, dtype=torch.uint8).float() / 127.5 - 1.0
        lh, lw = self.h // 8, self.w // 8
        weights = torch.ones(1, lh, lw)
        ratio = (self.h * self.w) / (bh * bw)
        weights[:, y0 // 8:(y0 + bh) // 8, x0 // 8:(x0 + bw) // 8] = math.log2(ratio) + 1.0
        # 448x448 nearest-neighbour view of the source for the VLM: 32x32 patches of 14 -> 1024 rows of 1176, 256 image tokens
        iy = (torch.arange(448) * self.h) // 448
        ix = (torch.arange(448) * self.w) // 448
        u8 = src_u8[:, iy][:, :, ix].permute(1, 2, 0).contiguous().numpy()
        from gpt_image_edit_b200.image_io import qwen_pixel_values
        pix, grid = qwen_pixel_values(u8)
        ids = synthetic_chat_tokens(pix.shape[0] // 4, n_text=self.n_text, vocab=self.vocab, seed_=self.seed * 7919 + i)[0]
        return dict(generated_image=tgt, ref_pixel_values=src, input_ids=ids, attention_mask=torch.ones_like(ids),
                    pixel_values=pix, image_grid_thw=grid[0], weights=weights, prompts="synthetic instruction")


def collate(samples: list) -> dict:
    """Equal-size samples only (the synthetic set has one resolution; the reference's mixed-size list path needs
    attention masks, which this engine does not implement)."""
    out = {}
    for k in ("generated_image", "ref_pixel_values", "input_ids", "attention_mask", "weights"):
        out[k] = torch.stack([s[k] for s in samples])
    out["pixel_values"] = torch.cat([s["pixel_values"] for s in samples], dim=0)
    out["image_grid_thw"] = torch.stack([s["image_grid_thw"] for s in samples])
    out["prompts"] = [s["prompts"] for s in samples]
    return out
