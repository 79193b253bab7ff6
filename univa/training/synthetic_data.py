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
                 vocab: int = 152064, target_sizes=None):
        self.h, self.w, self.length, self.seed, self.n_text, self.vocab = height, width, length, seed, n_text, vocab
        # mixed-size batches (train_denoiser.py:907-916): sample i's TARGET has size target_sizes[i % len]; the source
        # image keeps (height, width), as the reference stacks condition_pixel_values
        self.target_sizes = [tuple(int(v) for v in s) for s in target_sizes] if target_sizes else None

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
        th, tw = self.target_sizes[i % len(self.target_sizes)] if self.target_sizes else (self.h, self.w)
        if (th, tw) == (self.h, self.w):
            tgt = src.clone()
        else:                                   # integer nearest-neighbour view of the source at the target's size
            ty = (torch.arange(th) * self.h) // th
            tx = (torch.arange(tw) * self.w) // tw
            tgt = src_u8[:, ty][:, :, tx].float() / 127.5 - 1.0
        bh, bw = max(16, th // 4), max(16, tw // 4)
        y0 = int(torch.randint(0, th - bh + 1, (1,), generator=g))
        x0 = int(torch.randint(0, tw - bw + 1, (1,), generator=g))
        tgt[:, y0:y0 + bh, x0:x0 + bw] = torch.randint(0, 256, (3, bh, bw), generator=g, dtype=torch.uint8).float() / 127.5 - 1.0
        lh, lw = th // 8, tw // 8
        weights = torch.ones(1, lh, lw)
        ratio = (th * tw) / (bh * bw)
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


PAD_TOKEN_ID = 151643          # Qwen2.5-VL's pad token (<|endoftext|>)


def collate(samples: list, pad_token_id: int = PAD_TOKEN_ID, padding_side: str = "right") -> dict:
    """The reference's DataCollator (univa/dataset/data_collator.py:76-156) for these samples: prompts of different
    lengths are padded with the pad token on `padding_side` and `attention_mask = input_ids.ne(pad_token_id)` (:113-121);
    targets of one size are stacked, targets of different sizes stay a LIST of [1, 3, H_i, W_i] tensors (and their
    area-mask weights a list of [1, 1, h_i, w_i]), which is what the reference's loop tests for (train_denoiser.py:907,
    :1120)."""
    out = {"ref_pixel_values": torch.stack([s["ref_pixel_values"] for s in samples])}
    out["input_ids"] = torch.nn.utils.rnn.pad_sequence([s["input_ids"] for s in samples], batch_first=True,
                                                       padding_value=pad_token_id, padding_side=padding_side)
    out["attention_mask"] = out["input_ids"].ne(pad_token_id).long()
    mixed = len({tuple(s["generated_image"].shape) for s in samples}) > 1
    for k in ("generated_image", "weights"):
        out[k] = [s[k][None] for s in samples] if mixed else torch.stack([s[k] for s in samples])
    out["pixel_values"] = torch.cat([s["pixel_values"] for s in samples], dim=0)
    out["image_grid_thw"] = torch.stack([s["image_grid_thw"] for s in samples])
    out["prompts"] = [s["prompts"] for s in samples]
    return out
