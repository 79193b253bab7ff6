"""Aspect-ratio bucketing used by the CLI to pick the generation size (reference
univa/utils/anyres_util.py:22-78; host logic, results must be identical)."""
from __future__ import annotations

import math

RESOLUTIONS_17 = [(672, 1568), (688, 1504), (720, 1456), (752, 1392), (800, 1328), (832, 1248), (880, 1184),
                  (944, 1104), (1024, 1024), (1104, 944), (1184, 880), (1248, 832), (1328, 800), (1392, 752),
                  (1456, 720), (1504, 688), (1568, 672)]


def _reduced(pairs):
    return [(w // math.gcd(w, h), h // math.gcd(w, h)) for w, h in pairs]


RATIO = {
    "any_17ratio": _reduced(RESOLUTIONS_17),
    "any_11ratio": [(16, 9), (9, 16), (7, 5), (5, 7), (5, 4), (4, 5), (4, 3), (3, 4), (3, 2), (2, 3), (1, 1)],
    "any_9ratio": [(16, 9), (9, 16), (5, 4), (4, 5), (4, 3), (3, 4), (3, 2), (2, 3), (1, 1)],
    "any_7ratio": [(16, 9), (9, 16), (4, 3), (3, 4), (3, 2), (2, 3), (1, 1)],
    "any_5ratio": [(16, 9), (9, 16), (4, 3), (3, 4), (1, 1)],
    "any_1ratio": [(1, 1)],
}


def pick_ratio(orig_h: int, orig_w: int, anyres: str = "any_17ratio"):
    """Closest (w, h) ratio of the bucket list to the image's aspect ratio; first wins on ties."""
    target = orig_w / orig_h
    return min(RATIO[anyres], key=lambda wh: abs(wh[0] / wh[1] - target))


def compute_size(rw: int, rh: int, stride: int, *, min_pixels=None, max_pixels=None, anchor_pixels=None):
    bw, bh = rw * stride, rh * stride
    area = bw * bh
    if anchor_pixels is not None:
        goal = anchor_pixels
    elif min_pixels is not None and max_pixels is not None:
        goal = max_pixels if area > max_pixels else min_pixels if area < min_pixels else area
    else:
        goal = area
    k = math.sqrt(goal / area)
    new_w = max(stride, int(bw * k)) // stride * stride
    new_h = max(stride, int(bh * k)) // stride * stride
    return new_h, new_w


def dynamic_resize(orig_h: int, orig_w: int, anyres: str = "any_17ratio", anchor_pixels: int = 1024 * 1024,
                   stride: int = 32):
    """(h, w): the bucket ratio scaled by an INTEGER factor so the area lands near `anchor_pixels`."""
    rw, rh = pick_ratio(orig_h, orig_w, anyres)
    bw, bh = rw * stride, rh * stride
    s = max(1, round(math.sqrt(anchor_pixels / (bw * bh))))
    return (bh * s) // stride * stride, (bw * s) // stride * stride
