"""Aspect-ratio bucketing behind `--anyres` (the reference's univa/utils/anyres_util.py: pick_ratio :22-30,
compute_size :33-57, dynamic_resize :60-78).  Host logic; results are pinned to the reference's own outputs on a
112-case grid (tests/golden/host_ref.pt).

The bucket families are nested: every smaller family drops one landscape/portrait pair from the 11-ratio list, and
the 17-ratio family is the reduced form of the Kontext training resolutions the pipeline already carries."""
from __future__ import annotations

from fractions import Fraction
from math import sqrt

from gpt_image_edit_b200.pipeline import PREFERRED_KONTEXT_RESOLUTIONS as RESOLUTIONS_17

_ELEVEN = ((16, 9), (7, 5), (5, 4), (4, 3), (3, 2))        # landscape members, each followed by its transpose


def _family(drop=()):
    out = []
    for w, h in _ELEVEN:
        if (w, h) not in drop:
            out += [(w, h), (h, w)]
    return out + [(1, 1)]


def _lowest_terms(w, h):
    f = Fraction(w, h)
    return f.numerator, f.denominator


RATIO = {
    "any_17ratio": [_lowest_terms(w, h) for w, h in RESOLUTIONS_17],
    "any_11ratio": _family(),
    "any_9ratio": _family(drop={(7, 5)}),
    "any_7ratio": _family(drop={(7, 5), (5, 4)}),
    "any_5ratio": _family(drop={(7, 5), (5, 4), (3, 2)}),
    "any_1ratio": [(1, 1)],
}


def pick_ratio(orig_h: int, orig_w: int, anyres: str = "any_17ratio"):
    """(rw, rh) of the bucket whose w/h is nearest to the image's; the earlier entry wins a tie."""
    aspect = orig_w / orig_h
    best, best_d = None, float("inf")
    for rw, rh in RATIO[anyres]:
        d = abs(rw / rh - aspect)
        if d < best_d:
            best, best_d = (rw, rh), d
    return best


def _snap(v: float, stride: int) -> int:
    return max(stride, int(v)) // stride * stride


def compute_size(rw: int, rh: int, stride: int, *, min_pixels=None, max_pixels=None, anchor_pixels=None):
    """(h, w): the `stride`-unit bucket rescaled (real factor) to the anchor area, or clamped into
    [min_pixels, max_pixels], then floored to the stride grid."""
    unit_area = (rw * stride) * (rh * stride)
    if anchor_pixels is not None:
        want = anchor_pixels
    elif min_pixels is None or max_pixels is None:
        want = unit_area
    else:
        want = min(max(unit_area, min_pixels), max_pixels)
    k = sqrt(want / unit_area)
    return _snap(rh * stride * k, stride), _snap(rw * stride * k, stride)


def dynamic_resize(orig_h: int, orig_w: int, anyres: str = "any_17ratio", anchor_pixels: int = 1024 * 1024,
                   stride: int = 32):
    """(h, w): the bucket in `stride` units times an INTEGER factor chosen so that the area lands near the anchor."""
    rw, rh = pick_ratio(orig_h, orig_w, anyres)
    factor = max(1, round(sqrt(anchor_pixels / (rw * rh * stride * stride))))
    return rh * stride * factor, rw * stride * factor
