"""Logit-domain form of the reference's decision rule ``sigmoid(x) > lower_bound``.

``apply_thresholds`` (/root/reference/src/segma/inference.py:228-234) compares the fp32 sigmoid with an
fp32 threshold, strictly.  ``sigmoid(x) > t`` is not ``x > logit(t)`` in floating point (e.g.
``sigmoid(x) > 0.5`` is false for 0 < x <~ 6e-8), so the cut is found by bisection over float32 bit
patterns against torch's own fp32 sigmoid: ``logit_cut(t)`` is the largest x with ``sigmoid(x) <= t``;
the kernel then tests ``x > cut``, which is bit-exact with the reference for every finite logit wherever
torch's sigmoid is monotone around the cut (it is for the default 0.5 and most grid values).

torch's vectorised CPU sigmoid is *not* monotone at the ulp level everywhere (its exp wiggles by an ulp, e.g.
around t = 0.4 or 0.6), which makes ``sigmoid(x) > t`` itself implementation-defined within a couple of ulps
of the cut.  There the cut falls back to the correctly rounded definition -- float64 sigmoid rounded to
float32, which is monotone -- and agrees with any faithful fp32 evaluation except on those few bit patterns.
"""
from __future__ import annotations

import math
from functools import lru_cache

import numpy as np
import torch


def _ordered_to_float(k: np.ndarray) -> np.ndarray:
    """int64 keys in monotone order -> float32 values (inverse of the usual sign-flip trick)."""
    k = k.astype(np.int64)
    bits = np.where(k >= 0, k, (-(k + 1)) | 0x80000000).astype(np.uint32)
    return bits.view(np.float32)


def _float_to_ordered(x: float) -> int:
    b = int(np.float32(x).view(np.uint32))
    return b if b < 0x80000000 else -(b & 0x7FFFFFFF) - 1


def _active(keys: np.ndarray, t: np.float32) -> np.ndarray:
    x = torch.from_numpy(_ordered_to_float(keys).copy())
    return (x.sigmoid() > torch.tensor(t)).numpy()


def _active_exact(keys: np.ndarray, t: np.float32) -> np.ndarray:
    """Correctly rounded rule: float64 sigmoid, rounded to float32, strict compare."""
    x = _ordered_to_float(keys).astype(np.float64)
    with np.errstate(over="ignore"):
        s = (1.0 / (1.0 + np.exp(-x))).astype(np.float32)
    return s > t


def _bisect(active, t: np.float32) -> int | float:
    lo, hi = _float_to_ordered(-3.0e38), _float_to_ordered(3.0e38)
    if not active(np.array([hi]), t)[0]:
        return math.inf
    if active(np.array([lo]), t)[0]:
        return -math.inf
    while hi - lo > 1:  # invariant: lo inactive, hi active
        mid = (lo + hi) // 2
        if active(np.array([mid]), t)[0]:
            hi = mid
        else:
            lo = mid
    return lo


@lru_cache(maxsize=256)
def logit_cut(threshold: float) -> float:
    t = np.float32(threshold)
    if np.isnan(t):  # ``sigmoid(x) > nan`` is false for every frame
        return math.inf
    if t < 0.0:  # negative thresholds: every finite logit is active
        return -math.inf
    lo = _bisect(_active, t)
    if isinstance(lo, float):
        return lo
    around = np.arange(lo - 64, lo + 65)
    act = _active(around, t)
    if act[:65].any() or not act[65:].all():
        # torch's fp32 sigmoid is not monotone around this threshold: use the correctly rounded rule
        lo = _bisect(_active_exact, t)
        if isinstance(lo, float):
            return lo
    return float(_ordered_to_float(np.array([lo]))[0])


def cut_is_torch_exact(threshold: float) -> bool:
    """True if ``x > logit_cut(threshold)`` reproduces torch's fp32 ``sigmoid(x) > threshold`` on every bit
    pattern around the cut (i.e. torch's sigmoid is monotone there)."""
    t = np.float32(threshold)
    cut = logit_cut(threshold)
    if math.isinf(cut):
        return True
    k = _float_to_ordered(cut)
    around = np.arange(k - 64, k + 65)
    act = _active(around, t)
    return not act[:65].any() and bool(act[65:].all())
