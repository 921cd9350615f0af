"""Logit-domain form of the reference's decision rule ``sigmoid(x) > lower_bound``.

``apply_thresholds`` (/root/reference/src/segma/inference.py:228-234) compares the fp32 sigmoid with an
fp32 threshold, strictly.  ``sigmoid(x) > t`` is not ``x > logit(t)`` in floating point (e.g.
``sigmoid(x) > 0.5`` is false for 0 < x <~ 6e-8), so the cut is found by bisection over float32 bit
patterns against torch's own fp32 sigmoid: ``logit_cut(t)`` is the largest x with ``sigmoid(x) <= t``;
the kernel then tests ``x > cut``, which is bit-exact with the reference for every finite logit.
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


@lru_cache(maxsize=256)
def logit_cut(threshold: float) -> float:
    t = np.float32(threshold)
    if not (t >= 0.0):  # negative (or nan) thresholds: every finite logit is active
        return -math.inf
    lo, hi = _float_to_ordered(-3.0e38), _float_to_ordered(3.0e38)
    if _active(np.array([hi]), t)[0] == False:  # noqa: E712  (t >= 1: nothing is active)
        return math.inf
    if _active(np.array([lo]), t)[0]:
        return -math.inf
    while hi - lo > 1:  # invariant: lo inactive, hi active
        mid = (lo + hi) // 2
        if _active(np.array([mid]), t)[0]:
            hi = mid
        else:
            lo = mid
    # the sigmoid is monotone around the cut: check a neighbourhood so a non-monotone libm would be caught
    around = np.arange(lo - 64, lo + 65)
    act = _active(around, t)
    assert not act[:65].any() and act[65:].all(), "fp32 sigmoid is not monotone around the threshold"
    return float(_ordered_to_float(np.array([lo]))[0])
