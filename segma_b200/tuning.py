"""Threshold tuning on the device: the grid search of /root/reference/scripts/tune.py (``rttm_to_tensor`` 15-56,
``tune_multilabel`` 213-256) with the per-threshold sklearn F1 passes replaced by one histogram pass over
the logits (``segma_threshold_histogram``); SURVEY.md section 8f, row f4.  Same inputs and outputs:
``{label: {"lower_bound": best, "upper_bound": 1.0}}``.
"""
from __future__ import annotations

import math
from pathlib import Path

import torch

from . import ops
from .thresholds import logit_cut


def threshold_grid(precision: float = 0.1) -> tuple[torch.Tensor, int]:
    """The reference's grid: ``linspace(0, 1, n_steps).round(decimals=log10(n_steps))`` (tune.py:288-291)."""
    n_steps = int(1 / precision)
    return torch.linspace(0, 1, steps=n_steps).round(decimals=int(math.log10(n_steps))), n_steps


def rttm_to_tensor(rttm_path: Path, labels: list[str], frame_resolution_s: float = 0.02) -> torch.Tensor:
    """RTTM file -> (num_frames, num_labels) multi-hot float32 at 20 ms (tune.py:15-56)."""
    wanted = {lab: i for i, lab in enumerate(labels)}
    segs = []
    with open(rttm_path, "r") as f:
        for line in f:
            parts = line.strip().split()
            if len(parts) > 7 and parts[7] in wanted:
                segs.append((float(parts[3]), float(parts[4]), wanted[parts[7]]))
    total = max((s + d for s, d, _ in segs), default=0)
    n = math.ceil(total / frame_resolution_s)
    out = torch.zeros(n, len(labels), dtype=torch.float32)
    for s, d, c in segs:
        out[int(s / frame_resolution_s): min(math.ceil((s + d) / frame_resolution_s), n), c] = 1.0
    return out


def f1_from_histogram(hist: torch.Tensor) -> torch.Tensor:
    """(C, 2, K+1) counts of "cuts exceeded" -> (K, C) F1 per grid point and label, sklearn semantics
    (``average=None, zero_division=1.0``): prediction at grid point k is positive iff more than k cuts are exceeded."""
    h = hist.to(torch.float64).cpu()
    neg, pos = h[:, 0], h[:, 1]  # (C, K+1)
    suffix = lambda t: t.flip(-1).cumsum(-1).flip(-1)  # noqa: E731  sum over bins >= b
    tp = suffix(pos)[:, 1:]  # bins > k, k = 0..K-1
    fp = suffix(neg)[:, 1:]
    fn = pos.sum(-1, keepdim=True) - tp
    denom = 2 * tp + fp + fn
    f1 = torch.where(denom > 0, 2 * tp / denom.clamp(min=1), torch.ones_like(denom))
    return f1.T.contiguous()


def tune_multilabel(data_t: dict, thresholds, labels: list[str], n_steps: int | None = None) -> dict:
    """Grid search of the per-label ``lower_bound`` by F1 on ``data_t["val"]`` (``true``: (n, C) 0/1, ``pred``:
    (n, C) raw logits).  Ties keep the first (smallest) threshold, like ``max(dict, key=dict.get)`` over the
    reference's insertion order."""
    thr = [float(t) for t in thresholds]
    n_steps = len(thr) if n_steps is None else n_steps
    dev = torch.device("cuda")
    logits = data_t["val"]["pred"].to(dev, torch.float32).contiguous()
    truth = (data_t["val"]["true"] != 0).to(dev, torch.uint8).contiguous()
    order = sorted(range(len(thr)), key=lambda i: thr[i])
    cuts = [logit_cut(float(torch.tensor(thr[i], dtype=torch.float32))) for i in order]
    f1_sorted = f1_from_histogram(ops.threshold_histogram(logits, truth, cuts))  # rows follow `order`
    f1 = torch.empty_like(f1_sorted)
    f1[order] = f1_sorted
    best = {}
    digits = int(math.log10(n_steps))
    for c, lab in enumerate(labels):
        scores: dict[float, float] = {}
        for k, t in enumerate(thr):  # a repeated grid value keeps its first position and its last score, as a dict does
            scores[t] = float(f1[k, c])
        top = max(scores, key=scores.get)
        best[lab] = {"lower_bound": round(float(top), digits), "upper_bound": 1.0}
    return best
