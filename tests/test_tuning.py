"""Threshold tuning: the oracle's restatement of scripts/tune.py against sklearn on
the CPU, and the device histogram path against the oracle on the GPU."""
import math

import numpy as np
import pytest
import torch

from oracle import segma_oracle as O
from segma_b200 import tuning

LABELS = ["KCHI", "OCH", "MAL", "FEM"]


def _data(n, seed):
    g = torch.Generator().manual_seed(seed)
    truth = (torch.rand((n, 4), generator=g) < torch.tensor([0.3, 0.05, 0.5, 0.0])).float()
    logits = (truth * 2 - 1) * 1.5 + torch.randn((n, 4), generator=g) * 1.7
    return truth, logits


def test_grid_matches_reference_quirk():
    thr, n_steps = tuning.threshold_grid(0.1)
    assert n_steps == 10 and [round(float(t), 1) for t in thr] == [0.0, 0.1, 0.2, 0.3, 0.4, 0.6, 0.7, 0.8, 0.9, 1.0]
    thr, n_steps = tuning.threshold_grid(0.01)
    assert n_steps == 100 and len(thr) == 100


def test_oracle_f1_equals_sklearn():
    sklearn = pytest.importorskip("sklearn.metrics")
    truth, logits = _data(5000, 0)
    for t in (0.0, 0.3, 0.5, 0.9, 1.0):
        pred = logits.sigmoid() > t
        want = sklearn.f1_score(y_true=truth, y_pred=pred, average=None, labels=list(range(4)), zero_division=1.0)
        assert np.allclose(O.f1_per_label(truth, pred), want, rtol=0, atol=1e-12)


def test_f1_from_histogram_equals_oracle_on_cpu():
    """The suffix-sum reconstruction used by the product, fed with a CPU-built histogram."""
    from segma_b200.thresholds import logit_cut

    truth, logits = _data(20_000, 1)
    thr, _ = tuning.threshold_grid(0.1)
    cuts = torch.tensor([logit_cut(float(t)) for t in thr])
    b = (logits[:, :, None] > cuts[None, None, :]).sum(-1)  # cuts exceeded
    hist = torch.zeros((4, 2, len(thr) + 1), dtype=torch.int64)
    for c in range(4):
        for y in (0, 1):
            sel = truth[:, c] == y
            hist[c, y] = torch.bincount(b[sel, c], minlength=len(thr) + 1)
    f1 = tuning.f1_from_histogram(hist)
    for k, t in enumerate(thr):
        assert np.allclose(f1[k].numpy(), O.f1_per_label(truth, logits.sigmoid() > t), atol=1e-12)


def test_rttm_to_tensor(tmp_path):
    p = tmp_path / "a.rttm"
    p.write_text("SPEAKER a <NA> 0.02 0.04 <NA> <NA> KCHI <NA> <NA>\nSPEAKER a <NA> 1.0 0.5 <NA> <NA> FEM <NA> <NA>\n"
                 "SPEAKER a <NA> 0.0 9.0 <NA> <NA> OTHER <NA> <NA>\n")
    t = tuning.rttm_to_tensor(p, LABELS)
    assert t.shape == (75, 4)
    assert t[:, 0].nonzero().flatten().tolist() == [1, 2]
    assert t[:, 3].sum() == 25 and t[50, 3] == 1 and t[49, 3] == 0
    assert t[:, 1].sum() == 0 and t[:, 2].sum() == 0


@pytest.mark.gpu
@pytest.mark.parametrize("precision,n", [(0.1, 50_000), (0.01, 200_000)])
def test_device_tuning_equals_oracle(cuda, precision, n):
    truth, logits = _data(n, 2)
    thr, n_steps = tuning.threshold_grid(precision)
    got = tuning.tune_multilabel({"val": {"true": truth, "pred": logits}}, thr, LABELS, n_steps)
    want = O.tune_multilabel(truth, logits, thr, LABELS, n_steps)
    assert got == want
    # the per-threshold F1 table itself
    from segma_b200 import ops
    from segma_b200.thresholds import logit_cut

    cuts = sorted(logit_cut(float(t)) for t in thr)
    hist = ops.threshold_histogram(logits.cuda().contiguous(), (truth != 0).to(torch.uint8).cuda().contiguous(), cuts)
    assert int(hist.sum()) == n * 4
    f1 = tuning.f1_from_histogram(hist)
    for k, t in enumerate(sorted(float(t) for t in thr)):
        assert np.allclose(f1[k].numpy(), O.f1_per_label(truth, logits.sigmoid() > torch.tensor(t)), atol=1e-12)


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["p10", "p100", "p10_rare"])
def test_device_tuning_equals_reference_golden(cuda, name):
    """The histogram path against thresholds the reference's own ``tune_multilabel`` chose (tests/golden/tuning.npz,
    generated from /root/reference/scripts/tune.py by oracle/make_golden.py::tuning)."""
    from pathlib import Path

    g = np.load(Path(__file__).parent / "golden" / "tuning.npz")
    n, seed, n_steps = (int(v) for v in g[f"{name}_meta"])
    gen = torch.Generator().manual_seed(seed)
    truth = (torch.rand((n, 4), generator=gen) < torch.tensor([0.3, 0.05, 0.5, 0.0])).float()
    logits = (truth * 2 - 1) * 1.5 + torch.randn((n, 4), generator=gen) * 1.7
    thr = torch.from_numpy(g[f"{name}_thresholds"])
    best = tuning.tune_multilabel({"val": {"true": truth, "pred": logits}}, thr, LABELS, n_steps)
    assert [best[lab]["lower_bound"] for lab in LABELS] == g[f"{name}_best"].tolist()
