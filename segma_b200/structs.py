"""``Intervals``: the reference's multi-label interval set (/root/reference/src/segma/structs/interval.py:8-54)
with the per-label merge done by the interval post-processing kernel.

Same surface -- ``add``, ``intervals``, iteration, ``len``, ``repr`` -- and the same result: intervals of one label
that overlap or touch (``s <= previous end``, interval.py:26-31) become one, labels never mix, and the list is kept
``sorted()`` over all labels (interval.py:45).  Insertion order is free; nested and repeated intervals are absorbed.
``extend`` merges a whole batch in one kernel call, ``from_table`` / ``merged_table`` work on the int32
``(file, label, start, end)`` tables of ``ops.decode_intervals`` without leaving the device.

Host side: label -> index mapping and the (label, start, end) ordering of the rows (integer bookkeeping); the
merge itself (running maximum of the ends per label, group boundaries, compaction) is ``segma_postprocess_intervals``.
Coordinates are int32 on the device, as in the interval tables.
"""
from __future__ import annotations

from typing import Iterable, Iterator

import numpy as np
import torch

from . import ops

Interval = tuple  # (start, end, label)
_I32 = np.iinfo(np.int32)


class Intervals:
    def __init__(self, max_gap: int = 0) -> None:
        self.intervals: list[Interval] = []
        self.max_gap = int(max_gap)

    def add(self, interval: Interval) -> None:
        """Add one interval and re-reduce (interval.py:15-17)."""
        self.extend([interval])

    def extend(self, intervals: Iterable[Interval]) -> None:
        self.intervals = self._reduce_per_label(list(self.intervals) + [tuple(iv) for iv in intervals])

    def _reduce_per_label(self, intervals: list[Interval]) -> list[Interval]:
        if not intervals:
            return []
        labels: dict = {}
        for _, _, lab in intervals:
            labels.setdefault(lab, len(labels))  # first-appearance order, like the reference's defaultdict
        rows = np.array([(0, labels[lab], s, e) for s, e, lab in intervals], dtype=np.int64)
        if rows[:, 2:].min() < _I32.min or rows[:, 2:].max() > _I32.max:
            raise OverflowError("interval coordinates must fit int32 (the device table type)")
        order = np.lexsort((rows[:, 3], rows[:, 2], rows[:, 1]))  # by label, then start, then end
        table = torch.from_numpy(rows[order].astype(np.int32)).cuda()
        merged = ops.postprocess_intervals(table, self.max_gap, 0).cpu().numpy()
        names = list(labels)
        return sorted((int(s), int(e), names[int(c)]) for _, c, s, e in merged)

    @staticmethod
    def merged_table(table: torch.Tensor, max_gap_samples: int = 0, min_duration_samples: int = 0) -> torch.Tensor:
        """Device table in decode order (file-, label-major, time-ordered) -> merged / filtered device table."""
        return ops.postprocess_intervals(table.contiguous(), max_gap_samples, min_duration_samples)

    def __repr__(self):
        return "%s(%r)" % (self.__class__.__name__, self.intervals)

    def __iter__(self) -> Iterator[Interval]:
        return iter(self.intervals)

    def __len__(self):
        return len(self.intervals)
