"""Label <-> column index mapping for the multi-label heads.

Same contract as the reference's ``MultiLabelEncoder``
(/root/reference/src/segma/utils/encoders.py:51-119): ``base_labels`` fixes the column
order of the logits and the label-major order of the decoded intervals
(inference.py:258-262), ``_labels`` drives the default thresholds (inference.py:313).
"""
from __future__ import annotations

from collections.abc import Iterable

import numpy as np


class MultiLabelEncoder:
    def __init__(self, labels: list[str] | tuple[str, ...]) -> None:
        self._labels = labels
        self.n_labels = len(labels)
        self.map = {name: idx for idx, name in enumerate(labels)}
        self.rev_map = {idx: name for name, idx in self.map.items()}

    @property
    def labels(self) -> tuple[str, ...]:
        return tuple(self.map)

    @property
    def base_labels(self) -> tuple[str, ...]:
        return tuple(self._labels)

    def transform(self, label) -> int:
        return self.map[label]

    def __call__(self, labels=()) -> int:
        return self.transform(labels)

    def inv_transform(self, i: int) -> str:
        if i < 0 or i >= self.n_labels:
            raise ValueError(
                f"transformed index '{i}' is not assigned, only {self.n_labels} labels are available."
            )
        return self.rev_map[i]

    def one_hot(self, labels: Iterable[str] | str) -> np.ndarray:
        names = (labels,) if isinstance(labels, str) else labels
        out = np.zeros(self.n_labels, dtype=int)
        out[[self.transform(n) for n in names]] = 1
        return out

    def i_to_one_hot(self, i: int) -> np.ndarray:
        return self.one_hot(self.rev_map[i])

    def __len__(self) -> int:
        return self.n_labels

    def __contains__(self, label) -> bool:
        if isinstance(label, (list, tuple)):
            raise ValueError("Collections not supported, only single item membership makes sense")
        return label in self.map
