"""RTTM output format of the path.

``AudioAnnotation.to_rttm`` reproduces /root/reference/src/segma/annotation.py:86-104 byte for
byte (``round(x, 8)`` of float64 seconds, ``SPEAKER <uri> <NA> <start> <dur> <NA> <NA> <label>
<NA> <NA>``); ``frames_to_seconds`` is conversions.py:32-35.
"""
from __future__ import annotations

from dataclasses import dataclass

SAMPLE_RATE = 16_000
PRECISION = 8


def frames_to_seconds(f, sample_rate: int = SAMPLE_RATE):
    return f / sample_rate


@dataclass
class AudioAnnotation:
    uid: str
    start_time_s: float
    duration_s: float
    label: str
    PRECISION: int = PRECISION

    @property
    def end_time_s(self) -> float:
        return self.start_time_s + self.duration_s

    def to_rttm(self) -> str:
        return (
            f"SPEAKER {self.uid} <NA> {round(self.start_time_s, self.PRECISION)} "
            f"{round(self.duration_s, self.PRECISION)} <NA> <NA> {self.label} <NA> <NA>"
        )

    @classmethod
    def from_rttm(cls, line: str) -> "AudioAnnotation":
        parts = line.strip().split(" ")
        assert len(parts) in (9, 10)
        return cls(uid=parts[1], start_time_s=float(parts[3]), duration_s=float(parts[4]), label=parts[7])


def rttm_line(uri: str, start_sample: int, end_sample: int, label: str) -> str:
    """One RTTM line for the sample interval ``[start_sample, end_sample)`` (inference.py:275-283)."""
    # same text as AudioAnnotation(...).to_rttm(), without building the dataclass per interval
    return (f"SPEAKER {uri} <NA> {round(start_sample / SAMPLE_RATE, PRECISION)} "
            f"{round((end_sample - start_sample) / SAMPLE_RATE, PRECISION)} <NA> <NA> {label} <NA> <NA>")
