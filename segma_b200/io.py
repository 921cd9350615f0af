"""Audio access for the inference driver: header info and sample ranges of 16 kHz mono WAV files.

Stands in for the reference's torchcodec wrappers (/root/reference/src/segma/utils/io.py:18-47) with the
same names and shapes (``(n_channels, n_samples)`` float32 in [-1, 1]); RIFF/WAVE PCM16, PCM24, PCM32 and
float32 are parsed directly (no FFmpeg).  In-memory audio (numpy / torch 1-D arrays) is accepted
wherever the driver takes a path.  Decode + host->device staging is row f1 of SURVEY.md section 8f.
"""
from __future__ import annotations

import struct
from dataclasses import dataclass
from pathlib import Path

import numpy as np
import torch


@dataclass
class AudioInfo:
    sample_rate: int
    n_samples: int
    n_channels: int


@dataclass
class _WavLayout:
    sample_rate: int
    n_channels: int
    fmt: int  # 1 = PCM, 3 = IEEE float
    bits: int
    data_offset: int
    n_frames: int


def _parse_wav(path: Path) -> _WavLayout:
    with open(path, "rb") as f:
        head = f.read(12)
        if len(head) < 12 or head[:4] != b"RIFF" or head[8:12] != b"WAVE":
            raise ValueError(f"{path} is not a RIFF/WAVE file")
        fmt = None
        while True:
            ck = f.read(8)
            if len(ck) < 8:
                raise ValueError(f"{path}: no data chunk")
            cid, size = ck[:4], struct.unpack("<I", ck[4:])[0]
            if cid == b"fmt ":
                body = f.read(size + (size & 1))
                tag, ch, sr, _, _, bits = struct.unpack("<HHIIHH", body[:16])
                if tag == 0xFFFE and size >= 26:  # WAVE_FORMAT_EXTENSIBLE: sub-format GUID starts with the tag
                    tag = struct.unpack("<H", body[24:26])[0]
                fmt = (tag, ch, sr, bits)
            elif cid == b"data":
                if fmt is None:
                    raise ValueError(f"{path}: data chunk before fmt chunk")
                tag, ch, sr, bits = fmt
                offset = f.tell()
                avail = Path(path).stat().st_size - offset
                size = min(size, avail) if size not in (0, 0xFFFFFFFF) else avail
                return _WavLayout(sr, ch, tag, bits, offset, size // (ch * bits // 8))
            else:
                f.seek(size + (size & 1), 1)


def _read_frames(path: Path, lay: _WavLayout, start: int, count: int) -> np.ndarray:
    """-> float32 (n_channels, count)"""
    start = max(0, min(start, lay.n_frames))
    count = max(0, min(count, lay.n_frames - start))
    bps = lay.bits // 8
    with open(path, "rb") as f:
        f.seek(lay.data_offset + start * lay.n_channels * bps)
        raw = f.read(count * lay.n_channels * bps)
    if lay.fmt == 3 and lay.bits == 32:
        x = np.frombuffer(raw, dtype="<f4").astype(np.float32)
    elif lay.fmt == 3 and lay.bits == 64:
        x = np.frombuffer(raw, dtype="<f8").astype(np.float32)
    elif lay.fmt == 1 and lay.bits == 16:
        x = np.frombuffer(raw, dtype="<i2").astype(np.float32) / 32768.0
    elif lay.fmt == 1 and lay.bits == 32:
        x = np.frombuffer(raw, dtype="<i4").astype(np.float32) / 2147483648.0
    elif lay.fmt == 1 and lay.bits == 24:
        b = np.frombuffer(raw, dtype=np.uint8).reshape(-1, 3).astype(np.int32)
        v = b[:, 0] | (b[:, 1] << 8) | (b[:, 2] << 16)
        v = np.where(v >= 1 << 23, v - (1 << 24), v)
        x = v.astype(np.float32) / 8388608.0
    elif lay.fmt == 1 and lay.bits == 8:
        x = (np.frombuffer(raw, dtype=np.uint8).astype(np.float32) - 128.0) / 128.0
    else:
        raise ValueError(f"{path}: unsupported WAV encoding (format tag {lay.fmt}, {lay.bits} bits)")
    return np.ascontiguousarray(x.reshape(-1, lay.n_channels).T)


def _as_array(audio) -> np.ndarray | None:
    if isinstance(audio, torch.Tensor):
        return audio.detach().cpu().numpy()
    if isinstance(audio, np.ndarray):
        return audio
    return None


def get_audio_info(audio_p) -> AudioInfo:
    arr = _as_array(audio_p)
    if arr is not None:
        return AudioInfo(sample_rate=16_000, n_samples=int(arr.shape[-1]), n_channels=1 if arr.ndim == 1 else arr.shape[0])
    lay = _parse_wav(Path(audio_p))
    return AudioInfo(sample_rate=lay.sample_rate, n_samples=lay.n_frames, n_channels=lay.n_channels)


def get_samples_in_range(audio_p, start_f: int, duration_f: int) -> torch.Tensor:
    """Samples ``[start_f, start_f + duration_f)`` (to the end if ``duration_f < 0``) as (n_channels, n) float32."""
    arr = _as_array(audio_p)
    if arr is not None:
        a = arr.reshape(1, -1) if arr.ndim == 1 else arr
        end = a.shape[-1] if duration_f < 0 else start_f + duration_f
        return torch.from_numpy(np.ascontiguousarray(a[:, start_f:end], dtype=np.float32))
    p = Path(audio_p)
    lay = _parse_wav(p)
    count = lay.n_frames - start_f if duration_f < 0 else duration_f
    return torch.from_numpy(_read_frames(p, lay, start_f, count))


def get_all_samples(audio_p) -> torch.Tensor:
    return get_samples_in_range(audio_p, 0, -1)


def write_wav(path, pcm: np.ndarray, sample_rate: int = 16_000, subtype: str = "float32") -> None:
    """Minimal mono WAV writer (tests and synthetic data)."""
    pcm = np.asarray(pcm).reshape(-1)
    if subtype == "float32":
        data, tag, bits = pcm.astype("<f4").tobytes(), 3, 32
    elif subtype == "int16":
        data, tag, bits = np.clip(np.round(pcm * 32768.0), -32768, 32767).astype("<i2").tobytes(), 1, 16
    else:
        raise ValueError(subtype)
    hdr = struct.pack("<4sI4s4sIHHIIHH4sI", b"RIFF", 36 + len(data), b"WAVE", b"fmt ", 16, tag, 1, sample_rate,
                      sample_rate * bits // 8, bits // 8, bits, b"data", len(data))
    with open(path, "wb") as f:
        f.write(hdr)
        f.write(data)


# ---- device staging ------------------------------------------------------------------
_STAGE_CHUNK = 32 << 20  # bytes per pinned staging buffer


def stage_to_device(audio_p, device) -> torch.Tensor:
    """Whole file -> 1-D float32 tensor on ``device``.  Mono PCM16 / PCM32 / float32 WAV data is read straight
    from the file into two alternating pinned buffers, copied in its stored width and widened by
    ``segma_pcm_to_f32`` on the device; anything else goes through the host decoder."""
    from . import ops

    arr = _as_array(audio_p)
    if arr is None:
        lay = _parse_wav(Path(audio_p))
        fmt = {(1, 16): (ops.PCM_S16, np.dtype("<i2"), torch.int16), (1, 32): (ops.PCM_S32, np.dtype("<i4"), torch.int32),
               (3, 32): (ops.PCM_F32, np.dtype("<f4"), torch.float32)}.get((lay.fmt, lay.bits))
        if fmt is not None and lay.n_channels == 1:
            code, np_dt, t_dt = fmt
            n = lay.n_frames
            raw = torch.empty(n, dtype=t_dt, device=device)
            per = max(1, _STAGE_CHUNK // np_dt.itemsize)
            bufs = [torch.empty(min(per, max(n, 1)), dtype=t_dt).pin_memory() for _ in range(2)]
            events = [None, None]
            with open(audio_p, "rb") as f:
                f.seek(lay.data_offset)
                done, i = 0, 0
                while done < n:
                    cnt = min(per, n - done)
                    if events[i & 1] is not None:
                        events[i & 1].synchronize()  # the previous copy out of this buffer has finished
                    view = bufs[i & 1][:cnt].numpy()
                    got = f.readinto(memoryview(view).cast("B"))
                    if got != cnt * np_dt.itemsize:
                        raise ValueError(f"{audio_p}: truncated data chunk")
                    raw[done:done + cnt].copy_(bufs[i & 1][:cnt], non_blocking=True)
                    ev = torch.cuda.Event()
                    ev.record()
                    events[i & 1] = ev
                    done += cnt
                    i += 1
            return ops.pcm_to_f32(raw, code) if code != ops.PCM_F32 else raw
    t = get_samples_in_range(audio_p, 0, -1)
    if t.shape[0] != 1:
        raise ValueError(f"only mono audio is supported, got {t.shape[0]} channels")
    return t.reshape(-1).pin_memory().to(device, non_blocking=True)
