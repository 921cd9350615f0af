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


_COPY_STREAMS: dict[int, "torch.cuda.Stream"] = {}


def _copy_stream_for(device: torch.device) -> "torch.cuda.Stream":
    """One H2D staging stream per device, shared by all files (not one per file)."""
    idx = device.index if device.index is not None else torch.cuda.current_device()
    if idx not in _COPY_STREAMS:
        _COPY_STREAMS[idx] = torch.cuda.Stream(device=device)
    return _COPY_STREAMS[idx]


class PcmSource:
    """A file's samples as a 1-D float32 device tensor that is filled progressively.

    ``ensure(n)`` makes samples ``[0, n)`` valid for kernels launched afterwards on the current stream; host
    reads, PCIe copies (on a side stream) and the widening kernel therefore overlap the compute of earlier
    batches instead of preceding the whole file.  Sources: a device tensor (nothing to do), a pinned or
    pageable host array, or a mono PCM16 / PCM32 / float32 WAV file read in its stored width."""

    def __init__(self, audio_p, device, dev_out: torch.Tensor | None = None):
        """``dev_out``: a 1-D float32 device tensor of exactly this file's length to stage into (a slice of a buffer that
        packs several files); by default the source allocates its own."""
        from . import ops

        self._ops = ops
        self.device = torch.device(device)
        self._done = 0
        self._file = None
        self._host = None
        arr = audio_p if isinstance(audio_p, torch.Tensor) else _as_array(audio_p)
        if isinstance(arr, torch.Tensor) and arr.is_cuda:
            self.dev = arr.reshape(-1).to(torch.float32).contiguous()
            self.n_samples = self._done = self.dev.numel()
            if dev_out is not None:
                dev_out.copy_(self.dev)
                self.dev = dev_out
            return
        if arr is not None:
            host = torch.as_tensor(arr).reshape(-1)
            if host.dtype != torch.float32:
                host = host.to(torch.float32)
            self._host = host if host.is_pinned() else host.contiguous().pin_memory()
            self.n_samples = host.numel()
            self._fmt = (ops.PCM_F32, np.dtype("<f4"), torch.float32)
        else:
            lay = _parse_wav(Path(audio_p))
            fmt = {(1, 16): (ops.PCM_S16, np.dtype("<i2"), torch.int16), (1, 32): (ops.PCM_S32, np.dtype("<i4"), torch.int32),
                   (3, 32): (ops.PCM_F32, np.dtype("<f4"), torch.float32)}.get((lay.fmt, lay.bits))
            if fmt is None or lay.n_channels != 1:  # anything else goes through the host decoder
                t = get_samples_in_range(audio_p, 0, -1)
                if t.shape[0] != 1:
                    raise ValueError(f"only mono audio is supported, got {t.shape[0]} channels")
                self._host = t.reshape(-1).contiguous().pin_memory()
                self.n_samples = self._host.numel()
                self._fmt = (ops.PCM_F32, np.dtype("<f4"), torch.float32)
            else:
                self._fmt = fmt
                self.n_samples = lay.n_frames
                self._file = open(audio_p, "rb")
                self._file.seek(lay.data_offset)
                per = max(1, _STAGE_CHUNK // fmt[1].itemsize)
                self._bufs = [torch.empty(min(per, max(self.n_samples, 1)), dtype=fmt[2]).pin_memory() for _ in range(2)]
                self._buf_events = [None, None]
                self._turn = 0
        code, _, t_dt = self._fmt
        with torch.cuda.device(self.device):
            if dev_out is not None:
                assert dev_out.numel() == self.n_samples and dev_out.dtype == torch.float32 and dev_out.is_contiguous()
            self.dev = dev_out if dev_out is not None else torch.empty(self.n_samples, dtype=torch.float32, device=self.device)
            self._raw = self.dev if code == ops.PCM_F32 else torch.empty(self.n_samples, dtype=t_dt, device=self.device)
            # The caching allocator may hand back blocks whose previous owner (an earlier file's PCM, logits or
            # scratch) still has kernels queued on the current stream: the copy stream must not write them earlier.
            self._copy_stream = _copy_stream_for(self.device)
            self._copy_stream.wait_stream(torch.cuda.current_stream(self.device))

    def skip_to(self, first_sample: int) -> None:
        """Samples before ``first_sample`` will not be read by anyone (a rank that runs only some batches of the
        file): do not stage them."""
        first_sample = min(int(first_sample), self.n_samples)
        if first_sample <= self._done:
            return
        if self._file is not None:
            self._file.seek((first_sample - self._done) * self._fmt[1].itemsize, 1)
        self._done = first_sample

    def ensure(self, upto: int) -> None:
        upto = min(int(upto), self.n_samples)
        if upto <= self._done:
            return
        ops = self._ops
        code, np_dt, _ = self._fmt
        main = torch.cuda.current_stream(self.device)
        a = self._done
        if self._host is not None:
            with torch.cuda.stream(self._copy_stream):
                self._raw[a:upto].copy_(self._host[a:upto], non_blocking=True)
        else:
            per = self._bufs[0].numel()
            pos = a
            while pos < upto:
                cnt = min(per, upto - pos)
                i = self._turn & 1
                if self._buf_events[i] is not None:
                    self._buf_events[i].synchronize()  # the previous copy out of this pinned buffer has finished
                view = self._bufs[i][:cnt].numpy()
                got = self._file.readinto(memoryview(view).cast("B"))
                if got != cnt * np_dt.itemsize:
                    raise ValueError("truncated WAV data chunk")
                with torch.cuda.stream(self._copy_stream):
                    self._raw[pos:pos + cnt].copy_(self._bufs[i][:cnt], non_blocking=True)
                    ev = torch.cuda.Event()
                    ev.record(self._copy_stream)
                self._buf_events[i] = ev
                self._turn += 1
                pos += cnt
        main.wait_stream(self._copy_stream)
        if code != ops.PCM_F32:
            with torch.cuda.device(self.device):
                ops.pcm_to_f32(self._raw[a:upto], code, out=self.dev[a:upto])
        self._done = upto
        if self._done >= self.n_samples and self._file is not None:
            self._file.close()
            self._file = None

    def all(self) -> torch.Tensor:
        self.ensure(self.n_samples)
        return self.dev


def audio_n_samples(audio_p) -> int:
    """Samples of a file or array without staging it."""
    if isinstance(audio_p, torch.Tensor):
        return int(audio_p.numel())
    if isinstance(audio_p, np.ndarray):
        return int(audio_p.size)
    return int(get_audio_info(audio_p).n_samples)


def stage_to_device(audio_p, device) -> torch.Tensor:
    """Whole file -> 1-D float32 tensor on ``device`` (see ``PcmSource``)."""
    return PcmSource(audio_p, device).all()
