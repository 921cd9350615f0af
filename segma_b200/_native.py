"""ctypes binding of ``libsegma_b200.so`` (the C ABI declared in include/segma_b200.h).

There is no CPU or PyTorch fallback: if the library cannot be loaded (or built in-tree with nvcc)
``load()`` raises, and every op raises ``SegmaNativeError`` on a non-zero status.
"""
from __future__ import annotations

import ctypes as C
from pathlib import Path

_LIB = None


class SegmaNativeError(RuntimeError):
    pass


class GemmArgs(C.Structure):
    _fields_ = [
        ("a", C.c_void_p),
        ("a_batch_stride", C.c_int64),
        ("a_row_stride", C.c_int64),
        ("batch", C.c_int),
        ("rows_per_batch", C.c_int),
        ("a_rows_per_batch", C.c_int),
        ("k", C.c_int),
        ("conv_taps", C.c_int),
        ("conv_stride", C.c_int),
        ("w", C.c_void_p),
        ("n", C.c_int),
        ("bias", C.c_void_p),
        ("add_src", C.c_void_p),
        ("add_batch_rows", C.c_int64),
        ("out", C.c_void_p),
        ("out_batch_rows", C.c_int64),
        ("out_row_offset", C.c_int64),
        ("ldo", C.c_int64),
        ("flags", C.c_int),
        ("a_col_per_ntile", C.c_int),
        ("a_cols", C.c_int),
        ("force_bn", C.c_int),
    ]


_vp, _i, _i64, _f, _sz = C.c_void_p, C.c_int, C.c_int64, C.c_float, C.c_size_t

#: name -> (restype, argtypes); mirrors include/segma_b200.h one to one
SIGNATURES = {
    "segma_last_error": (C.c_char_p, []),
    "segma_version": (_i, []),
    "segma_device_check": (_i, []),
    "segma_sm_count": (_i, []),
    "segma_pcm_to_f32": (_i, [_vp, _i, _i64, _vp, _vp]),
    "segma_logmel_scratch_bytes": (_sz, [_i, _i]),
    "segma_logmel": (_i, [_vp, _i64, _i, _i, _i64, _vp, _vp, _vp, _vp]),
    "segma_logmel_set_filters": (_i, [_vp]),
    "segma_logmel_get_filters": (_i, [_vp]),
    "segma_w2v2_layer0": (_i, [_vp, _i64, _i, _i, _i64, _vp, _vp, _vp, _i, _vp, _vp, _i, _vp]),
    "segma_w2v2_layer0_at": (_i, [_vp, _i64, _i, _i, _vp, _vp, _vp, _vp, _i, _vp, _vp, _i, _vp]),
    "segma_wavlm_gate": (_i, [_vp, _i64, _i, _i, _vp, _vp, _vp, _vp, _vp]),
    "segma_gemm_f16": (_i, [C.POINTER(GemmArgs), _vp]),
    "segma_layernorm": (_i, [_vp, _vp, _vp, _i64, _i, _vp, _vp, _vp, _i, _i, _f, _f, _i, _i, _vp]),
    "segma_attention": (_i, [_vp, _i, _i, _i, _i, _vp, _vp, _i, _vp, _vp]),
    "segma_attention_rel": (_i, [_vp, _i, _i, _i, _i, _vp, _vp, _vp, _vp]),
    "segma_cast_f16": (_i, [_vp, _i64, _vp, _i64, _i64, _i, _vp]),
    "segma_cast_f16_split": (_i, [_vp, _i64, _vp, _i64, _i64, _i, _vp]),
    "segma_lstm_layer": (_i, [_vp, _vp, _i, _i, _i, _i, _vp, _vp, _vp]),
    "segma_heads": (_i, [_vp, _i, _i, _i, _i, _vp, _vp, _i, _vp, _i64, _i, _vp]),
    "segma_heads_at": (_i, [_vp, _i, _i, _i, _i, _vp, _vp, _i, _vp, _vp, _vp]),
    "segma_stitch": (_i, [_vp, _i, _i, _i, _i, _i, _vp, _i64, _vp]),
    "segma_decode_workspace_bytes": (_sz, [_i64, _i, _i]),
    "segma_decode_intervals": (_i, [_vp, _vp, _i, _i, _vp, _i, _vp, _i64, _vp, _vp, _sz, _vp]),
    "segma_decode_intervals_hysteresis": (_i, [_vp, _vp, _i, _i, _vp, _vp, _vp, _i64, _vp, _vp, _sz, _vp]),
    "segma_postprocess_intervals": (_i, [_vp, _i64, _i, _i, _vp, _vp, _i64, _vp, _vp]),
    "segma_threshold_histogram": (_i, [_vp, _vp, _i64, _i, _vp, _i, _vp, _vp]),
    "segma_threshold_mask": (_i, [_vp, _i64, _i, _vp, _i, _vp, _vp]),
}


def lib_path() -> Path:
    from .build import LIB_PATH

    return LIB_PATH


def load(build_if_missing: bool = True):
    """Load (building first if needed) the CUDA library; raises if that is impossible."""
    global _LIB
    if _LIB is not None:
        return _LIB
    import os

    from . import build as _build

    debug = os.environ.get("SEGMA_DEBUG", "0") not in ("", "0")  # the same kernels with device-side bounds asserts
    path = _build.DEBUG_LIB_PATH if debug else _build.LIB_PATH
    if not path.exists() or not _build.is_current(debug):
        # missing, or built from other sources than the ones in the tree (a stale library would be called with this
        # file's argument lists): rebuild, or refuse
        what = "missing" if not path.exists() else "older than its sources"
        if not build_if_missing:
            raise SegmaNativeError(f"{path} is {what}: run `python -m segma_b200.build`")
        try:
            _build.build(debug=debug)
        except Exception as e:  # noqa: BLE001
            raise SegmaNativeError(
                f"libsegma_b200.so is {what} and could not be built ({e}); segma_b200 has no CPU fallback"
            ) from e
    lib = C.CDLL(str(path))
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _LIB = lib
    return lib


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = load().segma_last_error().decode(errors="replace")
        raise SegmaNativeError(f"{what} failed with status {rc}: {msg}")
