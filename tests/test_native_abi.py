"""The C-ABI library builds for sm_100a without a GPU, loads, and exports every symbol that
include/segma_b200.h declares; the ctypes table mirrors the header one to one.  No compute calls here."""
import ctypes
import re
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
HEADER = ROOT / "include" / "segma_b200.h"


def _declared():
    text = re.sub(r"/\*.*?\*/", "", HEADER.read_text(), flags=re.S)
    return sorted(set(re.findall(r"SEGMA_API\s+[\w\s\*]+?\b(segma_\w+)\s*\(", text)))


@pytest.fixture(scope="module")
def lib():
    from segma_b200 import build

    return ctypes.CDLL(str(build.build()))


def test_header_declares_the_expected_surface():
    names = _declared()
    for must in ("segma_logmel", "segma_gemm_f16", "segma_attention", "segma_attention_rel", "segma_layernorm", "segma_lstm_layer", "segma_heads",
                 "segma_stitch", "segma_decode_intervals", "segma_threshold_mask", "segma_last_error"):
        assert must in names


def test_library_exports_every_declared_symbol(lib):
    missing = [n for n in _declared() if not hasattr(lib, n)]
    assert not missing, missing


def test_ctypes_table_matches_header():
    from segma_b200._native import SIGNATURES

    assert sorted(SIGNATURES) == _declared()


def test_version_and_error_text_without_gpu(lib):
    assert lib.segma_version() >= 100
    lib.segma_last_error.restype = ctypes.c_char_p
    assert isinstance(lib.segma_last_error(), bytes)
    lib.segma_decode_workspace_bytes.restype = ctypes.c_size_t
    lib.segma_decode_workspace_bytes.argtypes = [ctypes.c_int64, ctypes.c_int, ctypes.c_int]
    assert lib.segma_decode_workspace_bytes(180_000, 1, 4) > 180_000 * 4
    lib.segma_logmel_scratch_bytes.restype = ctypes.c_size_t
    assert lib.segma_logmel_scratch_bytes(128, 64000) >= 128 * 80 * 402 * 4


def test_builtin_mel_filters_match_transformers_values(lib):
    """The filterbank built in C++ (float64 slaney formula) equals the oracle's / transformers' matrix."""
    import numpy as np

    from oracle import segma_oracle as O

    out = np.empty((201, 80), dtype=np.float32)
    lib.segma_logmel_get_filters.argtypes = [ctypes.c_void_p]
    assert lib.segma_logmel_get_filters(out.ctypes.data) == 0
    ref = O.whisper_mel_filters().astype(np.float32)
    assert np.array_equal(out != 0, ref != 0) and int((ref != 0).sum()) == 391
    assert np.abs(out - ref).max() <= 1e-9


def test_product_has_no_cpu_path():
    import torch

    from segma_b200 import ops
    from segma_b200._native import SegmaNativeError

    with pytest.raises(SegmaNativeError):
        ops.threshold_mask(torch.zeros((4, 4)), [0.5] * 4)
    from segma_b200.inference import _cuda_device

    with pytest.raises(SegmaNativeError):
        _cuda_device("cpu")


def test_product_does_not_import_the_oracle():
    pkg = ROOT / "segma_b200"
    for f in pkg.rglob("*.py"):
        assert "oracle" not in re.sub(r'""".*?"""', "", f.read_text(), flags=re.S).replace("# oracle", ""), f


def test_debug_flavour_builds_and_exports_the_same_abi():
    """``python -m segma_b200.build --debug``: the same sources with device-side bounds asserts (SEGMA_DEBUG=1 loads
    it; compute-sanitizer is closed on the GPU pool).  Same symbols as the product library."""
    import ctypes

    from segma_b200 import _native, build

    path = build.build(debug=True)
    assert path.name == "libsegma_b200_debug.so" and path.exists()
    lib = ctypes.CDLL(str(path))
    assert not [name for name in _native.SIGNATURES if not hasattr(lib, name)]
