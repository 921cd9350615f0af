"""Builds ``libsegma_b200.so`` in-tree with nvcc for sm_100a (cross-compiles without a GPU).

    python -m segma_b200.build [--force]

The shared library has no torch dependency: it links only the static CUDA runtime and fetches the one
driver entry point it needs (cuTensorMapEncodeTiled) at run time.
"""
from __future__ import annotations

import hashlib
import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

PKG_DIR = Path(__file__).resolve().parent
CSRC = PKG_DIR / "csrc"
INCLUDE = PKG_DIR.parent / "include"
LIB_DIR = PKG_DIR / "lib"
LIB_PATH = LIB_DIR / "libsegma_b200.so"
OBJ_DIR = PKG_DIR / "build"
#: the same sources with -DSEGMA_DEBUG (device-side bounds asserts); loaded instead of the product library when
#: SEGMA_DEBUG=1 is set (compute-sanitizer is not available on the GPU pool)
DEBUG_LIB_PATH = LIB_DIR / "libsegma_b200_debug.so"
DEBUG_OBJ_DIR = PKG_DIR / "build_debug"

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC",
    "-Xcompiler", "-fvisibility=hidden",
    "--expt-relaxed-constexpr",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and Path(cand).exists():
            return cand
    raise RuntimeError("nvcc not found; the CUDA extension cannot be built")


def sources() -> list[Path]:
    return sorted(CSRC.glob("*.cu"))


def _digest() -> str:
    h = hashlib.sha256()
    for p in sources() + sorted(CSRC.glob("*.cuh")) + sorted(INCLUDE.glob("*.h")):
        h.update(p.name.encode())
        h.update(p.read_bytes())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def is_current(debug: bool = False) -> bool:
    stamp = LIB_DIR / ("build_debug.sha256" if debug else "build.sha256")
    lib = DEBUG_LIB_PATH if debug else LIB_PATH
    return lib.exists() and stamp.exists() and stamp.read_text().strip() == _digest()


def build(force: bool = False, verbose: bool = False, debug: bool = False) -> Path:
    lib_path, obj_dir = (DEBUG_LIB_PATH, DEBUG_OBJ_DIR) if debug else (LIB_PATH, OBJ_DIR)
    if not force and is_current(debug):
        return lib_path
    nvcc = _nvcc()
    obj_dir.mkdir(exist_ok=True)
    LIB_DIR.mkdir(exist_ok=True)
    flags = NVCC_FLAGS + (["-DSEGMA_DEBUG"] if debug else [])

    def compile_one(src: Path) -> Path:
        obj = obj_dir / (src.stem + ".o")
        cmd = [nvcc, *flags, "-I", str(INCLUDE), "-c", str(src), "-o", str(obj)]
        if verbose:
            print(" ".join(cmd), flush=True)
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed on {src.name}:\n{r.stdout}\n{r.stderr}")
        return obj

    with ThreadPoolExecutor(max_workers=min(8, len(sources()))) as ex:
        objs = list(ex.map(compile_one, sources()))
    cmd = [nvcc, "-shared", "-o", str(lib_path), *map(str, objs), "-gencode", "arch=compute_100a,code=sm_100a",
           "-cudart", "static"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    (LIB_DIR / ("build_debug.sha256" if debug else "build.sha256")).write_text(_digest() + "\n")
    return lib_path


if __name__ == "__main__":
    path = build(force="--force" in sys.argv, verbose=True, debug="--debug" in sys.argv)
    print(f"built {path}")
