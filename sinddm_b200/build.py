"""In-tree build of libsinddm_b200.so (sm_100a only).

`python -m sinddm_b200.build` or `build_library()`; objects go to sinddm_b200/csrc/build/, the shared
library to sinddm_b200/lib/.  nvcc cross-compiles without a GPU, so this runs on the CPU-only dev box and
the resulting .so travels to the B200 box with the repo snapshot.
"""
from __future__ import annotations

import concurrent.futures
import os
import shutil
import subprocess
import sys
from pathlib import Path

PKG_DIR = Path(__file__).resolve().parent
CSRC = PKG_DIR / "csrc"
OBJ_DIR = CSRC / "build"
LIB_DIR = PKG_DIR / "lib"
LIB_PATH = LIB_DIR / "libsinddm_b200.so"

SOURCES = [
    "host_common.cu",
    "tc_conv.cu",
    "tc_wgrad.cu",
    "simt_conv.cu",
    "simt_misc.cu",
    "dw_kernels.cu",
    "cond.cu",
    "diffusion_ops.cu",
    "fused_optim.cu",
    "net.cu",
    "capi.cu",
]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC",
    "-cudart", "shared",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and Path(cand).exists():
            return cand
    raise RuntimeError("nvcc not found; sinddm_b200 has no non-CUDA path")


def _newer(src: Path, dst: Path) -> bool:
    return (not dst.exists()) or src.stat().st_mtime > dst.stat().st_mtime


def _compile(nvcc: str, src: Path, obj: Path, verbose: bool) -> None:
    cmd = [nvcc, *NVCC_FLAGS, "-c", str(src), "-o", str(obj)]
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
    proc = subprocess.run(cmd, capture_output=True, text=True)
    if proc.returncode != 0:
        raise RuntimeError(f"nvcc failed for {src.name}:\n{proc.stdout}\n{proc.stderr}")
    if verbose:
        sys.stderr.write(proc.stderr)


def build_library(force: bool = False, verbose: bool = False) -> Path:
    nvcc = _nvcc()
    OBJ_DIR.mkdir(parents=True, exist_ok=True)
    LIB_DIR.mkdir(parents=True, exist_ok=True)
    headers = list(CSRC.glob("*.h")) + list(CSRC.glob("*.cuh")) + [PKG_DIR.parent / "include" / "sinddm_b200.h"]
    newest_header = max(h.stat().st_mtime for h in headers)
    jobs = []
    for name in SOURCES:
        src = CSRC / name
        obj = OBJ_DIR / (src.stem + ".o")
        if force or _newer(src, obj) or obj.stat().st_mtime < newest_header:
            jobs.append((src, obj))
    if jobs:
        with concurrent.futures.ThreadPoolExecutor(max_workers=min(8, len(jobs))) as pool:
            list(pool.map(lambda j: _compile(nvcc, j[0], j[1], verbose), jobs))
    objs = [OBJ_DIR / (Path(n).stem + ".o") for n in SOURCES]
    if jobs or not LIB_PATH.exists():
        cmd = [nvcc, "-shared", "-cudart", "shared", "-gencode", "arch=compute_100a,code=sm_100a",
               "-Xlinker", "-rpath=/usr/local/cuda/lib64", "-o", str(LIB_PATH), *map(str, objs)]
        proc = subprocess.run(cmd, capture_output=True, text=True)
        if proc.returncode != 0:
            raise RuntimeError(f"link failed:\n{proc.stdout}\n{proc.stderr}")
    return LIB_PATH


if __name__ == "__main__":
    path = build_library(force="--force" in sys.argv, verbose="-v" in sys.argv)
    print(path)
