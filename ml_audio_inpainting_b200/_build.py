"""In-tree build of the native libraries (no JIT cache: the built .so files travel with the repo).

  lib/libaip_b200.so   nvcc, sm_100a SASS only -- the product
  lib/libaip_emul.so   g++, host replay of the kernels' phase functions -- test support only
"""
from __future__ import annotations

import os
import shutil
import subprocess
from pathlib import Path

PKG = Path(__file__).resolve().parent
CSRC = PKG / "csrc"
LIB = PKG / "lib"
CUDA_LIB = LIB / "libaip_b200.so"
EMUL_LIB = LIB / "libaip_emul.so"

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-std=c++17", "-O3", "-lineinfo",
              "-Xptxas", "-v", "-shared", "-Xcompiler", "-fPIC"]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and Path(cand).exists():
            return cand
    raise RuntimeError("nvcc not found (set NVCC=/path/to/nvcc)")


def _stale(target: Path, sources) -> bool:
    if not target.exists():
        return True
    t = target.stat().st_mtime
    return any(Path(s).stat().st_mtime > t for s in sources)


def cuda_sources():
    return [CSRC / "aip_kernels.cu", CSRC / "aip_tiles.cuh", CSRC / "aip_core.cuh",
            CSRC / "aip_twiddles.inc", PKG.parent / "include" / "aip_b200.h"]


def build_cuda(force: bool = False, verbose: bool = False) -> Path:
    LIB.mkdir(exist_ok=True)
    if force or _stale(CUDA_LIB, cuda_sources()):
        cmd = [_nvcc(), *NVCC_FLAGS, "-o", str(CUDA_LIB), str(CSRC / "aip_kernels.cu")]
        res = subprocess.run(cmd, capture_output=True, text=True)
        (LIB / "ptxas_info.txt").write_text(res.stderr)
        if verbose:
            print(res.stderr)
        if res.returncode != 0:
            raise RuntimeError("nvcc failed:\n" + res.stdout + res.stderr)
    return CUDA_LIB


def build_emul(force: bool = False) -> Path:
    LIB.mkdir(exist_ok=True)
    srcs = [CSRC / "aip_emul.cpp", CSRC / "aip_tiles.cuh", CSRC / "aip_core.cuh", CSRC / "aip_twiddles.inc"]
    if force or _stale(EMUL_LIB, srcs):
        cmd = ["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-o", str(EMUL_LIB), str(CSRC / "aip_emul.cpp")]
        res = subprocess.run(cmd, capture_output=True, text=True)
        if res.returncode != 0:
            raise RuntimeError("g++ failed:\n" + res.stdout + res.stderr)
    return EMUL_LIB


if __name__ == "__main__":
    print(build_cuda(force=True, verbose=True))
    print(build_emul(force=True))
