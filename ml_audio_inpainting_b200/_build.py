"""In-tree build of the native libraries (no JIT cache: the built .so files travel with the repo).

  lib/libaip_b200.so   nvcc, sm_100a SASS only -- the product (one object per csrc/aip_*.cu, compiled in parallel)
  lib/libaip_emul.so   g++, host replay of the kernels' phase functions -- test support only
  lib/libaip_codec.so  gcc, the host-side FLAC codec of load_audio / save_audio (csrc/aip_flac.c) -- product, no CUDA

A library is rebuilt when the SHA-256 of its sources + flags differs from the one recorded next to it
(``lib/<name>.sha256``), never by file times: a pushed prebuilt .so whose sources changed is recompiled, and
``build(force=True)`` always compiles.
"""
from __future__ import annotations

import hashlib
import os
import shutil
import subprocess
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

PKG = Path(__file__).resolve().parent
CSRC = PKG / "csrc"
LIB = PKG / "lib"
OBJ = LIB / "obj"
CUDA_LIB = LIB / "libaip_b200.so"
EMUL_LIB = LIB / "libaip_emul.so"
CODEC_LIB = LIB / "libaip_codec.so"

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-std=c++17", "-O3", "-lineinfo",
              "-Xptxas", "-v", "-Xcompiler", "-fPIC"]
CUDA_UNITS = ["aip_fwd.cu", "aip_pow2.cu", "aip_inv.cu", "aip_mel.cu", "aip_misc.cu"]
CUDA_HEADERS = ["aip_core.cuh", "aip_tiles.cuh", "aip_device.cuh", "aip_host.h", "aip_twiddles.inc"]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and Path(cand).exists():
            return cand
    raise RuntimeError("nvcc not found (set NVCC=/path/to/nvcc)")


def _digest(sources, flags) -> str:
    h = hashlib.sha256()
    h.update("\0".join(flags).encode())
    for s in sources:
        h.update(str(Path(s).name).encode())
        h.update(Path(s).read_bytes())
    return h.hexdigest()


def _stale(target: Path, digest: str) -> bool:
    stamp = target.with_suffix(".sha256")
    return not (target.exists() and stamp.exists() and stamp.read_text().strip() == digest)


def cuda_units():
    return [CSRC / u for u in CUDA_UNITS if (CSRC / u).exists()]


def cuda_sources():
    return cuda_units() + [CSRC / h for h in CUDA_HEADERS] + [PKG.parent / "include" / "aip_b200.h"]


def build_cuda(force: bool = False, verbose: bool = False, extra_flags=()) -> Path:
    LIB.mkdir(exist_ok=True)
    flags = [*NVCC_FLAGS, *extra_flags]
    digest = _digest(cuda_sources(), flags)
    if not (force or _stale(CUDA_LIB, digest)):
        return CUDA_LIB
    OBJ.mkdir(exist_ok=True)
    nvcc = _nvcc()
    units = cuda_units()

    def compile_unit(src: Path):
        obj = OBJ / (src.stem + ".o")
        res = subprocess.run([nvcc, *flags, "-c", "-o", str(obj), str(src)], capture_output=True, text=True)
        return src, obj, res

    with ThreadPoolExecutor(max_workers=min(len(units), os.cpu_count() or 1)) as pool:
        done = list(pool.map(compile_unit, units))
    log = "".join(f"==== {src.name}\n{res.stderr}" for src, _, res in done)
    (LIB / "ptxas_info.txt").write_text(log)
    if verbose:
        print(log)
    for src, _, res in done:
        if res.returncode != 0:
            raise RuntimeError(f"nvcc failed on {src.name}:\n" + res.stdout + res.stderr)
    res = subprocess.run([nvcc, "-shared", "-o", str(CUDA_LIB), *[str(o) for _, o, _ in done]], capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc link failed:\n" + res.stdout + res.stderr)
    CUDA_LIB.with_suffix(".sha256").write_text(digest)
    return CUDA_LIB


def build_emul(force: bool = False) -> Path:
    LIB.mkdir(exist_ok=True)
    srcs = [CSRC / "aip_emul.cpp", CSRC / "aip_tiles.cuh", CSRC / "aip_core.cuh", CSRC / "aip_twiddles.inc"]
    flags = ["-O2", "-std=c++17", "-shared", "-fPIC"]
    digest = _digest(srcs, flags)
    if force or _stale(EMUL_LIB, digest):
        cmd = ["g++", *flags, "-o", str(EMUL_LIB), str(CSRC / "aip_emul.cpp")]
        res = subprocess.run(cmd, capture_output=True, text=True)
        if res.returncode != 0:
            raise RuntimeError("g++ failed:\n" + res.stdout + res.stderr)
        EMUL_LIB.with_suffix(".sha256").write_text(digest)
    return EMUL_LIB


def build_codec(force: bool = False) -> Path:
    """The host-side FLAC codec (plain C, no CUDA): gcc -> lib/libaip_codec.so."""
    LIB.mkdir(exist_ok=True)
    srcs = [CSRC / "aip_flac.c", PKG.parent / "include" / "aip_codec.h"]
    flags = ["-O2", "-std=c11", "-shared", "-fPIC"]
    digest = _digest(srcs, flags)
    if force or _stale(CODEC_LIB, digest):
        cc = shutil.which("gcc") or shutil.which("cc")
        if not cc:
            raise RuntimeError("no C compiler (gcc) to build lib/libaip_codec.so: the host codec has no pure-Python fallback")
        res = subprocess.run([cc, *flags, "-o", str(CODEC_LIB), str(CSRC / "aip_flac.c")], capture_output=True, text=True)
        if res.returncode != 0:
            raise RuntimeError("gcc failed:\n" + res.stdout + res.stderr)
        CODEC_LIB.with_suffix(".sha256").write_text(digest)
    return CODEC_LIB


if __name__ == "__main__":
    print(build_cuda(force=True, verbose=True))
    print(build_emul(force=True))
    print(build_codec(force=True))
