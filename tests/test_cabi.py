"""CPU tier: the C-ABI library loads and exports every symbol include/aip_b200.h declares (no compute calls
without a GPU), the host-only entry points answer correctly, and the product fails loudly without CUDA."""
import ctypes as C
import re
from pathlib import Path

import numpy as np
import pytest

from ml_audio_inpainting_b200 import _build, _cabi

ROOT = Path(__file__).resolve().parents[1]


@pytest.fixture(scope="module")
def lib():
    _build.build_cuda()
    return _cabi.load()


def declared_symbols():
    hdr = (ROOT / "include" / "aip_b200.h").read_text()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    return sorted(set(re.findall(r"\b(aip_[a-z0-9_]+)\s*\(", hdr)))


def test_every_declared_symbol_is_exported_and_bound(lib):
    syms = declared_symbols()
    assert len(syms) >= 15
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/aip_b200.h but not exported"
    assert sorted(_cabi.SIGNATURES) == syms, "ctypes table and header disagree"


def test_header_cites_reference_lines():
    hdr = (ROOT / "include" / "aip_b200.h").read_text()
    for cite in ("utils.py:192-234", "utils.py:316-327", "utils.py:328-332", "utils.py:313-314", "utils.py:84",
                 "models/CNNBLSTM/dataset.py", "models/GAN/dataset.py", "models/model_eval.py", "add_gaps.py"):
        assert cite in hdr


def test_host_only_entry_points(lib):
    assert lib.aip_num_frames(80000, 512, 192, 1) == 417
    assert lib.aip_num_frames(160000, 512, 192, 1) == 834
    assert lib.aip_num_frames(80000, 512, 128, 1) == 626
    assert lib.aip_num_frames(512, 512, 192, 0) == 1
    assert lib.aip_num_frames(100, 512, 192, 0) < 0
    assert lib.aip_istft_length(417, 512, 192, 1, 0) == 79872
    assert lib.aip_istft_length(626, 512, 128, 1, 0) == 80000
    assert lib.aip_istft_length(417, 512, 192, 1, 80000) == 80000
    assert lib.aip_istft_length(10, 512, 192, 0, 0) == 512 + 192 * 9
    assert b"sm_100a" in lib.aip_version()
    assert lib.aip_status_string(0) == b"ok"
    assert b"no fallback" in lib.aip_status_string(-3)
    d = _cabi.StftDesc(512, 192, 1, 0, None)
    assert lib.aip_istft_workspace_bytes(C.byref(d), 4, 100) == 0
    d2 = _cabi.StftDesc(2048, 512, 1, 0, None)      # tiled radix-16 inverse, overlap-add fused: no workspace either
    assert lib.aip_istft_workspace_bytes(C.byref(d2), 4, 100) == 0
    d3 = _cabi.StftDesc(2048, 16, 1, 0, None)       # more than two tiles reach a sample: frames go through the workspace
    assert lib.aip_istft_workspace_bytes(C.byref(d3), 4, 100) == 4 * 100 * 2048 * 4
    d4 = _cabi.StftDesc(4096, 1024, 1, 0, None)     # outside the tiled kernels' range
    assert lib.aip_istft_workspace_bytes(C.byref(d4), 4, 100) == 4 * 100 * 4096 * 4
    d5 = _cabi.StftDesc(512, 191, 1, 0, None)       # n_fft 512 with an odd hop: off the fused kernel, on the tiled one
    assert lib.aip_istft_workspace_bytes(C.byref(d5), 1, 10) == 0


def test_sass_is_sm100a_with_tma(lib):
    """The shipped object holds sm_100a code and the TMA bulk copy (UBLKCP) + mbarrier (SYNCS) instructions."""
    import shutil
    import subprocess
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not Path(cuobjdump).exists():
        pytest.skip("cuobjdump not available")
    out = subprocess.run([cuobjdump, "-lelf", str(_cabi.LIB_PATH)], capture_output=True, text=True).stdout
    assert "sm_100a" in out
    sass = subprocess.run([cuobjdump, "-sass", str(_cabi.LIB_PATH)], capture_output=True, text=True).stdout
    assert "UBLKCP" in sass and "SYNCS" in sass


def test_no_cpu_fallback():
    torch = pytest.importorskip("torch")
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from ml_audio_inpainting_b200 import frontend, spectral
    with pytest.raises(RuntimeError):
        spectral.stft(torch.zeros(2, 1000), None)
    with pytest.raises(RuntimeError):
        frontend.cnnblstm_batch(torch.zeros(2, 80000))
    with pytest.raises((RuntimeError, AssertionError)):
        spectral.get_plan(512, 192, 384)
    # a device-side entry point on a machine without a usable device reports an error, not a CPU result
    lib = _cabi.load()
    assert lib.aip_device_supported() == 0
    x = np.zeros(8, np.float32)
    rc = lib.aip_gap_mask_f32(x.ctypes.data, 8, 1, 8, None, None)
    assert rc != 0


def test_headers_are_plain_c_and_a_c_host_links_the_library(lib, tmp_path):
    """include/*.h compile as C99 and as C++11 on their own, and a host written in C -- no Python, no torch -- links the
    shared library and calls its host-only entry points (the device entry points need a GPU: tests/test_gpu_round2.py calls
    them raw through ctypes)."""
    import shutil
    import subprocess
    from ml_audio_inpainting_b200 import _build
    gcc = shutil.which("gcc")
    if not gcc:
        pytest.skip("no gcc")
    inc = ROOT / "include"
    for h in ("aip_b200.h", "aip_codec.h"):
        for cmd in ([gcc, "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-fsyntax-only", "-x", "c"],
                    [shutil.which("g++") or gcc, "-std=c++11", "-fsyntax-only", "-x", "c++"]):
            res = subprocess.run(cmd + [str(inc / h)], capture_output=True, text=True)
            assert res.returncode == 0, res.stderr
    src = tmp_path / "host.c"
    src.write_text('''
#include <stdio.h>
#include <string.h>
#include "aip_b200.h"
#include "aip_codec.h"
int main(void) {
  if (aip_num_frames(160000, 512, 192, 1) != 834) return 1;
  if (aip_istft_length(834, 512, 192, 1, 0) != 159936) return 2;
  if (strcmp(aip_status_string(AIP_OK), "ok") != 0) return 3;
  if (!strstr(aip_version(), "sm_100a")) return 4;
  if (aip_flac_info_read((const unsigned char*)"nope", 4, &(aip_flac_info){0}) >= 0) return 5;
  printf("%s\\n", aip_version());
  return 0;
}
''')
    exe = tmp_path / "host"
    libdir = _build.CUDA_LIB.parent
    _build.build_codec()
    res = subprocess.run([gcc, "-std=c99", "-I", str(inc), "-o", str(exe), str(src), f"-L{libdir}", "-laip_b200", "-laip_codec",
                          f"-Wl,-rpath,{libdir}"], capture_output=True, text=True)
    assert res.returncode == 0, res.stderr
    run = subprocess.run([str(exe)], capture_output=True, text=True)
    assert run.returncode == 0 and "sm_100a" in run.stdout, (run.returncode, run.stdout, run.stderr)
