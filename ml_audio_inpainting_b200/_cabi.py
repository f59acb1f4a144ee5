"""ctypes binding of include/aip_b200.h (the C ABI; plain pointers and sizes, no torch types).

There is no fallback: if the shared object is missing or the device is not a B200 this raises.
"""
from __future__ import annotations

import ctypes as C
import threading
from pathlib import Path

LIB_PATH = Path(__file__).resolve().parent / "lib" / "libaip_b200.so"

AIP_OK = 0
MAG_NONE, MAG_ABS, MAG_LOG10_EPS, MAG_LOG1P_POW, MAG_POW = 0, 1, 2, 3, 4
DOM_LINEAR, DOM_POW10, DOM_DB, DOM_EXPM1 = 0, 1, 2, 3


class StftDesc(C.Structure):
    """struct aip_stft_desc"""
    _fields_ = [("n_fft", C.c_int32), ("hop", C.c_int32), ("center", C.c_int32),
                ("win_length", C.c_int32), ("window", C.c_void_p)]


class AipError(RuntimeError):
    def __init__(self, status: int, what: str, where: str):
        super().__init__(f"{where}: {what} (status {status})")
        self.status = status


_P, _I64, _I32, _F, _SZ = C.c_void_p, C.c_int64, C.c_int32, C.c_float, C.c_size_t
_D = C.POINTER(StftDesc)

# name -> (restype, argtypes): exactly the entry points include/aip_b200.h declares
SIGNATURES = {
    "aip_num_frames": (_I64, [_I64, _I32, _I32, _I32]),
    "aip_istft_length": (_I64, [_I64, _I32, _I32, _I32, _I64]),
    "aip_stft_fwd_f32": (C.c_int, [_D, _P, _I64, _I64, _I64, _P, _P, _P, _I32, _I32, _F, _F, _I64,
                                   _P, _P, _P, _P, _P]),
    "aip_istft_f32": (C.c_int, [_D, _P, _P, _P, _I32, _P, _I64, _I64, _I64, _P, _P, _I64, _P, _SZ, _P]),
    "aip_istft_normalized_f32": (C.c_int, [_D, _P, _P, _P, _I32, _P, _I64, _I64, _I64, _P, _P, _I64, _P, _P, _I64, _P, _SZ, _P]),
    "aip_istft_blend_f32": (C.c_int, [_D, _P, _P, _P, _P, _I32, _I64, _I64, _I64, _P, _P, _I64, _P, _SZ, _P]),
    "aip_istft_handoff_f32": (C.c_int, [_D, _P, _P, _P, _I32, _P, _I32, _I64, _I64, _I64, _P, _P, _I64, _P, _P, _I64, _P, _SZ, _P]),
    "aip_istft_workspace_bytes": (_SZ, [_D, _I64, _I64]),
    "aip_inv_window_sumsquare_f32": (C.c_int, [_D, _I64, _I64, _P, _I64, _P]),
    "aip_random_phasors_f32": (C.c_int, [_P, _I64, C.c_uint64, _P]),
    "aip_griffinlim_f32": (C.c_int, [_D, _P, _P, _P, _I64, _I64, _I32, _F, _P, _P, _I64, _P, _SZ, _P]),
    "aip_griffinlim_c64_f32": (C.c_int, [_D, _P, _P, _P, _I64, _I64, _I32, _F, _P, _P, _I64, _P, _SZ, _P]),
    "aip_mel_project_f32": (C.c_int, [_P, _P, _P, _I64, _I64, _I64, _I64, _P, _P]),
    "aip_mel_inverse_f32": (C.c_int, [_P, _P, _I64, _I64, _I64, _I64, _I32, _P, _P, _P]),
    "aip_db_heuristic_f32": (C.c_int, [_P, _I64, _I64, _P, _P]),
    "aip_gap_zero_f32": (C.c_int, [_P, _I64, _P, _I64, _I64, _I64, _P, _P]),
    "aip_gap_mask_f32": (C.c_int, [_P, _I64, _I64, _I64, _P, _P]),
    "aip_frame_mask_f32": (C.c_int, [_P, _I64, _I64, _I64, _P, _I32, _P]),
    "aip_stft_gap_variants_workspace_bytes": (_SZ, [_I64, _I64]),
    "aip_stft_gap_variants_f32": (C.c_int, [_D, _P, _I64, _I64, _I64, _I64, _P, _I32, _I32, _F, _I64, _P, _P, _P, _SZ, _P]),
    "aip_peak_normalize_f32": (C.c_int, [_P, _I64, _P, _I64, _I64, _I64, _P, _P]),
    "aip_wave_to_pcm16_f32": (C.c_int, [_P, _I64, _P, _I64, _I64, _I64, _I32, _P, _P]),
    "aip_status_string": (C.c_char_p, [C.c_int]),
    "aip_debug_reload_env": (None, []),
    "aip_version": (C.c_char_p, []),
    "aip_device_supported": (C.c_int, []),
}

_lib = None
_lock = threading.Lock()


def load() -> C.CDLL:
    """Load libaip_b200.so (built in-tree by ``_build.build_cuda``) and type its entry points."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not LIB_PATH.exists():
            raise ImportError(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(nvcc, sm_100a).  ml_audio_inpainting_b200 has no CPU fallback.")
        lib = C.CDLL(str(LIB_PATH))
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


def check(status: int, where: str) -> None:
    if status != AIP_OK:
        raise AipError(status, load().aip_status_string(status).decode(), where)
