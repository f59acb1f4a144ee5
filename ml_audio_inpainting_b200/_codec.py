"""ctypes binding of the native host codec (include/aip_codec.h, csrc/aip_flac.c -> lib/libaip_codec.so).

The library is plain C built with gcc (``_build.build_codec``: rebuilt when the source hash changes); a missing compiler and a
missing library raise -- there is no pure-Python fallback on the product path (the Python codec is the checker, oracle/flac_port.py).
ctypes releases the GIL around every call, so files decode in parallel from a thread pool.
"""
from __future__ import annotations

import ctypes as C
import threading

from . import _build

_P, _I32, _I64, _SZ = C.c_void_p, C.c_int32, C.c_int64, C.c_size_t


class FlacInfoC(C.Structure):
    _fields_ = [("sample_rate", C.c_int32), ("channels", C.c_int32), ("bits_per_sample", C.c_int32),
                ("min_blocksize", C.c_int32), ("max_blocksize", C.c_int32), ("total_samples", C.c_int64),
                ("md5", C.c_uint8 * 16)]


# name -> (restype, argtypes): mirrors include/aip_codec.h (tests/test_codec.py checks the two against each other)
SIGNATURES = {
    "aip_flac_info_read": (C.c_int, [_P, _SZ, C.POINTER(FlacInfoC)]),
    "aip_flac_decode": (C.c_int64, [_P, _SZ, _I64, _P, _I64, C.POINTER(FlacInfoC)]),
    "aip_flac_encode16": (C.c_int64, [_P, _I64, _I32, _I32, _I32, C.c_char_p, _P, _SZ]),
    "aip_codec_status_string": (C.c_char_p, [C.c_int]),
}

_lib = None
_lock = threading.Lock()


def load() -> C.CDLL:
    """Build (if the source hash changed) and load the codec once; safe to call from the I/O thread pools."""
    global _lib
    if _lib is None:
        with _lock:
            if _lib is None:
                path = _build.build_codec()
                lib = C.CDLL(str(path))
                for name, (res, args) in SIGNATURES.items():
                    fn = getattr(lib, name)
                    fn.restype = res
                    fn.argtypes = args
                _lib = lib
    return _lib


def check(status: int, where: str) -> None:
    if status < 0:
        raise ValueError(f"{where}: FLAC: {load().aip_codec_status_string(int(status)).decode()}")
