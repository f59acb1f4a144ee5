"""ctypes access to lib/libaip_emul.so: the host replay of the n_fft = 512 kernels (test support)."""
import ctypes as C

import numpy as np

from ml_audio_inpainting_b200 import _build

_lib = None
FP = C.POINTER(C.c_float)
IP = C.POINTER(C.c_int)


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(str(_build.build_emul()))
    return _lib


def _f(a):
    return None if a is None else a.ctypes.data_as(FP)


def _i(a):
    return None if a is None else a.ctypes.data_as(IP)


def stft(wave, hop, window, center=True, gap_samples=None, zero_frames=None, mask_frames=None,
         mask_in_gap_is_one=True, mag_kind=0, eps=1e-9, power=1.0, t_out=None, want_spec=True,
         want_phase=False, want_mask=False, vec_ok=True, win_length=0):
    """win_length > 0 promises that `window` is a win_length window centre-padded with zeros (zero-tap pruning)."""
    wave = np.ascontiguousarray(wave, dtype=np.float32)
    B, L = wave.shape
    pad = 256 if center else 0
    T = 1 + (L + 2 * pad - 512) // hop
    t_out = T if t_out is None else t_out
    window = np.ascontiguousarray(window, dtype=np.float32)
    spec = np.zeros((B, 257, t_out, 2), np.float32) if want_spec else None
    mag = np.zeros((B, 257, t_out), np.float32) if mag_kind else None
    phase = np.zeros((B, 257, t_out), np.float32) if want_phase else None
    mask = np.full((B, 257, t_out), -1, np.float32) if want_mask else None
    conv = lambda a: None if a is None else np.ascontiguousarray(a, dtype=np.int32)
    g, z, m = conv(gap_samples), conv(zero_frames), conv(mask_frames)
    rc = lib().emul_stft512_fwd(_f(wave), B, L, C.c_longlong(L), hop, int(center), int(win_length), _f(window), _i(g), _i(z), _i(m),
                                int(mask_in_gap_is_one), mag_kind, C.c_float(eps), C.c_float(power), t_out,
                                _f(spec), _f(mag), _f(phase), _f(mask), int(vec_ok))
    assert rc == 0, rc
    out = {}
    if want_spec:
        out["spec"] = spec[..., 0] + 1j * spec[..., 1]
    if mag_kind:
        out["mag"] = mag
    if want_phase:
        out["phase"] = phase
    if want_mask:
        out["mask"] = mask
    return out


def stft_variants(wave, hop, window, G, gap_samples, mag_kind=2, eps=1e-9, center=True, t_out=None, vec_ok=True,
                  win_length=0, gap_len_max=None):
    """aip_stft_gap_variants_f32 replayed: the clean magnitudes come from the plain replay (``stft``)."""
    wave = np.ascontiguousarray(wave, dtype=np.float32)
    N, L = wave.shape
    pad = 256 if center else 0
    T = 1 + (L + 2 * pad - 512) // hop
    t_out = T if t_out is None else t_out
    window = np.ascontiguousarray(window, dtype=np.float32)
    g = np.ascontiguousarray(gap_samples, dtype=np.int32).reshape(N * G, 2)
    if gap_len_max is None:
        gap_len_max = int((g[:, 1] - g[:, 0]).max())
    clean = stft(wave, hop, window, center=center, mag_kind=mag_kind, eps=eps, t_out=t_out, want_spec=False,
                 vec_ok=vec_ok, win_length=win_length)["mag"]
    mag = np.full((N * G, 257, t_out), np.nan, np.float32)
    rc = lib().emul_stft512_variants(_f(wave), N, L, C.c_longlong(L), hop, int(center), int(win_length), _f(window), G, _i(g),
                                     int(gap_len_max), mag_kind, C.c_float(eps), t_out, _f(clean), _f(mag), int(vec_ok))
    assert rc == 0, rc
    return {"mag": mag, "clean_mag": clean}


def istft(hop, window, inv_wss, spec=None, mag=None, phase=None, mag_domain=0, db_flags=None, center=True, length=0,
          blend_in=None, blend_mask=None, win_length=0, peaks=None):
    window = np.ascontiguousarray(window, dtype=np.float32)
    inv_wss = np.ascontiguousarray(inv_wss, dtype=np.float32)
    if spec is not None:
        B, F, T = spec.shape
        sp = np.ascontiguousarray(np.stack([spec.real, spec.imag], -1), dtype=np.float32)
        mg = ph = None
    else:
        B, F, T = mag.shape
        sp = None
        mg = np.ascontiguousarray(mag, dtype=np.float32)
        ph = None if phase is None else np.ascontiguousarray(phase, dtype=np.float32)
    out_len = len(inv_wss)
    out = np.zeros((B, out_len), np.float32)
    fl = None if db_flags is None else np.ascontiguousarray(db_flags, dtype=np.int32)
    rc = lib().emul_istft512(_f(sp), _f(mg), _f(ph), mag_domain, _i(fl), B, T, int(length), hop, int(center), int(win_length),
                             _f(window), _f(inv_wss), _f(out), C.c_longlong(out_len),
                             _f(None if blend_in is None else np.ascontiguousarray(blend_in, dtype=np.float32)),
                             _f(None if blend_mask is None else np.ascontiguousarray(blend_mask, dtype=np.float32)),
                             _f(peaks))
    assert rc == 0, rc
    return out
