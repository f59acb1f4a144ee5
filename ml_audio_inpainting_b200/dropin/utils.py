"""Drop-in ``utils.py``: the reference's signatures (numpy in, numpy out, one clip per call), the
numeric work done by the B200 kernels.

Put this directory on ``sys.path`` (the reference's scripts do ``sys.path.append(repo_root); import
utils``) and the CNNBLSTM / GAN / model_eval scripts call it unchanged.  Every function cites the
reference lines it replaces.  There is no CPU fallback: without libaip_b200.so or a B200 the
transform functions raise.  For throughput use the batched API (``ml_audio_inpainting_b200.spectral``
/ ``.frontend``): these wrappers pay a host<->device round trip per clip.
"""
from __future__ import annotations

import sys
from pathlib import Path
from typing import Optional, Tuple, Union

import numpy as np

_REPO = Path(__file__).resolve().parents[2]
if str(_REPO) not in sys.path:
    sys.path.insert(0, str(_REPO))

from config import DEFAULT_SAMPLE_RATE                      # noqa: E402  (top-level import, like the reference)
from ml_audio_inpainting_b200 import audio_io, gaps         # noqa: E402


def _torch():
    import torch
    if not torch.cuda.is_available():
        raise RuntimeError("utils (B200 drop-in) needs a CUDA device: there is no CPU fallback")
    return torch


def _spectral():
    from ml_audio_inpainting_b200 import spectral
    return spectral


# --- Audio I/O ---------------------------------------------------------------------------------------

def load_audio(file_path: Union[str, Path], sample_rate: int = DEFAULT_SAMPLE_RATE, max_len: int = 5,
               mono: bool = True) -> Tuple[np.ndarray, int]:
    """reference utils.py:14-52.  Decode (built-in FLAC / WAV reader), mono, resample if the file rate
    differs, then truncate / right-pad to ``int(sample_rate * max_len)`` samples.  Any failure -> IOError."""
    try:
        # (only the first sample_rate * max_len samples are used: the decoder stops there when no resampling is needed)
        pcm, sr = audio_io.read_audio(file_path, max_samples=int(sample_rate * max_len), only_at_rate=sample_rate)
        audio_data = pcm.mean(axis=1, dtype=np.float32) if (mono and pcm.ndim == 2) else pcm.T.squeeze()
        if sr != sample_rate:
            import math
            import scipy.signal
            g = math.gcd(int(sr), int(sample_rate))
            audio_data = scipy.signal.resample_poly(audio_data, sample_rate // g, sr // g).astype(np.float32)
            sr = sample_rate
        max_samples = int(sample_rate * max_len)                        # utils.py:39
        if len(audio_data) > max_samples:
            audio_data = audio_data[:max_samples]
        else:
            audio_data = np.pad(audio_data, (0, max_samples - len(audio_data)), "constant")
        return audio_data, sr
    except Exception as e:
        raise IOError(f"Error loading audio file {file_path}: {str(e)}")


def _peak_normalize(audio_data: np.ndarray) -> np.ndarray:
    """librosa.util.normalize (norm=inf) on the device (aip_peak_normalize_f32)."""
    import ctypes as C
    torch = _torch()
    from ml_audio_inpainting_b200 import _cabi
    x = torch.from_numpy(np.ascontiguousarray(audio_data, dtype=np.float32)).cuda().reshape(1, -1)
    out = torch.empty_like(x)
    peaks = torch.empty(1, dtype=torch.float32, device=x.device)
    L = x.shape[1]
    _cabi.check(_cabi.load().aip_peak_normalize_f32(x.data_ptr(), L, out.data_ptr(), L, 1, L, peaks.data_ptr(),
                                                    C.c_void_p(torch.cuda.current_stream().cuda_stream)),
                "aip_peak_normalize_f32")
    return out[0].cpu().numpy().astype(audio_data.dtype if audio_data.dtype.kind == "f" else np.float32)


def save_audio(audio_data: np.ndarray, file_path: Union[str, Path], sample_rate: int = DEFAULT_SAMPLE_RATE,
               normalize: bool = True, file_format: str = "flac") -> None:
    """reference utils.py:54-89: mkdir, optional peak normalisation, write (PCM-16 FLAC / WAV)."""
    output_dir = Path(file_path).parent
    if output_dir and not output_dir.exists():
        try:
            output_dir.mkdir(parents=True, exist_ok=True)
        except Exception as e:
            raise IOError(f"Error creating directory {output_dir}: {str(e)}")
    audio_data = _peak_normalize(np.asarray(audio_data)) if normalize else audio_data
    try:
        audio_io.write_audio(file_path, np.asarray(audio_data), sample_rate, file_format)
    except Exception as e:
        raise IOError(f"Error saving audio to {file_path}: {str(e)}")


# --- Gap processing ----------------------------------------------------------------------------------

def create_gap_mask(audio_len_samples: int, gap_len_s: float, sample_rate: int = DEFAULT_SAMPLE_RATE,
                    gap_start_s: Optional[float] = None) -> Tuple[np.ndarray, Tuple[int, int]]:
    """reference utils.py:93-144.  Same np.random consumption; the dense mask is written by aip_gap_mask_f32."""
    import ctypes as C
    start, end, kind = gaps.gap_mask_interval(audio_len_samples, gap_len_s, sample_rate, gap_start_s)
    if kind == "all":
        print(f"Warning: Gap length ({gap_len_s}s) >= audio length. Returning all zeros mask.")
    torch = _torch()
    from ml_audio_inpainting_b200 import _cabi
    mask = torch.empty((1, audio_len_samples), dtype=torch.float32, device="cuda")
    g = torch.tensor([[start, end]], dtype=torch.int32, device="cuda")
    _cabi.check(_cabi.load().aip_gap_mask_f32(mask.data_ptr(), audio_len_samples, 1, audio_len_samples, g.data_ptr(),
                                              C.c_void_p(torch.cuda.current_stream().cuda_stream)), "aip_gap_mask_f32")
    return mask[0].cpu().numpy(), (start, end)


def _zero_range(audio: np.ndarray, start: int, length: int) -> np.ndarray:
    """[start, start+length) := 0 on the device; returns float64 like the reference's concatenate with np.zeros."""
    import ctypes as C
    torch = _torch()
    from ml_audio_inpainting_b200 import _cabi
    L = len(audio)
    x = torch.from_numpy(np.ascontiguousarray(audio, dtype=np.float32)).cuda().reshape(1, L)
    g = torch.tensor([[start, min(L, start + length)]], dtype=torch.int32, device="cuda")
    _cabi.check(_cabi.load().aip_gap_zero_f32(x.data_ptr(), L, x.data_ptr(), L, 1, L, g.data_ptr(),
                                              C.c_void_p(torch.cuda.current_stream().cuda_stream)), "aip_gap_zero_f32")
    return x[0].cpu().numpy().astype(np.float64)


def add_random_gap(file_path: Union[str, Path], gap_len: float, sample_rate: int = DEFAULT_SAMPLE_RATE,
                   mono: bool = True) -> Tuple[np.ndarray, Tuple[float, float]]:
    """reference utils.py:146-188: reload the file, zero a uniformly drawn range (exclusive upper bound),
    return (float64 audio, (start_s, end_s))."""
    audio_data, sr = load_audio(file_path, sample_rate=sample_rate, mono=mono)
    gap_length = int(gap_len * sample_rate)                             # utils.py:171
    audio_len = len(audio_data)
    if gap_length >= audio_len:                                         # utils.py:175-176
        raise ValueError(f"Gap length ({gap_length}s) exceeds audio length ({audio_len / sample_rate}s)")
    gap_start_idx = int(np.random.randint(0, audio_len - int(gap_len * sample_rate)))   # utils.py:179
    audio_new = _zero_range(audio_data, gap_start_idx, gap_length)
    return audio_new, (gap_start_idx / sample_rate, (gap_start_idx + gap_length) / sample_rate)


# --- Spectrograms ------------------------------------------------------------------------------------

def extract_spectrogram(audio_data: np.ndarray, n_fft: int = 2048, hop_length: int = 512,
                        win_length: Optional[int] = None, window: str = "hann", center: bool = True,
                        power: float = 1.0) -> np.ndarray:
    """reference utils.py:192-234: validates ``power`` and returns the COMPLEX STFT [1 + n_fft/2, T]
    (complex64 for float32 input, complex128 for float64 input -- computed in fp32 on the device)."""
    if power < 0:
        raise ValueError("Power must be non-negative")
    if win_length is None:
        win_length = n_fft
    audio_data = np.asarray(audio_data)
    if audio_data.ndim != 1:
        raise ValueError("extract_spectrogram takes a mono (1-D) signal")
    if not np.issubdtype(audio_data.dtype, np.floating):
        raise ValueError("Audio data must be floating-point")
    if not np.all(np.isfinite(audio_data)):
        raise ValueError("Audio buffer is not finite everywhere")
    torch = _torch()
    sp = _spectral()
    plan = sp.get_plan(n_fft, hop_length, win_length, window, center)
    x = torch.from_numpy(np.ascontiguousarray(audio_data, dtype=np.float32)).to(plan.device)
    S = sp.stft(x, plan)["spec"].cpu().numpy()
    return S if audio_data.dtype == np.float32 else S.astype(np.complex128)


def extract_mel_spectrogram(audio_data: np.ndarray, sample_rate: int = DEFAULT_SAMPLE_RATE, n_fft: int = 2048,
                            hop_length: int = 512, n_mels: int = 128, fmin: float = 0.0,
                            fmax: Optional[float] = None, power: float = 2.0) -> np.ndarray:
    """reference utils.py:236-277 (librosa.feature.melspectrogram with its defaults: hann window of n_fft taps, centred,
    Slaney mel basis): |STFT| ** power in the forward kernel's epilogue, then the filter-bank contraction
    (``aip_mel_project_f32``)."""
    if power < 0:
        raise ValueError("Power must be non-negative")
    audio_data = np.asarray(audio_data)
    torch = _torch()
    sp = _spectral()
    plan = sp.get_plan(n_fft, hop_length, n_fft, "hann", True)
    x = torch.from_numpy(np.ascontiguousarray(audio_data, dtype=np.float32)).to(plan.device)
    mag = sp.stft(x, plan, mag_kind=sp.MAG_POW, power=float(power), want_spec=False)["mag"]
    mel = sp.mel_project(mag, sample_rate, n_fft, n_mels, fmin, fmax).cpu().numpy()
    return mel if audio_data.dtype == np.float32 else mel.astype(np.float64)


def spectrogram_to_audio(spectrogram: np.ndarray, phase: Optional[np.ndarray] = None, phase_info: bool = False,
                         n_fft: int = 512, n_iter: int = 64, window: str = "hann", hop_length: int = 512,
                         win_length: Optional[int] = None, center: bool = True) -> np.ndarray:
    """reference utils.py:279-333: dB heuristic, then complex iSTFT / magnitude * exp(j phase) iSTFT /
    Griffin-Lim (n_iter, momentum 0.99, random initial phases drawn on the device).  The result has the dtype librosa
    would return (float64 for float64 / complex128 input or a float64 phase, else float32); the arithmetic is fp32."""
    torch = _torch()
    sp = _spectral()
    spectrogram = np.asarray(spectrogram)
    plan = sp.get_plan(n_fft, hop_length, win_length, window, center)
    dev = plan.device
    is_complex = np.iscomplexobj(spectrogram)
    wide = spectrogram.dtype in (np.float64, np.complex128) or \
        (not phase_info and phase is not None and np.asarray(phase).dtype == np.float64)
    out_dtype = np.float64 if wide else np.float32
    if is_complex:
        # np.max / np.mean of a complex array compare real parts first; the dB test is meant for real input
        if np.max(spectrogram) < 0 and np.mean(spectrogram) < 0:        # utils.py:313-314
            spectrogram = np.power(10.0, 0.05 * spectrogram)
        S = torch.from_numpy(np.ascontiguousarray(spectrogram, dtype=np.complex64)).to(dev)
        if phase_info:                                                  # utils.py:316-318
            return sp.istft(plan, spec=S).cpu().numpy().astype(out_dtype, copy=False)
        if phase is not None:                                           # utils.py:321-327
            S = S * torch.from_numpy(np.exp(1j * np.asarray(phase)).astype(np.complex64)).to(dev)
            return sp.istft(plan, spec=S).cpu().numpy().astype(out_dtype, copy=False)
        # utils.py:328-332 on a complex "magnitude" (tests/utils_test.py:624-645): librosa multiplies the phasors by it as it is
        return sp.griffinlim(plan, S, n_iter=n_iter).cpu().numpy().astype(out_dtype, copy=False)
    mag = torch.from_numpy(np.ascontiguousarray(spectrogram, dtype=np.float32)).to(dev)
    flags = sp.db_heuristic(mag.reshape(1, -1))                         # utils.py:313-314, on the device
    if phase_info:
        # librosa.istft of a real matrix: zero imaginary part
        return sp.istft(plan, mag=mag, phase=None, db_auto=True).cpu().numpy().astype(out_dtype, copy=False)
    if phase is not None:
        ph = torch.from_numpy(np.ascontiguousarray(phase, dtype=np.float32)).to(dev)
        return sp.istft(plan, mag=mag, phase=ph, db_auto=True).cpu().numpy().astype(out_dtype, copy=False)
    if int(flags.item()):
        mag = torch.pow(10.0, 0.05 * mag)
    return sp.griffinlim(plan, mag, n_iter=n_iter).cpu().numpy().astype(out_dtype, copy=False)       # utils.py:328-332


def mel_spectrogram_to_audio(mel_spectrogram: np.ndarray, sample_rate: int = DEFAULT_SAMPLE_RATE,
                             n_fft: int = 2048, hop_length: int = 512, n_iter: int = 32, n_mels: int = 128,
                             fmin: float = 0.0, fmax: Optional[float] = None, power: float = 2.0) -> np.ndarray:
    """reference utils.py:335-393: pinv(mel basis) @ mel and the square root for power spectrograms on the device
    (``aip_mel_inverse_f32``), then Griffin-Lim (hann window of n_fft taps, centred, random initial phases).  As in the
    reference, a negative projection becomes NaN under ``sqrt`` and spreads through Griffin-Lim."""
    torch = _torch()
    sp = _spectral()
    mel_spectrogram = np.asarray(mel_spectrogram)
    plan = sp.get_plan(n_fft, hop_length, n_fft, "hann", True)
    mel = torch.from_numpy(np.ascontiguousarray(mel_spectrogram, dtype=np.float32)).to(plan.device)
    mag = sp.mel_inverse(mel, sample_rate, n_fft, n_mels, fmin, fmax, take_sqrt=(power == 2.0))
    out = sp.griffinlim(plan, mag, n_iter=n_iter).cpu().numpy()
    return out if mel_spectrogram.dtype != np.float64 else out.astype(np.float64)


def visualize_spectrogram(*args, **kwargs):
    """reference utils.py:395-478 is host-side matplotlib plotting (out of scope for the GPU path): this
    forwards to a small matplotlib implementation when matplotlib is installed."""
    try:
        import matplotlib  # noqa: F401
    except ImportError as e:
        raise ImportError("visualize_spectrogram needs matplotlib, which is not installed") from e
    from ml_audio_inpainting_b200 import plotting
    return plotting.visualize_spectrogram(*args, **kwargs)
