"""Restatement of the third-party ``librosa`` (>= 0.10) routines the reference calls.

TEST INFRASTRUCTURE ONLY -- see ``oracle/__init__.py``.  Parity unpinned against the
reference's own outputs (librosa is absent here); pinned against torch.stft/istft.

librosa is not under /root/reference; what is restated here is its published
algorithm (numpy + ``scipy.fft``, which is the very backend librosa calls), anchored
on the reference's call sites:

  librosa.stft            <- utils.py:225-232
  librosa.istft           <- utils.py:317-318, utils.py:326-327
  librosa.griffinlim      <- utils.py:330-332, utils.py:386-392
  librosa.db_to_amplitude <- utils.py:314
  librosa.util.normalize  <- utils.py:84
  librosa.time_to_frames  <- models/CNNBLSTM/dataset.py:116-117, models/model_eval.py:148-149
  librosa.feature.melspectrogram <- utils.py:268-277
  librosa.filters.mel     <- utils.py:367-373

dtype flow is kept as librosa has it: the forward window product is float64 (float64
window x float32 frames) so the forward FFT runs in double and is rounded to
complex64 for float32 input; the inverse FFT of a complex64 matrix runs in single
precision inside scipy.fft, is multiplied by the float64 window and accumulated
into a float32 buffer; the window sum-of-squares is accumulated in the output dtype.
"""
from __future__ import annotations

import numpy as np
import scipy.fft
import scipy.signal

__all__ = [
    "get_window", "pad_center", "fft_window", "tiny", "stft", "istft",
    "window_sumsquare", "griffinlim", "time_to_samples", "samples_to_frames",
    "time_to_frames", "db_to_power", "db_to_amplitude", "normalize",
    "phasor", "n_frames_for", "hz_to_mel", "mel_to_hz", "mel_frequencies", "mel", "melspectrogram",
]


def tiny(x) -> float:
    """librosa.util.tiny: smallest positive normal number of x's (real) dtype."""
    x = np.asarray(x)
    if np.issubdtype(x.dtype, np.floating) or np.issubdtype(x.dtype, np.complexfloating):
        dtype = x.dtype
    else:
        dtype = np.dtype(np.float32)
    return float(np.finfo(dtype).tiny)


def get_window(window, Nx: int, fftbins: bool = True) -> np.ndarray:
    """librosa.filters.get_window: names/tuples/floats go to scipy, arrays pass through."""
    if callable(window):
        return np.asarray(window(Nx))
    if isinstance(window, (str, tuple)) or np.isscalar(window):
        return scipy.signal.get_window(window, Nx, fftbins=fftbins)
    w = np.asarray(window)
    if w.shape[0] != Nx:
        raise ValueError(f"Window size mismatch: {w.shape[0]} != {Nx}")
    return w


def pad_center(data: np.ndarray, size: int) -> np.ndarray:
    """librosa.util.pad_center along the last axis (zero padding)."""
    n = data.shape[-1]
    lpad = int((size - n) // 2)
    if lpad < 0:
        raise ValueError(f"Target size ({size}) must be at least input size ({n})")
    pads = [(0, 0)] * data.ndim
    pads[-1] = (lpad, int(size - n - lpad))
    return np.pad(data, pads, mode="constant")


def fft_window(window, win_length: int, n_fft: int) -> np.ndarray:
    """The float64, centre-padded analysis/synthesis window librosa.stft/istft build."""
    return pad_center(np.asarray(get_window(window, win_length, fftbins=True), dtype=np.float64), n_fft)


def n_frames_for(n_samples: int, n_fft: int, hop_length: int, center: bool = True) -> int:
    padded = n_samples + (2 * (n_fft // 2) if center else 0)
    return 1 + (padded - n_fft) // hop_length


def stft(y, n_fft=2048, hop_length=None, win_length=None, window="hann", center=True,
         dtype=None, pad_mode="constant") -> np.ndarray:
    """librosa.stft for a 1-D signal.  Returns [1 + n_fft//2, T], T contiguous.

    Called by the reference at utils.py:225-232 (win_length defaulted to n_fft at
    utils.py:222-223).  librosa >= 0.10 pads with zeros when center=True; the
    pre-0.10 behaviour ("reflect") is reachable through ``pad_mode``.
    """
    y = np.asarray(y)
    if y.ndim != 1:
        raise ValueError("oracle stft restates the mono path only")
    if not np.issubdtype(y.dtype, np.floating):
        raise ValueError("Audio data must be floating-point")
    if not np.all(np.isfinite(y)):
        raise ValueError("Audio buffer is not finite everywhere")
    if win_length is None:
        win_length = n_fft
    if hop_length is None:
        hop_length = int(win_length // 4)
    if not (isinstance(hop_length, (int, np.integer)) and hop_length > 0):
        raise ValueError(f"hop_length={hop_length} must be a positive integer")
    w = fft_window(window, win_length, n_fft)
    if center:
        y = np.pad(y, int(n_fft // 2), mode=pad_mode)
    elif n_fft > y.shape[-1]:
        raise ValueError(f"n_fft={n_fft} is too large for uncentered analysis of input signal of length={y.shape[-1]}")
    if y.shape[-1] < n_fft:
        raise ValueError("signal too short")
    n_frames = 1 + (y.shape[-1] - n_fft) // hop_length
    idx = np.arange(n_fft)[:, None] + hop_length * np.arange(n_frames)[None, :]
    frames = y[idx]                                     # [n_fft, T]
    if dtype is None:
        dtype = np.complex64 if y.dtype == np.float32 else np.complex128
    out = np.empty((1 + n_fft // 2, n_frames), dtype=dtype)
    out[...] = scipy.fft.rfft(w[:, None] * frames, axis=0)   # float64 product -> double FFT
    return out


def window_sumsquare(window, n_frames, hop_length=512, win_length=None, n_fft=2048,
                     dtype=np.float32) -> np.ndarray:
    """librosa.filters.window_sumsquare (norm=None): accumulated in ``dtype``."""
    if win_length is None:
        win_length = n_fft
    n = n_fft + hop_length * (n_frames - 1)
    x = np.zeros(n, dtype=dtype)
    win_sq = np.asarray(get_window(window, win_length, fftbins=True)) ** 2
    win_sq = pad_center(win_sq, n_fft)
    for i in range(n_frames):
        s = i * hop_length
        x[s:min(n, s + n_fft)] += win_sq[:max(0, min(n_fft, n - s))]
    return x


def _overlap_add(y: np.ndarray, ytmp: np.ndarray, hop_length: int) -> None:
    n_fft = ytmp.shape[0]
    N = n_fft
    for frame in range(ytmp.shape[1]):
        sample = frame * hop_length
        if N > y.shape[-1] - sample:
            N = y.shape[-1] - sample
        if N <= 0:
            break
        y[sample:sample + N] += ytmp[:N, frame]


def istft(stft_matrix, hop_length=None, win_length=None, n_fft=None, window="hann",
          center=True, dtype=None, length=None) -> np.ndarray:
    """librosa.istft for one [F, T] matrix (called at utils.py:317-318, :326-327).

    Output length hop*(T-1) when center=True and no ``length`` (the reference never
    passes ``length``).  scipy.fft.irfft ignores imag(DC) and imag(Nyquist).
    """
    S = np.asarray(stft_matrix)
    if S.ndim != 2:
        raise ValueError("oracle istft restates the mono path only")
    if n_fft is None:
        n_fft = 2 * (S.shape[-2] - 1)
    if win_length is None:
        win_length = n_fft
    if hop_length is None:
        hop_length = int(win_length // 4)
    w = fft_window(window, win_length, n_fft)
    if length:
        padded_length = length + 2 * (n_fft // 2) if center else length
        n_frames = min(S.shape[-1], int(np.ceil(padded_length / hop_length)))
    else:
        n_frames = S.shape[-1]
    if dtype is None:
        if S.dtype == np.complex64:
            dtype = np.float32
        elif np.issubdtype(S.dtype, np.complexfloating):
            dtype = np.float64
        else:                       # real input: librosa's dtype_c2r default is float32
            dtype = np.float32
    expected = n_fft + hop_length * (n_frames - 1)
    if length:
        expected = length
    elif center:
        expected -= 2 * (n_fft // 2)
    y = np.zeros(expected, dtype=dtype)

    if center:
        start_frame = int(np.ceil((n_fft // 2) / hop_length))
        ytmp = w[:, None] * scipy.fft.irfft(S[:, :start_frame], n=n_fft, axis=0)
        head_len = n_fft + hop_length * (start_frame - 1)
        head = np.zeros(head_len, dtype=dtype)
        _overlap_add(head, ytmp, hop_length)
        if y.shape[-1] < head_len - n_fft // 2:
            y[:] = head[n_fft // 2: y.shape[-1] + n_fft // 2]
        else:
            y[: head_len - n_fft // 2] = head[n_fft // 2:]
        offset = start_frame * hop_length - n_fft // 2
    else:
        start_frame = 0
        offset = 0

    if start_frame < n_frames:
        ytmp = w[:, None] * scipy.fft.irfft(S[:, start_frame:n_frames], n=n_fft, axis=0)
        _overlap_add(y[offset:], ytmp, hop_length)

    wss = window_sumsquare(window, n_frames, hop_length=hop_length, win_length=win_length,
                           n_fft=n_fft, dtype=dtype)
    start = n_fft // 2 if center else 0
    wss = wss[start:]
    if wss.shape[0] < y.shape[0]:
        wss = np.pad(wss, (0, y.shape[0] - wss.shape[0]))
    else:
        wss = wss[: y.shape[0]]
    nz = wss > tiny(wss)
    y[nz] /= wss[nz]
    return y


def phasor(angles) -> np.ndarray:
    """librosa.util.phasor (mag=None): cos + 1j*sin."""
    a = np.asarray(angles)
    return np.cos(a) + 1j * np.sin(a)


def griffinlim(S, n_iter=32, hop_length=None, win_length=None, n_fft=None, window="hann",
               center=True, dtype=None, length=None, pad_mode="constant", momentum=0.99,
               init="random", random_state=None, init_angles=None) -> np.ndarray:
    """librosa.griffinlim (fast Griffin-Lim, momentum 0.99) as called at utils.py:330-332.

    ``init_angles`` (not a librosa argument) injects the initial unit phasors so that
    the CUDA path can be compared value-for-value; with ``init_angles=None`` the
    librosa behaviour is kept: uniform random phases from ``default_rng(random_state)``
    (``init='random'``) or all-ones (``init=None``).
    """
    S = np.asarray(S)
    if random_state is None:
        rng = np.random.default_rng()
    elif isinstance(random_state, (int, np.integer)):
        rng = np.random.RandomState(seed=int(random_state))
    else:
        rng = random_state
    if momentum < 0:
        raise ValueError(f"griffinlim() called with momentum={momentum} < 0")
    if n_fft is None:
        n_fft = 2 * (S.shape[-2] - 1)
    if S.dtype in (np.float64, np.complex128):
        cdtype = np.complex128
    else:
        cdtype = np.complex64
    angles = np.empty(S.shape, dtype=cdtype)
    eps = tiny(angles)
    if init_angles is not None:
        angles[:] = init_angles
    elif init == "random":
        if isinstance(rng, np.random.RandomState):
            angles[:] = phasor(2 * np.pi * rng.random_sample(size=S.shape))
        else:
            angles[:] = phasor(2 * np.pi * rng.random(size=S.shape))
    elif init is None:
        angles[:] = 1.0
    else:
        raise ValueError(f"init={init} must either None or 'random'")
    tprev = None
    angles *= S
    kw = dict(hop_length=hop_length, win_length=win_length, n_fft=n_fft, window=window, center=center)
    for _ in range(n_iter):
        inverse = istft(angles, dtype=dtype, length=length, **kw)
        rebuilt = stft(inverse, pad_mode=pad_mode, **kw)
        angles[:] = rebuilt
        if tprev is not None:
            angles -= (momentum / (1 + momentum)) * tprev
        angles /= np.abs(angles) + eps
        angles *= S
        tprev = rebuilt
    return istft(angles, dtype=dtype, length=length, **kw)


def time_to_samples(times, sr=22050):
    return (np.asanyarray(times) * sr).astype(int)


def samples_to_frames(samples, hop_length=512, n_fft=None):
    offset = int(n_fft // 2) if n_fft is not None else 0
    samples = np.asanyarray(samples)
    return np.asarray(np.floor((samples - offset) // hop_length), dtype=int)


def time_to_frames(times, sr=22050, hop_length=512, n_fft=None):
    """librosa.time_to_frames: ``int(float64(t) * sr) // hop`` (truncation first).

    The truncation is why int((k/16000)*16000) == k-1 for some k;
    call sites: models/CNNBLSTM/dataset.py:116-117, models/model_eval.py:148-149.
    """
    return samples_to_frames(time_to_samples(times, sr=sr), hop_length=hop_length, n_fft=n_fft)


def db_to_power(S_db, ref=1.0):
    return ref * np.power(10.0, np.asarray(S_db) * 0.1)


def db_to_amplitude(S_db, ref=1.0):
    """librosa.db_to_amplitude = db_to_power(S_db, ref**2) ** 0.5 (utils.py:314)."""
    return db_to_power(S_db, ref=ref ** 2) ** 0.5


def normalize(S, norm=np.inf, axis=0):
    """librosa.util.normalize with norm=inf, threshold=None, fill=None (utils.py:84)."""
    S = np.asarray(S)
    if norm is None:
        return S
    if norm != np.inf:
        raise ValueError("oracle restates norm=inf only")
    threshold = tiny(S)
    mag = np.abs(S).astype(float)
    length = np.max(mag, axis=axis, keepdims=True)
    small = length < threshold
    length = np.where(small, 1.0, length)
    out = np.empty_like(S)
    out[:] = S / length
    return out


# --- mel (librosa.filters.mel / librosa.feature.melspectrogram; reference utils.py:268-277, :367-373) ---------------

def hz_to_mel(frequencies, htk=False):
    """librosa.hz_to_mel: Slaney's auditory-toolbox scale (linear below 1 kHz, log above) unless ``htk``."""
    f = np.asanyarray(frequencies)
    if htk:
        return 2595.0 * np.log10(1.0 + f / 700.0)
    f_min, f_sp = 0.0, 200.0 / 3
    mels = (f - f_min) / f_sp
    min_log_hz = 1000.0
    min_log_mel = (min_log_hz - f_min) / f_sp
    logstep = np.log(6.4) / 27.0
    if f.ndim:
        log_t = f >= min_log_hz
        mels[log_t] = min_log_mel + np.log(f[log_t] / min_log_hz) / logstep
    elif f >= min_log_hz:
        mels = min_log_mel + np.log(f / min_log_hz) / logstep
    return mels


def mel_to_hz(mels, htk=False):
    """librosa.mel_to_hz."""
    m = np.asanyarray(mels)
    if htk:
        return 700.0 * (10.0 ** (m / 2595.0) - 1.0)
    f_min, f_sp = 0.0, 200.0 / 3
    freqs = f_min + f_sp * m
    min_log_hz = 1000.0
    min_log_mel = (min_log_hz - f_min) / f_sp
    logstep = np.log(6.4) / 27.0
    if m.ndim:
        log_t = m >= min_log_mel
        freqs[log_t] = min_log_hz * np.exp(logstep * (m[log_t] - min_log_mel))
    elif m >= min_log_mel:
        freqs = min_log_hz * np.exp(logstep * (m - min_log_mel))
    return freqs


def mel_frequencies(n_mels=128, fmin=0.0, fmax=11025.0, htk=False):
    """librosa.mel_frequencies: n_mels points uniformly spaced on the mel axis."""
    return mel_to_hz(np.linspace(hz_to_mel(fmin, htk=htk), hz_to_mel(fmax, htk=htk), n_mels), htk=htk)


def mel(sr, n_fft, n_mels=128, fmin=0.0, fmax=None, htk=False, norm="slaney", dtype=np.float32):
    """librosa.filters.mel: triangular filters [n_mels, 1 + n_fft//2], area-normalised (``norm='slaney'``).

    The weight matrix is ``dtype`` (float32) from the start, as in librosa: every row is rounded to float32 when it is
    stored and the Slaney normalisation multiplies the rounded rows."""
    if fmax is None:
        fmax = float(sr) / 2
    n_mels = int(n_mels)
    weights = np.zeros((n_mels, int(1 + n_fft // 2)), dtype=dtype)
    fftfreqs = np.fft.rfftfreq(n=n_fft, d=1.0 / sr)
    mel_f = mel_frequencies(n_mels + 2, fmin=fmin, fmax=fmax, htk=htk)
    fdiff = np.diff(mel_f)
    ramps = np.subtract.outer(mel_f, fftfreqs)
    for i in range(n_mels):
        lower = -ramps[i] / fdiff[i]
        upper = ramps[i + 2] / fdiff[i + 1]
        weights[i] = np.maximum(0, np.minimum(lower, upper))
    if norm == "slaney":
        enorm = 2.0 / (mel_f[2:n_mels + 2] - mel_f[:n_mels])
        weights *= enorm[:, np.newaxis]
    elif norm is not None:
        raise ValueError("oracle restates norm='slaney' / None only")
    return weights


def melspectrogram(y=None, sr=22050, S=None, n_fft=2048, hop_length=512, win_length=None, window="hann", center=True,
                   pad_mode="constant", power=2.0, **kwargs):
    """librosa.feature.melspectrogram: mel basis contracted with |stft(y)| ** power (utils.py:268-277)."""
    if S is None:
        S = np.abs(stft(y, n_fft=n_fft, hop_length=hop_length, win_length=win_length, window=window, center=center,
                        pad_mode=pad_mode)) ** power
    else:
        n_fft = 2 * (S.shape[-2] - 1)
    mel_basis = mel(sr=sr, n_fft=n_fft, **kwargs)
    return np.einsum("...ft,mf->...mt", S, mel_basis, optimize=True)
