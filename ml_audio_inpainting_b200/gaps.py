"""Host-side integer index arithmetic of the gap path (bit-exact by construction).

Everything here is tiny scalar work that decides WHICH samples / frames a kernel
touches; it stays on the host in float64 / int64 because the reference's results
depend on float64 truncation quirks that an integer shortcut would not reproduce:

  * gap length in samples           int(gap_len_s * sr)                 utils.py:120, :171, add_gaps.py:24-25
  * random start, create_gap_mask   randint(0, L - g + 1)  (inclusive)  utils.py:132-134
  * random start, add_random_gap    randint(0, L - g)      (exclusive)  utils.py:179
  * seconds interval                (s / sr, (s + g) / sr)              utils.py:186
  * CNNBLSTM frame range            int(float64(t) * sr) // hop         librosa.time_to_frames via
                                                                        models/CNNBLSTM/dataset.py:116-117
  * GAN frame range                 [s0 // hop, ceil(s1 / hop)) clamped models/GAN/dataset.py:138-147

``int((k / 16000) * 16000) == k - 1`` for 741 values of k in [0, 80000], so the
CNNBLSTM frame range must go through the float64 seconds exactly as the reference does.
All random draws use the GLOBAL ``np.random`` stream, in the reference's order; a
vectorised ``randint(0, hi, size=B)`` yields the same stream as B scalar draws.
"""
from __future__ import annotations

import math
from typing import Optional, Tuple

import numpy as np

__all__ = [
    "gap_len_samples", "draw_starts_inclusive", "draw_starts_exclusive", "seconds_interval",
    "time_to_frames", "cnnblstm_frame_range", "gan_frame_range", "n_frames", "istft_length",
    "gap_mask_interval",
]


def gap_len_samples(gap_len_s: float, sample_rate: int) -> int:
    """int(gap_len_s * sample_rate), float64 product truncated toward zero."""
    return int(gap_len_s * sample_rate)


def gap_mask_interval(audio_len_samples: int, gap_len_s: float, sample_rate: int,
                      gap_start_s: Optional[float] = None) -> Tuple[int, int, str]:
    """The (start, end, kind) decision of utils.create_gap_mask (utils.py:120-138).

    kind is 'none' (g <= 0 -> all ones, (0, 0)), 'all' (g >= L -> all zeros, (0, L)) or
    'gap'.  Draws from np.random only in the 'gap' case with gap_start_s None.
    """
    g = gap_len_samples(gap_len_s, sample_rate)
    if g <= 0:
        return 0, 0, "none"
    if g >= audio_len_samples:
        return 0, audio_len_samples, "all"
    max_start = audio_len_samples - g
    if gap_start_s is None:
        start = int(np.random.randint(0, max_start + 1))
    else:
        start = int(gap_start_s * sample_rate)
    return start, start + g, "gap"


def draw_starts_inclusive(audio_len_samples: int, g: int, count: int) -> np.ndarray:
    """``count`` create_gap_mask starts: randint(0, L - g + 1) (utils.py:134)."""
    return np.random.randint(0, audio_len_samples - g + 1, size=count).astype(np.int64)


def draw_starts_exclusive(audio_len_samples: int, g: int, count: int) -> np.ndarray:
    """``count`` add_random_gap starts: randint(0, L - g) (utils.py:179)."""
    return np.random.randint(0, audio_len_samples - g, size=count).astype(np.int64)


def seconds_interval(start, g: int, sample_rate: int):
    """(start / sr, (start + g) / sr) as float64 (utils.py:186); scalar or array."""
    s = np.asarray(start, dtype=np.int64)
    t0 = s.astype(np.float64) / np.float64(sample_rate)
    t1 = (s + g).astype(np.float64) / np.float64(sample_rate)
    if s.ndim == 0:
        return float(t0), float(t1)
    return t0, t1


def time_to_frames(times, sr: int = 22050, hop_length: int = 512):
    """librosa.time_to_frames: (asarray(t) * sr).astype(int) // hop_length."""
    samples = (np.asanyarray(times, dtype=np.float64) * sr).astype(np.int64)
    frames = samples // hop_length
    return int(frames) if frames.ndim == 0 else frames


def cnnblstm_frame_range(start, g: int, sample_rate: int, hop_length: int):
    """Frame range [f0, f1) the CNNBLSTM dataset marks 1 for a gap starting at sample ``start``."""
    t0, t1 = seconds_interval(start, g, sample_rate)
    return time_to_frames(t0, sample_rate, hop_length), time_to_frames(t1, sample_rate, hop_length)


def gan_frame_range(s0, s1, hop_length: int, num_frames: int):
    """[s0 // hop, ceil(s1 / hop)) clamped to [0, num_frames] (GAN/dataset.py:138-147)."""
    s0 = np.asarray(s0, dtype=np.int64)
    s1 = np.asarray(s1, dtype=np.int64)
    f0 = np.maximum(0, s0 // hop_length)
    f1 = np.minimum(num_frames, np.ceil(s1.astype(np.float64) / hop_length).astype(np.int64))
    if f0.ndim == 0:
        return int(f0), int(f1)
    return f0, f1


def n_frames(n_samples: int, n_fft: int, hop_length: int, center: bool = True) -> int:
    """Frame count of librosa.stft: 1 + (L + 2*(n_fft//2) - n_fft) // hop when centred."""
    padded = n_samples + (2 * (n_fft // 2) if center else 0)
    if padded < n_fft:
        raise ValueError(f"n_fft={n_fft} is too large for input signal of length={n_samples}")
    return 1 + (padded - n_fft) // hop_length


def istft_length(num_frames: int, n_fft: int, hop_length: int, center: bool = True,
                 length: Optional[int] = None) -> int:
    """Output length of librosa.istft."""
    if length:
        return int(length)
    n = n_fft + hop_length * (num_frames - 1)
    return n - 2 * (n_fft // 2) if center else n


def cnnblstm_crop_frames(sample_rate: int, max_len_s: float, hop_length: int) -> int:
    """ceil(sr * max_len / hop): the time dimension the CNNBLSTM dataset allocates (dataset.py:89)."""
    return math.ceil(sample_rate * max_len_s / hop_length)
