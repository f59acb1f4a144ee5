"""Restatement of the reference's ``utils.py`` / ``add_gaps.py`` hot-path functions.

TEST INFRASTRUCTURE ONLY -- see ``oracle/__init__.py`` ("parity unpinned": the
reference's utils.py imports librosa/soundfile/matplotlib at module top, utils.py:2-6,
none installable here, so the reference itself cannot be imported).

File decoding is host I/O outside the path: every function that takes a path in the
reference takes the already-decoded float32 samples here (``decoded``), i.e. what
``librosa.load(path, sr=sample_rate, mono=True)`` yields for a file whose sample rate
already equals ``sample_rate`` (int16 / 32768 -> float32, mean over channels).
"""
from __future__ import annotations

import numpy as np

from . import librosa_port as lr

DEFAULT_SAMPLE_RATE = 16000          # config.py:27


def load_audio_from_decoded(decoded, sample_rate=DEFAULT_SAMPLE_RATE, max_len=5):
    """utils.py:38-50: truncate to int(sr*max_len) samples or right-pad with zeros."""
    audio = np.asarray(decoded)
    max_samples = int(sample_rate * max_len)                    # utils.py:39
    if len(audio) > max_samples:                                # utils.py:40-41
        audio = audio[:max_samples]
    else:                                                       # utils.py:42-48
        audio = np.pad(audio, (0, max_samples - len(audio)), "constant")
    return audio, sample_rate


def create_gap_mask(audio_len_samples, gap_len_s, sample_rate=DEFAULT_SAMPLE_RATE, gap_start_s=None):
    """utils.py:93-144.  Consumes the GLOBAL np.random stream exactly like the reference."""
    gap_len_samples = int(gap_len_s * sample_rate)              # utils.py:120
    if gap_len_samples <= 0:                                    # utils.py:122-124
        return np.ones(audio_len_samples, dtype=np.float32), (0, 0)
    if gap_len_samples >= audio_len_samples:                    # utils.py:126-129
        print(f"Warning: Gap length ({gap_len_s}s) >= audio length. Returning all zeros mask.")
        return np.zeros(audio_len_samples, dtype=np.float32), (0, audio_len_samples)
    max_start = audio_len_samples - gap_len_samples             # utils.py:132
    if gap_start_s is None:
        start = np.random.randint(0, max_start + 1)             # utils.py:134 (inclusive of max_start)
    else:
        start = int(gap_start_s * sample_rate)                  # utils.py:136
    end = start + gap_len_samples                               # utils.py:138
    mask = np.ones(audio_len_samples, dtype=np.float32)         # utils.py:141-142
    mask[start:end] = 0.0
    return mask, (start, end)


def add_random_gap_from_audio(audio_data, gap_len, sample_rate=DEFAULT_SAMPLE_RATE):
    """utils.py:168-188 after the load: returns (float64 audio with a zeroed range, (t0, t1) seconds)."""
    audio_data = np.asarray(audio_data)
    gap_length = int(gap_len * sample_rate)                     # utils.py:171
    audio_len = len(audio_data)
    if gap_length >= audio_len:                                 # utils.py:175-176
        raise ValueError(f"Gap length ({gap_length}s) exceeds audio length ({audio_len/sample_rate}s)")
    start = np.random.randint(0, audio_len - int(gap_len * sample_rate))   # utils.py:179 (exclusive)
    silence = np.zeros(gap_length)                              # utils.py:180 (float64!)
    audio_new = np.concatenate([audio_data[:start], silence, audio_data[start + gap_length:]])
    interval = (start / sample_rate, (start + gap_length) / sample_rate)   # utils.py:186
    return audio_new, interval


def insert_gap_from_audio(y, gap_start, gap_duration, sample_rate=16000):
    """add_gaps.py:24-32 after the load (the file write is host I/O)."""
    y = np.asarray(y)
    gap_start_idx = int(gap_start * sample_rate)
    gap_length = int(gap_duration * sample_rate)
    silence = np.zeros(gap_length)
    return np.concatenate([y[:gap_start_idx], silence, y[gap_start_idx + gap_length:]])


def extract_spectrogram(audio_data, n_fft=2048, hop_length=512, win_length=None, window="hann",
                        center=True, power=1.0):
    """utils.py:192-234: validates ``power`` then returns the COMPLEX STFT."""
    if power < 0:
        raise ValueError("Power must be non-negative")
    if win_length is None:
        win_length = n_fft
    return lr.stft(audio_data, n_fft=n_fft, hop_length=hop_length, win_length=win_length,
                   window=window, center=center)


def spectrogram_to_audio(spectrogram, phase=None, phase_info=False, n_fft=512, n_iter=64,
                         window="hann", hop_length=512, win_length=None, center=True,
                         _gl_init_angles=None, _gl_random_state=None):
    """utils.py:279-333.  ``_gl_*`` are oracle-only hooks to make Griffin-Lim repeatable."""
    spectrogram = np.asarray(spectrogram)
    if np.max(spectrogram) < 0 and np.mean(spectrogram) < 0:    # utils.py:313-314
        spectrogram = lr.db_to_amplitude(spectrogram)
    if phase_info:                                              # utils.py:316-318
        return lr.istft(spectrogram, n_fft=n_fft, hop_length=hop_length, win_length=win_length,
                        window=window, center=center)
    if phase is not None:                                       # utils.py:321-327
        complex_spectrogram = spectrogram * np.exp(1j * phase)
        return lr.istft(complex_spectrogram, n_fft=n_fft, hop_length=hop_length,
                        win_length=win_length, window=window, center=center)
    return lr.griffinlim(spectrogram, n_fft=n_fft, n_iter=n_iter, hop_length=hop_length,   # utils.py:330-332
                         win_length=win_length, window=window, center=center,
                         init_angles=_gl_init_angles, random_state=_gl_random_state)


def peak_normalize(audio_data):
    """The numeric part of utils.save_audio (utils.py:84): librosa.util.normalize."""
    return lr.normalize(audio_data)
