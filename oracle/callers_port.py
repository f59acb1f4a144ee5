"""Restatement of the per-item numpy epilogues the reference's callers wrap around utils.py.

TEST INFRASTRUCTURE ONLY -- see ``oracle/__init__.py`` (parity unpinned; see there).

  cnnblstm_item     <- models/CNNBLSTM/dataset.py:89-119  (one of the gaps_per_audio iterations)
  cnnblstm_getitem  <- models/CNNBLSTM/dataset.py:74-121  (the whole item: gaps_per_audio iterations stacked)
  gan_item          <- models/GAN/dataset.py:104-166
  eval_frontend_*   <- models/model_eval.py:61-111, :146-154
  eval_backend      <- models/model_eval.py:130-143, :179-192 (spectrogram_to_audio with phase reuse)
"""
from __future__ import annotations

import math

import numpy as np

from . import librosa_port as lr
from . import utils_port as up


def cnnblstm_item(audio_data, sample_rate=16000, max_len_s=5.0, gap_len_s=0.2,
                  n_fft=512, hop_len=192, win_len=384):
    """One iteration of the gap loop of LibriSpeechDataset.__getitem__ (dataset.py:93-119).

    ``audio_data`` is what utils.load_audio returned (float32 [sr*5]).  Draws one number
    from the global np.random stream (via add_random_gap).  Returns float32 log-magnitude
    of the gapped audio, float32 (t0, t1), float32 frame mask (1 in the gap) and the
    complex64 target, all cropped to ceil(sr*max_len/hop) frames.
    """
    n_t = math.ceil(sample_rate * max_len_s / hop_len)                       # dataset.py:89
    audio_gap, gap_int_s = up.add_random_gap_from_audio(audio_data, gap_len_s, sample_rate)   # dataset.py:98
    target = up.extract_spectrogram(audio_data, n_fft=n_fft, hop_length=hop_len, win_length=win_len)   # :102
    spec_gap = np.abs(up.extract_spectrogram(audio_gap, n_fft=n_fft, hop_length=hop_len, win_length=win_len))  # :103
    spec_gap = np.log10(spec_gap + 1e-9)                                      # dataset.py:106
    target = np.asarray(target[:, :n_t], dtype=np.complex64)                  # dataset.py:110
    spec_gap = np.asarray(spec_gap[:, :n_t], dtype=np.float32)                # dataset.py:111
    gap_int = np.asarray(gap_int_s, dtype=np.float32)                         # dataset.py:112
    mask = np.zeros_like(spec_gap, dtype=np.float32)                          # dataset.py:115
    f0 = int(lr.time_to_frames(gap_int_s[0], sr=sample_rate, hop_length=hop_len))   # dataset.py:116
    f1 = int(lr.time_to_frames(gap_int_s[1], sr=sample_rate, hop_length=hop_len))   # dataset.py:117
    mask[:, f0:f1] = 1                                                        # dataset.py:118
    return dict(spectrogram_gap=spec_gap, gap_int_s=gap_int, gap_mask=mask,
                spectrogram_target_phase=target, gap_frames=(f0, f1), gap_int_s64=gap_int_s)


def cnnblstm_getitem(audio_data, gaps_per_audio=25, sample_rate=16000, max_len_s=5.0, gap_len_s=0.2,
                     n_fft=512, hop_len=192, win_len=384):
    """LibriSpeechDataset.__getitem__ (dataset.py:74-121) for one decoded file: ``gaps_per_audio`` iterations of the
    loop body above stacked into the four pre-allocated arrays (dataset.py:88-91).  Consumes ``gaps_per_audio`` draws
    of the global np.random stream, in order."""
    items = [cnnblstm_item(audio_data, sample_rate, max_len_s, gap_len_s, n_fft, hop_len, win_len)
             for _ in range(gaps_per_audio)]                                               # dataset.py:93
    return dict(spectrogram_gaps=np.stack([it["spectrogram_gap"] for it in items]),       # dataset.py:111
                gap_ints=np.stack([it["gap_int_s"] for it in items]),                     # dataset.py:112
                gap_masks=np.stack([it["gap_mask"] for it in items]),                     # dataset.py:119
                spectrogram_target_phases=np.stack([it["spectrogram_target_phase"] for it in items]),   # dataset.py:110
                gap_frames=np.array([it["gap_frames"] for it in items]))


def gan_frame_mask_range(gap_start_sample, gap_end_sample, hop_length, num_frames):
    """models/GAN/dataset.py:138-147: [s0//hop, ceil(s1/hop)) clamped to [0, T]."""
    f0 = gap_start_sample // hop_length
    f1 = int(np.ceil(gap_end_sample / hop_length))
    return max(0, f0), min(num_frames, f1)


def gan_item(original_audio, sample_rate=16000, gap_len_s=0.2, n_fft=512, hop_length=128,
             win_length=512, window="hann", power=1.0, spec_normalize=True, gap_start_s=None):
    """SpeechInpaintingDataset.__getitem__ steps 2-5 (GAN/dataset.py:104-166), without the channel dim."""
    mask_t, (s0, s1) = up.create_gap_mask(len(original_audio), gap_len_s, sample_rate, gap_start_s)   # :104-108
    impaired = original_audio * mask_t                                        # :109
    S = up.extract_spectrogram(original_audio, n_fft=n_fft, hop_length=hop_length,
                               win_length=win_length, window=window, power=power)      # :112-118
    mag = np.abs(S) ** power                                                  # :121
    mag = np.log1p(mag) if spec_normalize else mag                            # :122
    phase = np.angle(S)                                                       # :123
    Si = up.extract_spectrogram(impaired, n_fft=n_fft, hop_length=hop_length,
                                win_length=win_length, window=window, power=power)     # :126-132
    imag_ = np.abs(Si)                                                        # :134
    imag_ = np.log1p(imag_) if spec_normalize else imag_                      # :135
    f0, f1 = gan_frame_mask_range(s0, s1, hop_length, mag.shape[1])           # :138-147
    spec_mask = np.ones_like(mag, dtype=np.float32)                           # :150-152
    if f1 > f0:
        spec_mask[:, f0:f1] = 0
    return dict(original_magnitude=mag.astype(np.float32), impaired_magnitude=imag_.astype(np.float32),
                mask=spec_mask, original_phase=phase.astype(np.float32),
                gap_samples=(s0, s1), gap_frames=(f0, f1))


def eval_frontend_gan(audio, sr=16000, n_fft=512, hop_length=128, win_length=512,
                      gap_len_s=0.08, gap_start_s=2.0):
    """models/model_eval.py:61-111 (GAN branch inputs)."""
    mask_t, (s0, s1) = up.create_gap_mask(len(audio), gap_len_s, sr, gap_start_s=gap_start_s)   # :66-71
    impaired = audio * mask_t                                                 # :73
    S = up.extract_spectrogram(audio, n_fft=n_fft, hop_length=hop_length, win_length=win_length)
    mag = np.log1p(np.abs(S))                                                 # :84-85
    phase = np.angle(S)                                                       # :86
    Si = up.extract_spectrogram(impaired, n_fft=n_fft, hop_length=hop_length, win_length=win_length)
    imag_ = np.log1p(np.abs(Si))                                              # :96-97
    f0, f1 = gan_frame_mask_range(s0, s1, hop_length, mag.shape[1])           # :100-107
    spec_mask = np.ones_like(mag, dtype=np.float32)
    if f1 > f0:
        spec_mask[:, f0:f1] = 0
    return dict(original_magnitude=mag, original_phase=phase, impaired_magnitude=imag_,
                mask=spec_mask, gap_samples=(s0, s1), gap_frames=(f0, f1), original_spectrogram=S)


def eval_frontend_cnnlstm(audio, sr=16000, n_fft=512, hop_length=192, win_length=384,
                          t0=2.0, t1=2.08):
    """models/model_eval.py:146-154: spectrogram-domain gap, log10(|S*(1-mask)| + 1e-9)."""
    S = up.extract_spectrogram(audio, n_fft=n_fft, hop_length=hop_length, win_length=win_length)
    spec_mask = np.zeros_like(S, dtype=np.float32)                            # :147
    f0 = int(lr.time_to_frames(t0, sr=sr, hop_length=hop_length))             # :148
    f1 = int(lr.time_to_frames(t1, sr=sr, hop_length=hop_length))             # :149
    spec_mask[:, f0:f1] = 1                                                   # :150
    log_imp = np.log10(np.abs(S * (1 - spec_mask)) + 1e-9)                    # :154
    return dict(original_spectrogram=S, mask=spec_mask, log_impaired_magnitude=log_imp,
                original_phase=np.angle(S), gap_frames=(f0, f1))


def eval_backend(magnitude, original_phase, n_fft=512, hop_length=192, win_length=384):
    """models/model_eval.py:131-140 / :180-189: spectrogram_to_audio with the original phase."""
    return up.spectrogram_to_audio(magnitude, phase=original_phase, phase_info=False,
                                   n_fft=n_fft, hop_length=hop_length, win_length=win_length)
