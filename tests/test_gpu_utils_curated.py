"""GPU tier: a curated restatement of the reference's own test-suite (tests/utils_test.py:149-1182, 36 tests) run against
the B200 drop-in ``utils``.

How the curation was done: the reference's file was executed UNCHANGED in the build container against the reference's own
``utils.py`` (tests/refshim.py: librosa := the oracle's restatement, soundfile := the package codec).  19 of its 36 tests pass
there; the other 17 fail against the reference's own code, for these reasons (they are NOT carried over as written):

  stale vs. the shipped utils.py
    test_extract_spectrogram (:260)                   asserts T <= 1 + (L - win) // hop, false for center=True (167 > 165)
    test_extract_mel_spectrogram (:309)               the same frame-count assertion
    test_extract_spectrogram_power_values (:277)      expects spec(power=2) == spec(power=1) ** 2; extract_spectrogram ignores power
    test_extract_mel_spectrogram_power_values (:327)  expects mel(p=2) == mel(p=1) ** 2: false for any filter bank (sum of squares)
    test_spectrogram_to_audio_with_phase_init (:958)  passes phase_initialization=, which spectrogram_to_audio does not accept
    test_save_audio_error_handling (:537)             expects IOError under /root/unauthorized: only when not running as root
  plotting (matplotlib; out of scope)
    test_visualize_spectrogram, _return_figure, _with_gap (:551-:610), test_end_to_end_pipeline (:725, its last step plots)
  Griffin-Lim from UNSEEDED random phases on a COMPLEX "magnitude", judged by time-domain / spectral correlation
    test_full_reconstruction_spectrogram (:624), test_griffin_lim_reconstruction_quality (:851), test_griffin_lim_convergence
    (:907), test_stft_window_effects (:1006), test_hop_length_effects (:1058), test_real_audio_file_reconstruction (:1112)
    -- extract_spectrogram returns the complex STFT and these tests hand it to spectrogram_to_audio(phase_info=False), so
    librosa multiplies random phasors by the TRUE phase; the result is uncorrelated with the input (measured with the
    reference's own utils.py on the oracle: correlation -0.09, spectral correlation 0.40 for the sine).  What they were
    meant to pin is carried in tests/test_gpu_round2.py on magnitude input (the five signals, the thresholds 0.9 / 0.7, the
    iteration sweep) and, for the complex input itself, as value parity with injected phasors.

The 19 tests that do hold are restated below in this file's own words, plus the still-meaningful half of the stale ones
(shapes with the correct frame count).  Fixtures follow the reference's (:39-105): a 2-s 440 + 880 Hz sine as a WAV file,
a stereo variant, |STFT| and mel of a 440 Hz sine at config.py's parameters.
"""
import os
import sys
from pathlib import Path

import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

from oracle import librosa_port as lr          # noqa: E402  (checker)

ROOT = Path(__file__).resolve().parents[1]
DROPIN = ROOT / "ml_audio_inpainting_b200" / "dropin"
SR, N_FFT, WIN, HOP = 16000, 512, 384, 192      # config.py:27-30


@pytest.fixture(scope="module")
def utils(tmp_path_factory):
    os.environ["AIP_OUTPUT_DIR"] = str(tmp_path_factory.mktemp("out3"))
    sys.path.insert(0, str(DROPIN))
    for m in ("utils", "config", "add_gaps"):
        sys.modules.pop(m, None)
    import utils as u
    yield u
    sys.path.remove(str(DROPIN))
    for m in ("utils", "config", "add_gaps"):
        sys.modules.pop(m, None)


def two_tone(seconds=2):
    t = np.linspace(0, seconds, int(SR * seconds))
    return 0.5 * np.sin(2 * np.pi * 440 * t) + 0.3 * np.sin(2 * np.pi * 880 * t)


@pytest.fixture(scope="module")
def wav_file(tmp_path_factory):
    from ml_audio_inpainting_b200 import audio_io
    x = two_tone()
    p = tmp_path_factory.mktemp("wav") / "tone.wav"
    audio_io.write_audio(p, x, SR, "wav")
    return p, x


@pytest.fixture(scope="module")
def stereo_file(tmp_path_factory):
    from ml_audio_inpainting_b200 import audio_io
    t = np.linspace(0, 2, 2 * SR)
    st = np.stack([0.5 * np.sin(2 * np.pi * 440 * t), 0.5 * np.sin(2 * np.pi * 880 * t)], 1)
    p = tmp_path_factory.mktemp("wav2") / "stereo.wav"
    audio_io.write_audio(p, st, SR, "wav")
    return p


# ---- load_audio (:149-212) ----
def test_load_pads_truncates_downmixes(utils, wav_file, stereo_file):
    path, _ = wav_file
    for kw, n in ((dict(max_len=5), 5 * SR), (dict(max_len=1), SR), (dict(sample_rate=16000, max_len=5), 5 * SR)):
        a, sr = utils.load_audio(path, **({"sample_rate": SR} | kw))
        assert sr == SR and a.ndim == 1 and a.shape[0] == n and a.dtype == np.float32
    a, sr = utils.load_audio(stereo_file, sample_rate=SR, max_len=5, mono=True)
    assert sr == SR and a.ndim == 1 and a.shape[0] == 5 * SR
    with pytest.raises(IOError):
        utils.load_audio("nonexistent_file.wav")


# ---- add_random_gap (:216-255) ----
def test_random_gap_interval_shape_silence_and_error(utils, wav_file):
    path, _ = wav_file
    y, iv = utils.add_random_gap(path, gap_len=0.5, sample_rate=SR)
    assert isinstance(iv, tuple) and len(iv) == 2 and np.isclose(iv[1] - iv[0], 0.5)
    base, _ = utils.load_audio(path, sample_rate=SR)
    assert y.shape == base.shape and y.dtype == np.float64
    assert np.allclose(y[int(iv[0] * SR):int(iv[1] * SR)], 0)
    with pytest.raises(ValueError):
        utils.add_random_gap(path, gap_len=10, sample_rate=SR)


# ---- extract_spectrogram / extract_mel_spectrogram (:260-358): what still holds ----
def test_spectrogram_shapes_and_power_validation(utils, wav_file):
    _, x = wav_file
    S = utils.extract_spectrogram(x, n_fft=N_FFT, hop_length=HOP, win_length=WIN)
    assert S.shape == (N_FFT // 2 + 1, 1 + len(x) // HOP) and np.iscomplexobj(S)       # center=True: 1 + L // hop frames
    assert np.array_equal(utils.extract_spectrogram(x, n_fft=N_FFT, hop_length=HOP, win_length=WIN, power=2.0), S)   # power ignored
    M = utils.extract_mel_spectrogram(x, sample_rate=SR, n_fft=N_FFT, hop_length=HOP, n_mels=128)
    assert M.shape == (128, 1 + len(x) // HOP) and np.all(M >= 0)
    with pytest.raises(ValueError):
        utils.extract_spectrogram(np.zeros(1000), power=-1.0)
    with pytest.raises(ValueError):
        utils.extract_mel_spectrogram(np.zeros(1000), power=-1.0)


# ---- spectrogram_to_audio (:361-416) ----
def test_back_end_griffinlim_and_phase_paths(utils):
    t = np.linspace(0, 2, 2 * SR)
    x = 0.5 * np.sin(2 * np.pi * 440 * t)
    mag = np.abs(lr.stft(x, n_fft=N_FFT, hop_length=HOP, win_length=WIN))
    y = utils.spectrogram_to_audio(mag, phase_info=False, hop_length=HOP, win_length=WIN, n_fft=N_FFT, n_iter=64)
    assert isinstance(y, np.ndarray) and y.ndim == 1 and len(y) > 0 and not np.allclose(y, 0)
    S = lr.stft(x, n_fft=N_FFT, hop_length=HOP, win_length=WIN)
    z = utils.spectrogram_to_audio(S, phase_info=True, hop_length=HOP, win_length=WIN)
    n = min(len(x), len(z))
    assert np.corrcoef(x[:n], z[:n])[0, 1] > 0.9


# ---- mel_spectrogram_to_audio (:420-492, :672-723) ----
def test_mel_back_end_and_mel_round_trip(utils, wav_file):
    _, x = wav_file
    t = np.linspace(0, 2, 2 * SR)
    tone = 0.5 * np.sin(2 * np.pi * 440 * t)
    for power in (2.0, 1.0):
        mel = lr.melspectrogram(y=tone, sr=SR, n_fft=N_FFT, hop_length=HOP, n_mels=128, power=power)
        y = utils.mel_spectrogram_to_audio(mel, sample_rate=SR, n_fft=N_FFT, hop_length=HOP, n_iter=32, n_mels=128, power=power)
        assert isinstance(y, np.ndarray) and y.ndim == 1 and len(y) > 0 and not np.allclose(y, 0)
    mel = utils.extract_mel_spectrogram(x, sample_rate=SR, n_fft=N_FFT, hop_length=HOP, n_mels=128, power=2.0)
    y = utils.mel_spectrogram_to_audio(mel, sample_rate=SR, n_fft=N_FFT, hop_length=HOP, n_iter=32, n_mels=128, power=2.0)
    n = min(len(x), len(y))
    a = lr.melspectrogram(y=x[:n], sr=SR, n_mels=128)
    b = lr.melspectrogram(y=np.asarray(y[:n], dtype=np.float64), sr=SR, n_mels=128)
    assert np.corrcoef(a.ravel(), b.ravel())[0, 1] > 0


# ---- save_audio (:494-535) ----
def test_save_normalises_and_creates_directories(utils, tmp_path, wav_file):
    from ml_audio_inpainting_b200 import audio_io
    _, x = wav_file
    out = tmp_path / "output_audio.wav"
    utils.save_audio(x, out, sample_rate=SR, file_format="wav")
    pcm, sr = audio_io.read_audio(out)
    assert out.exists() and sr == SR and np.max(np.abs(pcm)) <= 1.0 and np.max(np.abs(pcm)) > 0.99      # peak-normalised
    nested = tmp_path / "nested" / "directory" / "output_audio.wav"
    assert not nested.parent.exists()
    utils.save_audio(np.random.default_rng(0).random(SR), nested, sample_rate=SR)
    assert nested.parent.is_dir() and nested.exists()


# ---- visualize_spectrogram (:612-622): argument validation only (plotting is out of scope) ----
def test_visualize_rejects_bad_power(utils):
    with pytest.raises((ValueError, ImportError)):          # ImportError: matplotlib is not installed in this image
        utils.visualize_spectrogram(np.random.default_rng(0).random((100, 100)), power=3)


# ---- the two tight librosa properties (:780-849) through the drop-in, at fp32 accuracy ----
def test_round_trip_properties_through_the_drop_in(utils):
    t = np.linspace(0, 1, SR, endpoint=False)
    x = np.sin(2 * np.pi * 440 * t)
    S = utils.extract_spectrogram(x, n_fft=N_FFT, hop_length=HOP, win_length=WIN)
    y = utils.spectrogram_to_audio(S, phase_info=True, n_fft=N_FFT, hop_length=HOP, win_length=WIN)
    n = len(y) - N_FFT
    assert S.dtype == np.complex128 and y.dtype == np.float64 and np.abs(y[:n] - x[:n]).max() < 5e-6     # 1e-10 in float64; fp32 here
    S2 = np.abs(S) * np.exp(1j * np.angle(S))
    y2 = utils.spectrogram_to_audio(np.abs(S), phase=np.angle(S), n_fft=N_FFT, hop_length=HOP, win_length=WIN)
    assert np.abs(S2 - S).max() < 1e-10 and np.abs(y2[:n] - x[:n]).max() < 5e-6
