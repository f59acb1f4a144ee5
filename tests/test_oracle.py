"""CPU tier: pins the oracle (oracle/*.py, the CPU restatement of librosa + utils.py + callers).

The reference holds no golden vectors for this path and librosa cannot be installed here, so the oracle is
"parity unpinned" against the reference's own outputs; what pins it:
  * torch.stft / torch.istft on CPU (independent implementations),
  * the tight properties the reference's own tests assert (tests/utils_test.py:780-809, :811-849),
  * the length / frame facts the shipped artefacts pin, and tests/golden/anchors.json (regression).
"""
import json
from pathlib import Path

import numpy as np
import pytest
import torch

from oracle import callers_port as cp
from oracle import librosa_port as lr
from oracle import utils_port as up

GOLD = Path(__file__).parent / "golden"


def sine(seconds=2.0, sr=16000):
    t = np.arange(int(sr * seconds)) / sr
    return (0.5 * np.sin(2 * np.pi * 440 * t) + 0.3 * np.sin(2 * np.pi * 880 * t)).astype(np.float32)


@pytest.mark.parametrize("n_fft,hop,wl", [(512, 192, 384), (512, 128, 512), (2048, 512, 2048), (256, 64, 200)])
def test_stft_matches_torch(n_fft, hop, wl):
    rng = np.random.default_rng(0)
    x = rng.standard_normal(12345).astype(np.float32)
    S = lr.stft(x, n_fft=n_fft, hop_length=hop, win_length=wl)
    w = torch.from_numpy(lr.fft_window("hann", wl, n_fft).astype(np.float32))
    St = torch.stft(torch.from_numpy(x), n_fft, hop, n_fft, window=w, center=True, pad_mode="constant",
                    return_complex=True).numpy()
    assert S.shape == St.shape == (1 + n_fft // 2, 1 + len(x) // hop)
    assert S.dtype == np.complex64
    assert np.abs(S - St).max() / np.abs(S).max() < 2e-6


def test_istft_matches_torch():
    rng = np.random.default_rng(1)
    x = rng.standard_normal(16000).astype(np.float32)
    S = lr.stft(x, n_fft=512, hop_length=128, win_length=512)
    y = lr.istft(S, hop_length=128, win_length=512, n_fft=512)
    w = torch.from_numpy(lr.fft_window("hann", 512, 512).astype(np.float32))
    yt = torch.istft(torch.from_numpy(S), 512, 128, 512, window=w, center=True).numpy()
    assert y.dtype == np.float32 and y.shape == yt.shape == (128 * (S.shape[1] - 1),)
    assert np.abs(y - yt).max() / np.abs(y).max() < 2e-6


def test_reference_tight_round_trip_float64():
    """tests/utils_test.py:780-809: istft(stft(x), length=len(x)) == x to 1e-10 in float64 at P1."""
    t = np.linspace(0, 1, 16000, endpoint=False)
    x = np.sin(2 * np.pi * 440 * t)
    S = lr.stft(x, n_fft=512, hop_length=192, win_length=384, window="hann", center=True)
    assert S.dtype == np.complex128
    y = lr.istft(S, hop_length=192, win_length=384, window="hann", center=True, length=len(x))
    n = len(x) - 512        # the last partial hop is not covered by a full window sum
    np.testing.assert_allclose(y[:n], x[:n], atol=1e-10)
    # tests/utils_test.py:811-849: magnitude * exp(j * phase) recombination
    S2 = np.abs(S) * np.exp(1j * np.angle(S))
    np.testing.assert_allclose(S2, S, atol=1e-10)
    y2 = lr.istft(S2, hop_length=192, win_length=384, length=len(x))
    np.testing.assert_allclose(y2[:n], x[:n], atol=1e-10)


def test_shapes_lengths_like_the_shipped_artefacts():
    x = np.zeros(80000, np.float32)
    assert lr.stft(x, n_fft=512, hop_length=192, win_length=384).shape == (257, 417)
    assert lr.stft(x, n_fft=512, hop_length=128, win_length=512).shape == (257, 626)
    assert len(lr.istft(np.zeros((257, 417), np.complex64), hop_length=192, win_length=384)) == 79872
    assert len(lr.istft(np.zeros((257, 626), np.complex64), hop_length=128, win_length=512)) == 80000
    assert lr.time_to_frames(2.0, sr=16000, hop_length=192) == 166
    assert lr.time_to_frames(2.08, sr=16000, hop_length=192) == 173
    assert cp.gan_frame_mask_range(32000, 33280, 128, 626) == (250, 260)


def test_window_padding_and_exact_zero():
    w = lr.fft_window("hann", 384, 512)
    assert w.shape == (512,) and np.all(w[:64] == 0) and np.all(w[448:] == 0) and w[64] == 0.0
    assert abs(w[64 + 192] - 1.0) < 1e-15


def test_time_to_frames_float64_truncation_quirk():
    """int((k / 16000) * 16000) == k - 1 for 741 of the k in [0, 80000]."""
    k = np.arange(80001)
    back = ((k / 16000) * 16000).astype(int)
    assert int((back == k - 1).sum()) == 741 and int((back == k).sum()) == 80001 - 741
    naive = k // 192
    quirk = lr.time_to_frames(k / 16000, sr=16000, hop_length=192)
    diff = np.flatnonzero(naive != quirk)
    assert diff.tolist() == [64320, 64704, 65088, 65472]


def test_gap_functions_and_rng_order():
    np.random.seed(5)
    m, (s0, s1) = up.create_gap_mask(80000, 0.2, 16000)
    np.random.seed(5)
    assert s0 == np.random.randint(0, 80000 - 3200 + 1) and s1 == s0 + 3200       # inclusive upper bound
    assert m.dtype == np.float32 and m.sum() == 80000 - 3200 and np.all(m[s0:s1] == 0)
    m, iv = up.create_gap_mask(1000, 0.0, 16000)
    assert iv == (0, 0) and np.all(m == 1)
    m, iv = up.create_gap_mask(1000, 1.0, 16000)
    assert iv == (0, 1000) and np.all(m == 0)
    m, iv = up.create_gap_mask(80000, 0.08, 16000, gap_start_s=2.0)
    assert iv == (32000, 33280)
    x = np.ones(80000, np.float32)
    np.random.seed(6)
    y, (t0, t1) = up.add_random_gap_from_audio(x, 0.1)
    np.random.seed(6)
    s = np.random.randint(0, 80000 - 1600)                                         # exclusive upper bound
    assert y.dtype == np.float64 and y.shape == x.shape and t0 == s / 16000 and t1 == (s + 1600) / 16000
    assert np.all(y[s:s + 1600] == 0) and y.sum() == 80000 - 1600
    with pytest.raises(ValueError):
        up.add_random_gap_from_audio(np.ones(100, np.float32), 1.0)
    with pytest.raises(ValueError):
        up.extract_spectrogram(x, power=-1)
    # vectorised draws == sequential scalar draws
    np.random.seed(9)
    a = np.random.randint(0, 1000, size=50)
    np.random.seed(9)
    b = np.array([np.random.randint(0, 1000) for _ in range(50)])
    assert np.array_equal(a, b)


def test_load_audio_pad_truncate():
    a, sr = up.load_audio_from_decoded(np.ones(100000, np.float32))
    assert a.shape == (80000,) and sr == 16000
    a, _ = up.load_audio_from_decoded(np.ones(1000, np.float32))
    assert a.shape == (80000,) and a[:1000].sum() == 1000 and a[1000:].sum() == 0


def test_spectrogram_to_audio_branches():
    x = sine(1.0)
    S = up.extract_spectrogram(x, n_fft=512, hop_length=192, win_length=384)
    y = up.spectrogram_to_audio(S, phase_info=True, n_fft=512, hop_length=192, win_length=384)
    n = len(y)
    assert np.corrcoef(x[:n], y)[0, 1] > 0.999
    y2 = up.spectrogram_to_audio(np.abs(S), phase=np.angle(S), n_fft=512, hop_length=192, win_length=384)
    assert np.abs(y - y2).max() < 1e-5
    db = 20 * np.log10(np.abs(S) / np.abs(S).max() * 0.5 + 1e-12)
    assert db.max() < 0
    y3 = up.spectrogram_to_audio(db, phase=np.angle(S), n_fft=512, hop_length=192, win_length=384)
    ref = lr.istft(lr.db_to_amplitude(db) * np.exp(1j * np.angle(S)), hop_length=192, win_length=384, n_fft=512)
    assert np.allclose(y3, ref)


def test_griffinlim_statistics_and_determinism():
    """tests/utils_test.py:851-956: spectral correlation > 0.9 for a sine after Griffin-Lim."""
    x = sine(0.5)
    mag = np.abs(lr.stft(x, n_fft=512, hop_length=128, win_length=512))
    y = lr.griffinlim(mag, n_iter=32, hop_length=128, win_length=512, n_fft=512, random_state=0)
    mag2 = np.abs(lr.stft(y, n_fft=512, hop_length=128, win_length=512))
    T = min(mag.shape[1], mag2.shape[1])
    assert np.corrcoef(mag[:, :T].ravel(), mag2[:, :T].ravel())[0, 1] > 0.9
    ang = np.exp(2j * np.pi * np.random.default_rng(3).random(mag.shape)).astype(np.complex64)
    a = lr.griffinlim(mag, n_iter=4, hop_length=128, win_length=512, n_fft=512, init_angles=ang)
    b = lr.griffinlim(mag, n_iter=4, hop_length=128, win_length=512, n_fft=512, init_angles=ang)
    assert np.array_equal(a, b)


def test_normalize_and_db():
    y = np.array([0.5, -2.0, 1.0], np.float32)
    assert np.allclose(lr.normalize(y), y / 2.0)
    assert np.array_equal(lr.normalize(np.zeros(4, np.float32)), np.zeros(4, np.float32))
    assert np.allclose(lr.db_to_amplitude(np.array([-20.0, 0.0])), [0.1, 1.0])


def test_golden_anchors(golden_clips):
    """Regression anchors generated by tests/golden/make_golden.py from the 9 reference clips."""
    anchors = json.loads((GOLD / "anchors.json").read_text())
    assert sorted(anchors) == sorted(golden_clips) and len(anchors) == 9
    for name in sorted(golden_clips)[:3]:
        x, a = golden_clips[name], anchors[name]
        assert x.shape == (80000,) and x.dtype == np.float32
        S = lr.stft(x, n_fft=512, hop_length=192, win_length=384)
        assert list(S.shape) == a["shape"]
        assert abs(float(np.abs(S).max()) - a["max_abs_S"]) < 1e-6 * a["max_abs_S"]
        assert abs(float(np.abs(S).astype(np.float64).sum()) - a["sum_abs_S"]) < 1e-9 * a["sum_abs_S"]
        ev = cp.eval_frontend_cnnlstm(x)
        assert list(ev["gap_frames"]) == a["cnnlstm_gap_frames"]
        assert float(ev["log_impaired_magnitude"].min()) == a["min_log10_specgap"] == -9.0
        y = lr.istft(S, hop_length=192, win_length=384, n_fft=512)
        assert len(y) == a["istft_len"]
        n = len(y)
        snr = 10 * np.log10((x[:n].astype(np.float64) ** 2)[512:n - 512].sum()
                            / ((x[:n] - y).astype(np.float64) ** 2)[512:n - 512].sum())
        assert abs(snr - a["roundtrip_snr_db"]) < 0.5 and snr > 135


def test_caller_epilogues_shapes():
    x = np.random.default_rng(0).standard_normal(80000).astype(np.float32) * 0.1
    np.random.seed(1)
    it = cp.cnnblstm_item(x)
    assert it["spectrogram_gap"].shape == (257, 417) and it["spectrogram_gap"].dtype == np.float32
    assert it["spectrogram_target_phase"].dtype == np.complex64 and it["gap_mask"].sum() > 0
    f0, f1 = it["gap_frames"]
    assert np.all(it["gap_mask"][:, f0:f1] == 1) and it["gap_mask"].sum() == 257 * (f1 - f0)
    g = cp.gan_item(x)
    assert g["original_magnitude"].shape == (257, 626) and g["mask"].min() == 0 and g["mask"].max() == 1


@pytest.mark.parametrize("sr,n_fft,n_mels,fmin,fmax", [(16000, 2048, 128, 0.0, None), (16000, 512, 128, 0.0, None),
                                                       (16000, 512, 64, 50.0, 7000.0), (22050, 1024, 80, 0.0, None)])
def test_mel_matches_torchaudio(sr, n_fft, n_mels, fmin, fmax):
    """librosa.filters.mel restated (Slaney scale, norm='slaney') against torchaudio's independent implementation, and
    librosa.feature.melspectrogram as the contraction of that basis with |stft| ** power (utils.py:268-277)."""
    torchaudio = pytest.importorskip("torchaudio")
    m = lr.mel(sr=sr, n_fft=n_fft, n_mels=n_mels, fmin=fmin, fmax=fmax)
    t = torchaudio.functional.melscale_fbanks(n_fft // 2 + 1, fmin, fmax if fmax else sr / 2, n_mels, sr, norm="slaney",
                                              mel_scale="slaney").T.numpy()
    assert m.dtype == np.float32 and m.shape == t.shape == (n_mels, n_fft // 2 + 1)
    assert np.abs(m - t).max() / np.abs(m).max() < 2e-5
    x = sine(0.5)
    S = np.abs(lr.stft(x, n_fft=n_fft, hop_length=n_fft // 4)) ** 2.0
    M = lr.melspectrogram(y=x, sr=sr, n_fft=n_fft, hop_length=n_fft // 4, n_mels=n_mels, fmin=fmin, fmax=fmax, power=2.0)
    assert M.shape == (n_mels, S.shape[1]) and np.allclose(M, m @ S, rtol=1e-5, atol=1e-9 * S.max())
    # the product builds the same table (two separately written forms of the same published algorithm)
    from ml_audio_inpainting_b200.spectral import mel_basis
    assert np.array_equal(mel_basis(sr, n_fft, n_mels, fmin, fmax), m)
