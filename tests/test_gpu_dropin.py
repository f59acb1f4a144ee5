"""GPU tier: the drop-in numpy API (dropin/utils.py etc.) against the oracle's restatement of the reference's
utils.py, the host-buffer pipeline, the bulk preprocess path, and full-size property tests."""
import os
import sys
from pathlib import Path

import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

ROOT = Path(__file__).resolve().parents[1]
DROPIN = ROOT / "ml_audio_inpainting_b200" / "dropin"

from oracle import callers_port as cp      # noqa: E402
from oracle import librosa_port as lr      # noqa: E402
from oracle import utils_port as up        # noqa: E402

TOL = 1e-4


def relerr(a, b):
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))


@pytest.fixture(scope="module")
def utils(tmp_path_factory):
    os.environ["AIP_OUTPUT_DIR"] = str(tmp_path_factory.mktemp("out"))
    sys.path.insert(0, str(DROPIN))
    for m in ("utils", "config", "add_gaps"):
        sys.modules.pop(m, None)
    import utils as u
    yield u
    sys.path.remove(str(DROPIN))
    for m in ("utils", "config", "add_gaps"):
        sys.modules.pop(m, None)


@pytest.fixture(scope="module")
def flac_file(tmp_path_factory, golden_clips):
    from ml_audio_inpainting_b200 import audio_io
    name = sorted(golden_clips)[1]
    path = tmp_path_factory.mktemp("audio") / f"{name}.flac"
    x = np.concatenate([golden_clips[name], golden_clips[name][:20000]])      # 6.25 s: load_audio must truncate
    audio_io.write_audio(path, x, 16000, "flac")
    return path, golden_clips[name]


def test_config_constants(utils):
    import config
    assert (config.DEFAULT_SAMPLE_RATE, config.DEFAULT_N_FFT, config.DEFAULT_HANN_WINDOW_SIZE,
            config.DEFAULT_HANN_HOP_LENGTH) == (16000, 512, 384, 192)
    assert config.DEFAULT_GAP_START_TIME == 2.0 and config.DEFAULT_GAP_DURATION == 0.5
    assert config.SUPPORTED_FORMATS == [".flac", ".wav", ".mp3"] and Path(config.OUTPUT_DIR).is_dir()


def test_load_audio(utils, flac_file):
    path, clip = flac_file
    a, sr = utils.load_audio(path)
    assert sr == 16000 and a.shape == (80000,) and a.dtype == np.float32
    assert np.abs(a - clip).max() <= 1.0 / 32767 + 1e-7          # PCM-16 round trip of the fixture
    b, _ = utils.load_audio(path, max_len=8)
    assert b.shape == (128000,) and np.all(b[100000:] == 0)
    with pytest.raises(IOError):
        utils.load_audio(path.parent / "nope.flac")


def test_create_gap_mask_matches_reference_semantics(utils):
    for args in [(80000, 0.2, 16000, None), (80000, 0.08, 16000, 2.0), (1000, 0.0, 16000, None), (1000, 1.0, 16000, None)]:
        np.random.seed(11)
        m, iv = utils.create_gap_mask(*args)
        np.random.seed(11)
        rm, riv = up.create_gap_mask(*args)
        assert iv == riv and m.dtype == np.float32 and np.array_equal(m, rm)


def test_add_random_gap_and_insert_gap(utils, flac_file, tmp_path):
    path, clip = flac_file
    loaded, _ = utils.load_audio(path)
    np.random.seed(5)
    y, (t0, t1) = utils.add_random_gap(path, 0.1)
    np.random.seed(5)
    ry, (rt0, rt1) = up.add_random_gap_from_audio(loaded, 0.1)
    assert (t0, t1) == (rt0, rt1) and y.dtype == np.float64 and np.array_equal(y, ry)
    with pytest.raises(ValueError):
        utils.add_random_gap(path, 6.0)
    import add_gaps
    out = tmp_path / "gap.flac"
    add_gaps.insert_gap(path, out, 2.0, 0.5)
    z, _ = utils.load_audio(out)
    ref = up.insert_gap_from_audio(loaded, 2.0, 0.5)
    assert np.abs(z - ref.astype(np.float32)).max() <= 1.0 / 32767 + 1e-7 and np.all(z[32000:40000] == 0)


@pytest.mark.parametrize("kw", [dict(n_fft=512, hop_length=192, win_length=384), dict(n_fft=512, hop_length=128, win_length=512),
                                dict(), dict(n_fft=1024, hop_length=256, window="hamming")])
def test_extract_spectrogram(utils, golden_clips, kw):
    x = golden_clips[sorted(golden_clips)[2]]
    S = utils.extract_spectrogram(x, **kw)
    ref = up.extract_spectrogram(x, **kw)
    assert S.shape == ref.shape and S.dtype == np.complex64
    assert relerr(S, ref) < TOL
    S64 = utils.extract_spectrogram(x.astype(np.float64), **kw)
    assert S64.dtype == np.complex128
    with pytest.raises(ValueError):
        utils.extract_spectrogram(x, power=-1.0)


def test_spectrogram_to_audio_branches(utils, golden_clips):
    x = golden_clips[sorted(golden_clips)[3]]
    kw = dict(n_fft=512, hop_length=192, win_length=384)
    S = up.extract_spectrogram(x, **kw)
    y = utils.spectrogram_to_audio(S, phase_info=True, **kw)
    ref = up.spectrogram_to_audio(S, phase_info=True, **kw)
    assert y.shape == ref.shape == (79872,) and y.dtype == np.float32 and relerr(y, ref) < TOL
    y = utils.spectrogram_to_audio(np.abs(S), phase=np.angle(S), **kw)
    assert relerr(y, up.spectrogram_to_audio(np.abs(S), phase=np.angle(S), **kw)) < TOL
    db = (20 * np.log10(np.abs(S) / np.abs(S).max() * 0.5 + 1e-12)).astype(np.float32)       # max < 0, mean < 0
    y = utils.spectrogram_to_audio(db, phase=np.angle(S), **kw)
    assert relerr(y, up.spectrogram_to_audio(db, phase=np.angle(S), **kw)) < 2 * TOL
    # CNNBLSTM/train.py:181-183 style call: only n_fft given -> hop 512 / win 512 defaults are honoured
    y = utils.spectrogram_to_audio(np.abs(S), phase=np.angle(S), n_fft=512)
    assert y.shape == (512 * (S.shape[1] - 1),)
    # Griffin-Lim (random phases): the reference's statistical criterion, tests/utils_test.py:897-902
    t = np.arange(16000) / 16000
    sine = (0.5 * np.sin(2 * np.pi * 440 * t)).astype(np.float32)
    mag = np.abs(up.extract_spectrogram(sine, n_fft=512, hop_length=128, win_length=512))
    yg = utils.spectrogram_to_audio(mag, n_fft=512, n_iter=32, hop_length=128, win_length=512)
    mag2 = np.abs(up.extract_spectrogram(yg, n_fft=512, hop_length=128, win_length=512))
    T = min(mag.shape[1], mag2.shape[1])
    assert np.corrcoef(mag[:, :T].ravel(), mag2[:, :T].ravel())[0, 1] > 0.9


def test_save_audio_peak_normalises(utils, golden_clips, tmp_path):
    x = 0.25 * golden_clips[sorted(golden_clips)[4]]
    utils.save_audio(x, tmp_path / "sub" / "a.flac")
    y, _ = utils.load_audio(tmp_path / "sub" / "a.flac")
    ref = lr.normalize(x)
    assert np.abs(y - ref).max() <= 1.5 / 32767
    utils.save_audio(x, tmp_path / "b.wav", normalize=False, file_format="wav")
    z, _ = utils.load_audio(tmp_path / "b.wav")
    assert np.abs(z - x).max() <= 1.0 / 32767


def test_host_pipeline_matches_device_path():
    from ml_audio_inpainting_b200 import frontend, spectral as sp
    B, L = 37, 16000
    rng = np.random.default_rng(3)
    x = (0.1 * rng.standard_normal((B, L))).astype(np.float32)
    starts = rng.integers(0, L - 3200, B)
    gaps = np.stack([starts, starts + 3200], 1).astype(np.int32)
    plan = sp.get_plan(512, 192, 384, "hann", True, "cuda:0")
    T = plan.num_frames(L)
    h_in = torch.from_numpy(x).pin_memory()
    h_out = torch.empty((B, 257, T), dtype=torch.float32).pin_memory()
    pipe = frontend.HostPipeline(plan, B, L, chunk=8)
    pipe.logmag_gap(h_in, gaps, h_out)
    ref = sp.stft(torch.from_numpy(x).cuda(), plan, gap_samples=gaps, mag_kind=sp.MAG_LOG10_EPS, want_spec=False)["mag"]
    assert torch.equal(h_out, ref.cpu())


def test_preprocess_batch_matches_reference_loop():
    from ml_audio_inpainting_b200 import preprocess
    N, L = 9, 80000
    rng = np.random.default_rng(8)
    x = (0.3 * rng.standard_normal((N, L))).astype(np.float32)
    np.random.seed(21)
    res = preprocess.preprocess_batch(torch.from_numpy(x).cuda(), gap_len=0.1, want_logmag=True)
    np.random.seed(21)
    for b in range(N):
        ry, (t0, t1) = up.add_random_gap_from_audio(x[b], 0.1)                # pre_process_dataset.py:38
        assert np.array_equal(res["audio_gap"][b].cpu().numpy(), ry.astype(np.float32))
        assert tuple(res["gap_int_s"][b]) == (t0, t1)
        assert np.array_equal(res["audio_gap_normalized"][b].cpu().numpy(), lr.normalize(ry.astype(np.float32)))   # :41
        ref = np.abs(lr.stft(ry.astype(np.float32), n_fft=512, hop_length=192, win_length=384))
        assert relerr(10.0 ** res["logmag_gap"][b].cpu().numpy().astype(np.float64), ref + 1e-9) < TOL


def test_full_size_properties():
    """BASELINE configs[1]/[2] shapes (10 s clips): size-independent properties + sampled oracle checks."""
    from ml_audio_inpainting_b200 import spectral as sp
    B, L = 1024, 160000
    gen = torch.Generator(device="cuda").manual_seed(5)
    x = (0.1 * torch.randn((B, L), generator=gen, device="cuda")).clamp_(-1, 1)
    plan = sp.get_plan(512, 192, 384, "hann", True, "cuda:0")
    S = sp.stft(x, plan)["spec"]
    assert tuple(S.shape) == (B, 257, 834)
    # linearity: stft(a x1 + b x2) == a stft(x1) + b stft(x2)
    y = 0.5 * x[:64] - 0.25 * x[64:128]
    lin = 0.5 * S[:64] - 0.25 * S[64:128]
    assert float((sp.stft(y, plan)["spec"] - lin).abs().max() / lin.abs().max()) < TOL
    # Parseval-type checksum per clip, against a float64 evaluation of the same identity on the device
    w = torch.from_numpy(sp.fft_window("hann", 384, 512)).cuda()
    e_spec = (S.abs().double() ** 2)
    e_spec = e_spec[:, 0] + e_spec[:, 256] + 2 * e_spec[:, 1:256].sum(1)          # [B, T]: sum_k |X_k|^2 over the full spectrum
    frames = torch.nn.functional.pad(x[:8].double(), (256, 256)).unfold(1, 512, 192) * w    # [8, T, 512]
    e_time = 512 * (frames ** 2).sum(-1)
    assert float(((e_spec[:8] - e_time).abs() / e_time.clamp_min(1e-12)).max()) < 1e-4
    # round trip at full size: SNR floor
    r = sp.istft(plan, spec=S)
    n = r.shape[1]
    err = (r[:, 512:n - 512] - x[:, 512:n - 512]).double()
    snr = 10 * torch.log10((x[:, 512:n - 512].double() ** 2).sum(1) / (err ** 2).sum(1))
    assert float(snr.min()) >= 100.0
    # sampled clips against the oracle
    for b in (0, 511, 1023):
        ref = lr.stft(x[b].cpu().numpy(), n_fft=512, hop_length=192, win_length=384)
        assert relerr(S[b].cpu().numpy(), ref) < TOL
    # idempotence of the gap epilogue: zeroing the gap in the waveform first gives bit-identical output
    gaps = torch.tensor([[1000 * (b % 100), 1000 * (b % 100) + 3200] for b in range(B)], dtype=torch.int32, device="cuda")
    a = sp.stft(x, plan, gap_samples=gaps, mag_kind=sp.MAG_LOG10_EPS, want_spec=False)["mag"]
    xz = x.clone()
    for b in range(0, B, 97):
        xz[b, int(gaps[b, 0]):int(gaps[b, 1])] = 0
        assert torch.equal(a[b], sp.stft(xz[b:b + 1], plan, mag_kind=sp.MAG_LOG10_EPS, want_spec=False)["mag"][0])


def test_config5_end_to_end_cnnblstm_inference(golden_clips):
    """BASELINE configs[4] shape on one GPU: GPU front-end -> random-init model -> GPU back-end (model_eval.py:146-192)."""
    from ml_audio_inpainting_b200 import frontend, spectral as sp
    from tests.support_cnnblstm import StandInBLSTMCNN
    names = sorted(golden_clips)
    x = np.stack([golden_clips[n] for n in names])                              # 9 clips x 5 s
    xd = torch.from_numpy(x).cuda()
    torch.manual_seed(0)
    model = StandInBLSTMCNN().cuda().eval()
    ev = frontend.eval_cnnlstm_batch(xd)                                         # model_eval.py:146-154
    with torch.no_grad():
        rec = model.reconstruct_spectrogram(ev["log_impaired_magnitude"], ev["mask"])    # :157-160
    assert tuple(rec.shape) == (9, 257, 417)
    # outside the gap frames the blend returns the input bit for bit
    keep = ev["mask"] == 0
    assert torch.equal(rec[keep], ev["log_impaired_magnitude"][keep])
    y = frontend.backend_batch(rec, ev["original_phase"], mag_domain=sp.DOM_POW10).cpu().numpy()     # :163, :179-189 (10** fused)
    rec_np, ph = rec.cpu().numpy(), ev["original_phase"].cpu().numpy()
    # the same hand-off with the blend fused too (raw network output in, waveform out: one kernel)
    with torch.no_grad():
        raw = model(ev["log_impaired_magnitude"].unsqueeze(1))
    yf = frontend.cnnblstm_backend_batch(raw, ev["log_impaired_magnitude"], ev["mask"], ev["original_phase"]).cpu().numpy()
    for b in range(9):
        ref = cp.eval_backend(10.0 ** rec_np[b], ph[b])                          # oracle: 10** then spectrogram_to_audio(phase=...)
        assert y[b].shape == ref.shape == (79872,)
        assert relerr(y[b], ref) < 2 * TOL
        assert relerr(yf[b], ref) < 2 * TOL
