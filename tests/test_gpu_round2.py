"""GPU tier, round 2: the parity holes the round-1 review listed, each against the ORACLE (never against the kernel itself).

  * the headline kernel build stft512_fwd_kernel<log10, zero-tap pruning, T_out = 834> on gapped 10-s clips (BASELINE configs[1] shape)
  * the GAN back-end: expm1 prologue, the generator hand-off of models/model_eval.py:118-140 and of models/GAN/train.py:470-482
  * Griffin-Lim at the iteration counts the reference uses (32: utils.py:340, 64: utils.py:284) with injected phasors and the
    MEASURED drift written down, the reference's statistical criteria for its five test signals, complex "magnitudes"
  * dtype flow of spectrogram_to_audio, preprocess_tree on a file tree, mel front / back-end
  * more than 64 forward launches in flight on 8 streams (the round-1 tile-counter ring allowed two launches to share a counter)
"""
import os
import sys
from pathlib import Path

import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

from oracle import callers_port as cp          # noqa: E402  (checker)
from oracle import librosa_port as lr          # noqa: E402
from oracle import utils_port as up            # noqa: E402

ROOT = Path(__file__).resolve().parents[1]
DROPIN = ROOT / "ml_audio_inpainting_b200" / "dropin"
TOL = 1e-4
SR = 16000


def relerr(a, b):
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))


def _noise(B, L, seed=0):
    rng = np.random.default_rng(seed)
    return np.clip(0.1 * rng.standard_normal((B, L)), -1, 1).astype(np.float32)


@pytest.fixture(scope="module")
def sp():
    from ml_audio_inpainting_b200 import spectral
    return spectral


@pytest.fixture(scope="module")
def utils(tmp_path_factory):
    os.environ["AIP_OUTPUT_DIR"] = str(tmp_path_factory.mktemp("out2"))
    sys.path.insert(0, str(DROPIN))
    for m in ("utils", "config", "add_gaps"):
        sys.modules.pop(m, None)
    import utils as u
    yield u
    sys.path.remove(str(DROPIN))
    for m in ("utils", "config", "add_gaps"):
        sys.modules.pop(m, None)


# ------------------------------------------------------------------------------------------- (i) the headline build
def test_headline_kernel_build_against_the_oracle(sp):
    """BASELINE configs[1] shape: 10-s clips (L = 160 000 -> T = 834), n_fft 512 / win 384 / hop 192, gap 0.2 s, log10(|S| + 1e-9):
    the shape-specialised instantiation the benchmark times, on gapped clips, against librosa's restatement."""
    B, L, g = 6, 160000, 3200
    x = _noise(B, L, seed=834)
    starts = np.array([0, 31999, 64320, 100001, L - g - 1, L - g])           # clip start, tile borders, the float64 quirk value, clip end
    gaps = np.stack([starts, starts + g], 1)
    plan = sp.get_plan(512, 192, 384, "hann", True, "cuda:0")
    mag = sp.stft(torch.from_numpy(x).cuda(), plan, gap_samples=gaps, mag_kind=sp.MAG_LOG10_EPS, eps=1e-9, want_spec=False)["mag"]
    assert tuple(mag.shape) == (B, 257, 834)
    mag = mag.cpu().numpy()
    for b in range(B):
        xg = x[b].copy()
        xg[gaps[b, 0]:gaps[b, 1]] = 0
        ref = np.abs(lr.stft(xg, n_fft=512, hop_length=192, win_length=384)).astype(np.float64)
        assert relerr(10.0 ** mag[b].astype(np.float64), ref + 1e-9) < TOL, b
        # frames that lie wholly inside the gap see only zeros: exactly log10(1e-9)
        t_in = [t for t in range(834) if t * 192 - 256 + 64 >= gaps[b, 0] and t * 192 - 256 + 448 <= gaps[b, 1]]
        assert len(t_in) >= 14 and np.all(mag[b][:, t_in] == np.float32(-9.0))
    # the generic-T_out build of the same kernel (switch off the shape specialisation): bit-identical
    with sp.experiment_env(AIP_FWD_NO_SHAPE="1"):
        mag2 = sp.stft(torch.from_numpy(x).cuda(), plan, gap_samples=gaps, mag_kind=sp.MAG_LOG10_EPS, eps=1e-9,
                       want_spec=False)["mag"].cpu().numpy()
    assert np.array_equal(mag, mag2)


# ------------------------------------------------------------------------------------------- (ii) GAN back-end
def test_gan_backend_expm1_and_handoffs(sp, golden_clips):
    from ml_audio_inpainting_b200 import frontend
    names = sorted(golden_clips)[:3]
    x = np.stack([golden_clips[n] for n in names])
    xd = torch.from_numpy(x).cuda()
    g = frontend.eval_gan_batch(xd)                                            # models/model_eval.py:61-111
    plan = sp.get_plan(512, 128, 512, "hann", True, xd.device)
    rng = np.random.default_rng(5)
    gen = (g["original_magnitude"].cpu().numpy() + 0.05 * rng.standard_normal(g["original_magnitude"].shape)).astype(np.float32)
    gen_d = torch.from_numpy(gen).cuda()
    om, ph, mk = (g[k].cpu().numpy() for k in ("original_magnitude", "original_phase", "mask"))
    # (a) expm1 prologue: istft(expm1(log1p-magnitude) e^{j phase}) reproduces the clip
    y = sp.istft(plan, mag=g["original_magnitude"], phase=g["original_phase"], mag_domain=sp.DOM_EXPM1).cpu().numpy()
    for b in range(3):
        ref = lr.istft((np.expm1(om[b]) * np.exp(1j * ph[b])).astype(np.complex64), hop_length=128, win_length=512, n_fft=512)
        assert relerr(y[b], ref) < TOL
        assert relerr(y[b][512:-512], x[b][512:80000 - 512]) < 2 * TOL
    # (b) model_eval.py:118-140: the generator output goes to spectrogram_to_audio as it is, with the original phase
    y = frontend.backend_batch(gen_d, g["original_phase"], n_fft=512, hop_length=128, win_length=512).cpu().numpy()
    for b in range(3):
        assert relerr(y[b], cp.eval_backend(gen[b], ph[b], hop_length=128, win_length=512)) < TOL
    # (c) models/GAN/train.py:470-482: combined = generated (1 - mask) + original mask (mask 1 OUTSIDE the gap), handed over
    #     in the log1p domain as it is (DOM_LINEAR), and with the log1p undone (DOM_EXPM1); + save_audio's normalisation
    combined = gen * (1 - mk) + om * mk
    for dom, fn in ((sp.DOM_LINEAR, lambda m: m), (sp.DOM_EXPM1, np.expm1)):
        y = sp.istft_blend(plan, gen_d, g["original_magnitude"], g["mask"], g["original_phase"], mag_domain=dom,
                           mask_keeps_input=True).cpu().numpy()
        pk = torch.empty(3, device="cuda")
        yn = sp.istft_blend(plan, gen_d, g["original_magnitude"], g["mask"], g["original_phase"], mag_domain=dom,
                            mask_keeps_input=True, normalize=True, peaks_out=pk).cpu().numpy()
        for b in range(3):
            ref = up.spectrogram_to_audio(fn(combined[b]).astype(np.float32), phase=ph[b], n_fft=512, hop_length=128, win_length=512)
            assert relerr(y[b], ref) < TOL, (dom, b)
            assert relerr(yn[b], up.peak_normalize(ref)) < TOL and abs(float(pk[b]) - np.abs(ref).max()) < TOL * np.abs(ref).max()


# ------------------------------------------------------------------------------------------- (iii) Griffin-Lim
def _signals():
    """the five signals of the reference's tests (tests/utils_test.py:114-145), noise seeded"""
    t = np.linspace(0, 2, 2 * SR)
    import scipy.signal
    tt = np.arange(0, 2, 1.0 / SR)
    impulse = np.zeros(2 * SR)
    impulse[::SR // 10] = 1.0
    return {"sine": 0.5 * np.sin(2 * np.pi * 440 * t),
            "sine_combo": 0.5 * np.sin(2 * np.pi * 440 * t) + 0.3 * np.sin(2 * np.pi * 880 * t),
            "chirp": scipy.signal.chirp(tt, 20, 2, 8000, method="logarithmic", phi=-90),
            "impulse": impulse,
            "noise": np.random.default_rng(11).standard_normal(2 * SR) * 0.1}


def _spec_corr(a, b):
    A = np.abs(lr.stft(a.astype(np.float32)))           # librosa.stft defaults, as the reference's test takes them
    Bm = np.abs(lr.stft(b.astype(np.float32)))
    return float(np.corrcoef(A.ravel(), Bm.ravel())[0, 1])


# Measured on a B200 (this test prints the values): the fp32 CUDA iteration against the oracle (complex64 state, float64 FFTs)
# from IDENTICAL initial phasors, relative max-abs error of the waveform on two 1-s noise clips:
#     n_iter   0        1        2        8        32                 64
#     error    2.4e-7   7.5e-7   2.0e-6   1.3e-5   4.6e-5 / 7.2e-4    1.7e-4 / 8.4e-4
# Griffin-Lim is a fixed-point iteration without contraction: rounding differences grow with the iteration count, so beyond 8
# iterations the 1e-4 bound of the single transforms no longer holds for the waveform (stated bound: 3e-3 = 3.5 x the measured
# worst case); what stays tight is the thing the algorithm optimises -- the magnitude spectrogram of the result (spectral
# convergence) -- asserted against the oracle's own value (measured: equal to 5 digits at every count).
GL_WAVE_BOUND = {0: 1e-4, 1: 1e-4, 2: 1e-4, 8: 1e-4, 32: 3e-3, 64: 3e-3}


def test_griffinlim_value_parity_at_the_references_iteration_counts(sp):
    L = 16000
    x = _noise(2, L, seed=33)
    plan = sp.get_plan(512, 192, 384, "hann", True, "cuda:0")
    mag = sp.stft(torch.from_numpy(x).cuda(), plan, mag_kind=sp.MAG_ABS, want_spec=False)["mag"]
    m = mag.cpu().numpy()
    ang = np.exp(2j * np.pi * np.random.default_rng(1).random(m.shape)).astype(np.complex64)

    def sc(y, target):          # spectral convergence ||  |stft(y)| - target || / || target ||
        Y = np.abs(lr.stft(y, n_fft=512, hop_length=192, win_length=384))
        return float(np.linalg.norm(Y - target) / np.linalg.norm(target))

    report = {}
    for n_iter in (0, 1, 2, 8, 32, 64):
        y = sp.griffinlim(plan, mag, n_iter=n_iter, init_angles=torch.from_numpy(ang).cuda()).cpu().numpy()
        for b in range(2):
            ref = lr.griffinlim(m[b], n_iter=n_iter, hop_length=192, win_length=384, n_fft=512, init_angles=ang[b])
            e = relerr(y[b], ref)
            report[(n_iter, b)] = (e, sc(y[b], m[b]), sc(ref, m[b]))
            assert e < GL_WAVE_BOUND[n_iter], (n_iter, b, e)
            if n_iter >= 8:     # same quality as the oracle's result: spectral convergence within 3 % (relative) of its value
                assert abs(sc(y[b], m[b]) - sc(ref, m[b])) < 0.03 * sc(ref, m[b]) + 1e-3, report[(n_iter, b)]
    print("griffinlim (n_iter, clip) -> (waveform relerr, spectral convergence cuda, oracle):", report)


def test_griffinlim_statistical_criteria_five_signals(sp, utils):
    """tests/utils_test.py:851-905 on magnitude input: spectral correlation > 0.9 for sine / sine combination / chirp and
    > 0.7 for impulse train / noise after 100 iterations from random phases; :907-956: the correlation does not fall as the
    iteration count grows over 10 / 32 / 64 / 100 (measured on the spectral correlation, which is sign-invariant -- the
    reference's time-domain correlation depends on the unseeded random start)."""
    P = dict(n_fft=512, hop_length=192, win_length=384)
    for name, sig in _signals().items():
        mag = np.abs(utils.extract_spectrogram(sig.astype(np.float32), **P))
        torch.manual_seed(7)
        y = utils.spectrogram_to_audio(mag, phase_info=False, n_iter=100, **P)
        c = _spec_corr(sig[:len(y)], y)
        ref = up.spectrogram_to_audio(mag, phase_info=False, n_iter=100, _gl_random_state=7, **P)
        c_ref = _spec_corr(sig[:len(ref)], ref)
        # the reference's thresholds (0.9 tonal / 0.7 impulse, noise) where librosa's own algorithm -- the oracle, from its own
        # random start -- meets them; for white noise it does not (measured 0.44 on magnitude input: a noise spectrogram taken
        # with another n_fft is uncorrelated detail), so there the requirement is "as good as the oracle"
        want = 0.9 if name in ("sine", "sine_combo", "chirp") else 0.7
        assert c > min(want, c_ref - 0.02), (name, c, c_ref)
        assert c > c_ref - 0.05, (name, c, c_ref)          # never worse than the oracle
        print(f"griffinlim 100 it, {name}: spectral correlation cuda {c:.4f} oracle {c_ref:.4f}")
    sig = _signals()["sine_combo"]
    mag = np.abs(utils.extract_spectrogram(sig.astype(np.float32), **P))
    cs = []
    for n_iter in (10, 32, 64, 100):
        torch.manual_seed(3)
        y = utils.spectrogram_to_audio(mag, phase_info=False, n_iter=n_iter, **P)
        cs.append(_spec_corr(sig[:len(y)], y))
    assert cs[0] <= cs[-1] + 1e-3 and all(cs[i] <= cs[i + 1] + 5e-3 for i in range(3)), cs


def test_griffinlim_on_a_complex_magnitude(sp, utils):
    """tests/utils_test.py:624-645 hands extract_spectrogram's COMPLEX output to spectrogram_to_audio(phase_info=False):
    librosa multiplies the phasors by it as it is.  Value parity with injected phasors, and the call through the drop-in."""
    x = _noise(2, 8000, seed=77)
    plan = sp.get_plan(512, 192, 384, "hann", True, "cuda:0")
    S = sp.stft(torch.from_numpy(x).cuda(), plan)["spec"]
    s = S.cpu().numpy()
    ang = np.exp(2j * np.pi * np.random.default_rng(2).random(s.shape)).astype(np.complex64)
    for n_iter, tol in ((0, 1e-4), (1, 1e-4), (2, 2e-4), (4, 1e-3)):
        y = sp.griffinlim(plan, S, n_iter=n_iter, init_angles=torch.from_numpy(ang).cuda()).cpu().numpy()
        for b in range(2):
            ref = lr.griffinlim(s[b], n_iter=n_iter, hop_length=192, win_length=384, n_fft=512, init_angles=ang[b])
            assert relerr(y[b], ref) < tol, (n_iter, relerr(y[b], ref))
    y = utils.spectrogram_to_audio(s[0], phase_info=False, n_fft=512, hop_length=192, win_length=384, n_iter=8)
    assert y.dtype == np.float32 and y.shape == (192 * (s.shape[2] - 1),) and np.isfinite(y).all() and not np.allclose(y, 0)


# ------------------------------------------------------------------------------------------- (v) dtype flow
def test_spectrogram_to_audio_dtype_flow(utils):
    x = _noise(1, 6000, seed=5)[0]
    P = dict(n_fft=512, hop_length=192, win_length=384)
    S32 = utils.extract_spectrogram(x, **P)
    S64 = utils.extract_spectrogram(x.astype(np.float64), **P)
    assert S32.dtype == np.complex64 and S64.dtype == np.complex128
    cases = [(S32, None, True, np.float32), (S64, None, True, np.float64),
             (np.abs(S32), np.angle(S32), False, np.float32), (np.abs(S64), np.angle(S64), False, np.float64),
             (np.abs(S32), np.angle(S64), False, np.float64), (np.abs(S32), None, False, np.float32),
             (np.abs(S64), None, False, np.float64)]
    for spec, phase, info, want in cases:
        y = utils.spectrogram_to_audio(spec, phase=phase, phase_info=info, n_iter=2, **P)
        ref = up.spectrogram_to_audio(spec, phase=phase, phase_info=info, n_iter=2, _gl_random_state=0, **P)
        assert y.dtype == want == ref.dtype and y.shape == ref.shape
        if info or phase is not None:
            assert relerr(y, ref) < TOL


# ------------------------------------------------------------------------------------------- (vi) preprocess_tree
def test_preprocess_tree_against_the_reference_loop(tmp_path, golden_clips):
    """pre_process_dataset.py:19-43 on a small LibriSpeech-shaped tree: same os.walk order, one np.random draw per file in that
    order, gap zeroed, peak-normalised, FLAC written whatever the suffix; .mp3 skipped with a warning (no decoder)."""
    from ml_audio_inpainting_b200 import audio_io, preprocess
    names = sorted(golden_clips)[:3]
    src, dst = tmp_path / "in", tmp_path / "out"
    layout = [("84/121123", names[0] + ".flac"), ("84/121550", names[1] + ".flac"), ("174/50561", names[2] + ".wav")]
    for sub, f in layout:
        (src / sub).mkdir(parents=True, exist_ok=True)
        clip = np.concatenate([golden_clips[f.rsplit(".", 1)[0]], np.zeros(1234, np.float32)])      # longer than 5 s: truncated
        audio_io.write_audio(src / sub / f, clip, SR, f.rsplit(".", 1)[1])
    (src / "84/121123" / "notes.txt").write_text("not audio")
    (src / "84/121123" / "song.mp3").write_bytes(b"ID3")
    np.random.seed(2024)
    n = preprocess.preprocess_tree(src, dst, gap_len=0.1, supported_formats=[".flac", ".wav", ".mp3"], progress=False)
    assert n == 3
    after = np.random.randint(0, 1 << 30)
    # the reference loop, restated, in the same walk order
    np.random.seed(2024)
    seen = 0
    for root, subdirs, files in os.walk(src, topdown=True):
        if len(subdirs) == 0:
            for f in files:
                if Path(f).suffix in (".flac", ".wav"):
                    pcm, sr = audio_io.read_audio(Path(root) / f)
                    audio, _ = up.load_audio_from_decoded(pcm.reshape(-1))
                    y, _ = up.add_random_gap_from_audio(audio, 0.1)
                    want = audio_io._to_int16(up.peak_normalize(y), 32768.0)
                    out = dst / os.path.relpath(root, src) / f
                    assert out.read_bytes()[:4] == b"fLaC"                       # save_audio's default format, utils.py:59
                    got, info = audio_io.decode_flac(out.read_bytes(), verify_md5=True)
                    d = np.abs(np.asarray(got).reshape(-1).astype(np.int64) - want.astype(np.int64))
                    assert len(d) == 80000 and d.max() <= 1 and (d > 0).mean() < 2e-3
                    seen += 1
    assert seen == 3 and after == np.random.randint(0, 1 << 30)
    assert not (dst / "84/121123" / "song.mp3").exists() and not (dst / "84/121123" / "notes.txt").exists()


# ------------------------------------------------------------------------------------------- (vii) launches in flight
def test_many_forward_launches_in_flight_on_many_streams(sp):
    """200 forward launches round-robin on 8 streams, none waited for until the end: every output bit-identical to a
    serial run.  (Round 1 rotated 64 process-wide tile counters: launch k + 64 could share the counter of a running launch.)"""
    B, L = 48, 40000
    plan = sp.get_plan(512, 192, 384, "hann", True, "cuda:0")
    xs = [torch.from_numpy(_noise(B, L, seed=s)).cuda() for s in range(4)]
    want = [sp.stft(x, plan, mag_kind=sp.MAG_LOG10_EPS, want_spec=False)["mag"].clone() for x in xs]
    torch.cuda.synchronize()
    streams = [torch.cuda.Stream() for _ in range(8)]
    outs = []
    for i in range(200):
        with torch.cuda.stream(streams[i % 8]):
            outs.append((i % 4, sp.stft(xs[i % 4], plan, mag_kind=sp.MAG_LOG10_EPS, want_spec=False)["mag"]))
    torch.cuda.synchronize()
    for k, o in outs:
        assert torch.equal(o, want[k])
    # and from several host threads sharing ONE stream
    import threading
    shared = torch.cuda.Stream()
    res = [None] * 16

    def work(j):
        with torch.cuda.stream(shared):
            res[j] = sp.stft(xs[j % 4], plan, mag_kind=sp.MAG_LOG10_EPS, want_spec=False)["mag"]

    th = [threading.Thread(target=work, args=(j,)) for j in range(16)]
    [t.start() for t in th]
    [t.join() for t in th]
    torch.cuda.synchronize()
    for j in range(16):
        assert torch.equal(res[j], want[j % 4])


# ------------------------------------------------------------------------------------------- (viii) mel
@pytest.mark.parametrize("n_fft,hop,n_mels,power", [(512, 192, 128, 2.0), (2048, 512, 128, 2.0), (1024, 256, 80, 1.0)])
def test_mel_front_and_back_end(sp, utils, n_fft, hop, n_mels, power):
    """utils.py:236-277 / :335-393 against librosa.feature.melspectrogram / filters.mel restated in the oracle."""
    t = np.linspace(0, 2, 2 * SR)
    x = (0.5 * np.sin(2 * np.pi * 440 * t) + 0.1 * np.random.default_rng(4).standard_normal(len(t))).astype(np.float32)
    mel = utils.extract_mel_spectrogram(x, sample_rate=SR, n_fft=n_fft, hop_length=hop, n_mels=n_mels, power=power)
    ref = lr.melspectrogram(y=x, sr=SR, n_fft=n_fft, hop_length=hop, n_mels=n_mels, fmin=0.0, fmax=None, power=power)
    assert mel.dtype == ref.dtype == np.float32 and mel.shape == ref.shape == (n_mels, 1 + len(x) // hop)
    assert relerr(mel, ref) < TOL
    with pytest.raises(ValueError):
        utils.extract_mel_spectrogram(x, power=-1.0)
    # back-end, projection step: pinv(basis) @ mel (+ sqrt) on the device against numpy
    inv = np.linalg.pinv(lr.mel(sr=SR, n_fft=n_fft, n_mels=n_mels, fmin=0.0, fmax=None))
    lin = np.dot(inv, ref)
    got = sp.mel_inverse(torch.from_numpy(ref).cuda(), SR, n_fft, n_mels, take_sqrt=False).cpu().numpy()
    assert relerr(got, lin) < TOL
    with np.errstate(invalid="ignore"):
        want = np.sqrt(lin)
    got = sp.mel_inverse(torch.from_numpy(ref).cuda(), SR, n_fft, n_mels, take_sqrt=True).cpu().numpy()
    # where the projection is clearly positive / negative the square root / NaN must agree; |lin| ~ 0 may round to either side
    # (negative values below 1e-9 of the clip's peak are taken as 0 on the device: aip_b200.h)
    clear = np.abs(lin) > 1e-4 * np.abs(lin).max()
    assert np.array_equal(np.isnan(got)[clear], np.isnan(want)[clear])
    ok = clear & ~np.isnan(want)
    assert np.abs(got[ok] - want[ok]).max() < TOL * np.nanmax(want)
    # the whole back-end (tests/utils_test.py:420-443): a finite, non-silent waveform of librosa's length
    y = utils.mel_spectrogram_to_audio(ref, sample_rate=SR, n_fft=n_fft, hop_length=hop, n_iter=4, n_mels=n_mels, power=1.0)
    assert y.ndim == 1 and len(y) == hop * (ref.shape[1] - 1) and np.isfinite(y).all() and not np.allclose(y, 0)


# ------------------------------------------------------------------------------------------- guard regions (round-2 kernels)
@pytest.mark.parametrize("L,hop,win", [(16000, 192, 384), (16001, 192, 384), (7777, 128, 512), (700, 192, 384)])
def test_round2_kernels_write_inside_their_outputs(sp, L, hop, win):
    """compute-sanitizer is closed on this GPU pool: the outputs of the kernels added in round 2 (TMA-staged inverse for even
    and the direct one for odd T, the hand-off in both mask conventions, Griffin-Lim with staged / direct loads, mel projection
    and inverse) sit between sentinel guard regions that must survive, and every output element must have been written."""
    B, G, SENT = 3, 4096, -12345.0
    x = torch.from_numpy(_noise(B, L, seed=L)).cuda()
    plan = sp.get_plan(512, hop, win, "hann", True, "cuda:0")
    T, F = plan.num_frames(L), 257

    def guarded(shape):
        n = int(np.prod(shape))
        flat = torch.full((n + 2 * G,), SENT, dtype=torch.float32, device="cuda")
        return flat, flat[G:G + n].view(shape)

    def check(flat, view, what):
        torch.cuda.synchronize()
        ref = torch.full((G,), SENT, dtype=torch.float32, device="cuda")
        assert torch.equal(flat[:G], ref) and torch.equal(flat[-G:], ref), what
        assert not bool(view.eq(SENT).any()), what

    r = sp.stft(x, plan, mag_kind=sp.MAG_ABS, want_spec=True, want_phase=True)
    for Tn in (T, T - 1):                                   # one even, one odd frame count: staged and direct stage A
        S, mag, ph = (r[k][:, :, :Tn].contiguous() for k in ("spec", "mag", "phase"))
        n = plan.istft_length(Tn)
        flat, y = guarded((B, n))
        sp.istft(plan, spec=S, out=y)
        check(flat, y, ("istft", Tn))
        mask = torch.zeros_like(mag)
        mask[:, :, 3:6] = 1
        for keeps in (False, True):
            for dom in (sp.DOM_LINEAR, sp.DOM_POW10, sp.DOM_EXPM1):
                flat, y = guarded((B, n))
                sp.istft_blend(plan, mag * 0.5, mag * 0.1, mask, ph, mag_domain=dom, mask_keeps_input=keeps, normalize=True, out=y)
                check(flat, y, ("handoff", Tn, keeps, dom))
        y = sp.griffinlim(plan, mag, n_iter=3)              # state arrays are internal; the waveform must be finite and complete
        assert tuple(y.shape) == (B, n) and bool(torch.isfinite(y).all())
    flat, mel = guarded((B, 40, T))
    sp.mel_project(r["mag"], SR, 512, 40, out=mel)
    check(flat, mel, "mel_project")
    for sq in (False, True):
        flat, back = guarded((B, F, T))
        sp.mel_inverse(mel, SR, 512, 40, take_sqrt=sq, out=back)
        torch.cuda.synchronize()
        ref = torch.full((G,), SENT, dtype=torch.float32, device="cuda")
        assert torch.equal(flat[:G], ref) and torch.equal(flat[-G:], ref)
        assert not bool(back.eq(SENT).any())


# ------------------------------------------------------------------------------------------- 16-bit tail of save_audio
def test_wave_to_pcm16_is_save_audios_quantisation(sp):
    """aip_wave_to_pcm16_f32 = librosa.util.normalize + soundfile's float -> PCM_16 for FLAC (utils.py:83-87), bit for bit against
    the host codec's conversion of the oracle's normalised waveform -- odd lengths and pitches, a silent clip (peak < tiny: left
    alone), samples at +-peak (+32768 clips to 32767, -32768 stays), un-normalised input beyond +-1 (clipped)."""
    from ml_audio_inpainting_b200 import audio_io
    rng = np.random.default_rng(5)
    for B, L, pitch in ((7, 8001, 8001), (3, 4096, 4100), (5, 1, 1), (2, 159936, 159937)):
        x = (0.3 * rng.standard_normal((B, pitch))).astype(np.float32)
        x[0, :L] *= 7.0                                             # far beyond +-1 before normalisation
        if B > 1:
            x[1] = 0.0
        if B > 2 and L > 8:
            x[2, :L] = np.clip(x[2, :L], -0.5, 0.5)
            x[2, 3], x[2, 7] = 0.75, -0.75                          # the peak on both signs
        xd = torch.from_numpy(x).cuda()[:, :L]
        peaks = torch.empty(B, device="cuda")
        got = sp.wave_to_pcm16(xd, normalize=True, peaks_out=peaks).cpu().numpy()
        want = np.stack([audio_io._to_int16(up.peak_normalize(x[b, :L]), 32768.0) for b in range(B)])
        assert got.dtype == np.int16 and np.array_equal(got, want)
        assert np.array_equal(peaks.cpu().numpy(), np.abs(x[:, :L]).max(1))
        if B > 2 and L > 8:
            assert got[2, 3] == 32767 and got[2, 7] == -32768
        raw = sp.wave_to_pcm16(xd, normalize=False).cpu().numpy()
        assert np.array_equal(raw, audio_io._to_int16(x[:, :L], 32768.0))
        given = sp.wave_to_pcm16(xd, peaks=peaks).cpu().numpy()
        assert np.array_equal(given, want)
    one = sp.wave_to_pcm16(torch.from_numpy(x[0, :L]).cuda())
    assert one.shape == (L,) and np.array_equal(one.cpu().numpy(), want[0])


@pytest.mark.parametrize("n_fft,hop,win", [(512, 192, 384), (512, 128, 512), (1024, 256, 1024)])
def test_inverse_straight_to_pcm16(sp, n_fft, hop, win):
    """istft(..., normalize=True, pcm16=True) and the hand-off with pcm16: the scaling pass writes the 16-bit samples instead --
    identical to quantising the normalised float result, on the fused n_fft = 512 path and on the generic one."""
    from ml_audio_inpainting_b200 import audio_io
    B, L = 5, 24000
    x = _noise(B, L, seed=n_fft + hop)
    plan = sp.get_plan(n_fft, hop, win, "hann", True, "cuda:0")
    S = sp.stft(torch.from_numpy(x).cuda(), plan)["spec"]
    y = sp.istft(plan, spec=S, normalize=True).cpu().numpy()
    keep = torch.empty((B, plan.istft_length(S.shape[2])), device="cuda")
    peaks = torch.empty(B, device="cuda")
    q = sp.istft(plan, spec=S, normalize=True, pcm16=True, out=keep, peaks_out=peaks)
    assert q.dtype == torch.int16 and np.array_equal(q.cpu().numpy(), audio_io._to_int16(y, 32768.0))
    raw = sp.istft(plan, spec=S).cpu().numpy()
    assert np.array_equal(keep.cpu().numpy(), raw)                 # `out` keeps the un-normalised waveform
    assert np.array_equal(peaks.cpu().numpy(), np.abs(raw).max(1))
    assert np.array_equal(sp.istft(plan, spec=S, pcm16=True).cpu().numpy(), audio_io._to_int16(raw, 32768.0))
    if n_fft == 512:
        mag, ph = S.abs().log10().clamp_min(-9.0), S.angle()
        mo = torch.full_like(mag, -2.0)
        mask = torch.zeros_like(mag)
        mask[:, :, 40:50] = 1.0
        for kw in (dict(normalize=True), dict(normalize=False)):
            f = sp.istft_blend(plan, mo, mag, mask, ph, **kw).cpu().numpy()
            qq = sp.istft_blend(plan, mo, mag, mask, ph, pcm16=True, **kw).cpu().numpy()
            assert np.array_equal(qq, audio_io._to_int16(f, 32768.0)), kw


# ------------------------------------------------------------------------------------------- power-of-two n_fft other than 512
@pytest.mark.parametrize("n_fft,hop,win,L,center", [
    (2048, 512, 2048, 40000, True), (2048, 512, None, 9000, True), (1024, 256, 1024, 30001, True), (1024, 333, 800, 20000, True),
    (256, 64, 256, 12000, True), (256, 64, 200, 12000, False), (128, 32, 128, 5000, True), (64, 16, 64, 3000, True),
    (2048, 2048, 2048, 30000, True), (1024, 255, 1024, 8192, False), (512, 191, 384, 20000, True)])
def test_tiled_radix16_path_for_other_nfft(sp, n_fft, hop, win, L, center):
    """n_fft in {64 .. 2048} \\ {512 with an even hop} -- the reference's own defaults (utils.py:192-193: n_fft 2048, hop 512) -- run
    the tiled radix-16 kernels (csrc/aip_pow2.cu).  Against the oracle: complex output, every epilogue, gaps across tile borders,
    odd hops (scalar loads), no centring, frame counts that are no multiple of the tile; the inverse for complex and
    magnitude + phase input with `length`; and against the one-frame-per-CTA kernels they replace (AIP_POW2=0)."""
    B = 3
    win = win or n_fft
    x = _noise(B, L, seed=n_fft + hop)
    xd = torch.from_numpy(x).cuda()
    plan = sp.get_plan(n_fft, hop, win, "hann", center, "cuda:0")
    T = plan.num_frames(L)
    g = max(1, min(L // 4, 3 * n_fft))
    gaps = np.array([[0, g], [L // 2 - g // 2, L // 2 - g // 2 + g], [L - g, L]])
    out = sp.stft(xd, plan, gap_samples=gaps, want_spec=True, want_phase=True, mag_kind=sp.MAG_LOG10_EPS)
    with sp.experiment_env(AIP_POW2="0"):
        old = sp.stft(xd, plan, gap_samples=gaps, want_spec=True, want_phase=True, mag_kind=sp.MAG_LOG10_EPS)
    S = out["spec"].cpu().numpy()
    assert S.shape == (B, n_fft // 2 + 1, T)
    for b in range(B):
        xg = x[b].copy()
        xg[gaps[b, 0]:gaps[b, 1]] = 0
        ref = lr.stft(xg, n_fft=n_fft, hop_length=hop, win_length=win, center=center)
        assert ref.shape == S[b].shape and relerr(S[b], ref) < TOL, (b, relerr(S[b], ref))
        assert relerr(S[b], old["spec"][b].cpu().numpy()) < TOL
        big = np.abs(ref) > 1e-3 * np.abs(ref).max()
        assert np.abs(out["mag"][b].cpu().numpy() - np.log10(np.abs(ref) + 1e-9))[big].max() < 1e-3
        assert np.abs(np.exp(1j * out["phase"][b].cpu().numpy()) - np.exp(1j * np.angle(ref)))[big].max() < 2e-3
    # crop + masks + spectrum-domain gap through the general emitter
    if T > 9:
        frm = np.array([[0, 3], [T // 2, T // 2 + 4], [T - 5, T - 1]])
        o2 = sp.stft(xd, plan, mag_kind=sp.MAG_LOG1P_POW, want_spec=False, want_mask=True, mask_frames=frm,
                     mask_in_gap_is_one=True, zero_frames=frm, t_out=T - 1)
        for b in range(B):
            ref = np.abs(lr.stft(x[b], n_fft=n_fft, hop_length=hop, win_length=win, center=center))[:, :T - 1]
            ref[:, frm[b, 0]:frm[b, 1]] = 0
            m = np.zeros_like(ref)
            m[:, frm[b, 0]:frm[b, 1]] = 1
            assert np.array_equal(o2["mask"][b].cpu().numpy(), m)
            assert np.abs(o2["mag"][b].cpu().numpy() - np.log1p(ref)).max() < 1e-4 * max(1.0, np.log1p(ref).max())
    # inverse
    if center:
        Sc = sp.stft(xd, plan)["spec"]
        y = sp.istft(plan, spec=Sc).cpu().numpy()
        with sp.experiment_env(AIP_POW2_OLA_FAST="0"):      # the general gather and the power-of-two-hop one add in the same order
            assert np.array_equal(sp.istft(plan, spec=Sc).cpu().numpy(), y)
        with sp.experiment_env(AIP_POW2_SPAN="0"):          # per-warp global loads instead of the staged span: same samples
            assert torch.equal(sp.stft(xd, plan)["spec"], Sc)
        # (hop == n_fft: no overlap, the window sum-square reaches ~0 at the frame borders and the division amplifies rounding
        #  there -- compare where it is well conditioned, as test_gpu_parity does for the reference's hop-512 default)
        wss = lr.window_sumsquare("hann", T, hop_length=hop, win_length=win, n_fft=n_fft, dtype=np.float32)[n_fft // 2:]
        for b in range(B):
            ref = lr.istft(Sc[b].cpu().numpy(), hop_length=hop, win_length=win, n_fft=n_fft)
            ok = wss[:len(ref)] > 1e-2
            assert y[b].shape == ref.shape and relerr(y[b][ok], ref[ok]) < TOL, relerr(y[b][ok], ref[ok])
        length = L - 7
        y2 = sp.istft(plan, mag=Sc.abs(), phase=Sc.angle(), length=length).cpu().numpy()
        with sp.experiment_env(AIP_POW2="0"):
            y2_old = sp.istft(plan, mag=Sc.abs(), phase=Sc.angle(), length=length).cpu().numpy()
        ref = lr.istft(Sc[1].cpu().numpy(), hop_length=hop, win_length=win, n_fft=n_fft, length=length)
        ok = np.zeros(length, bool)
        ok[:min(length, len(wss))] = wss[:length] > 1e-2
        assert y2.shape == (B, length) and relerr(y2[1][ok], ref[ok]) < 5 * TOL and relerr(y2[:, ok], y2_old[:, ok]) < 5 * TOL
    else:
        # no centring: the first / last samples have a single tapering frame over them (wss -> 0) -- compare where it is conditioned
        Sc = sp.stft(xd, plan)["spec"]
        y = sp.istft(plan, spec=Sc).cpu().numpy()
        wss = lr.window_sumsquare("hann", T, hop_length=hop, win_length=win, n_fft=n_fft, dtype=np.float32)
        for b in range(B):
            ref = lr.istft(Sc[b].cpu().numpy(), hop_length=hop, win_length=win, n_fft=n_fft, center=False)
            ok = wss[:len(ref)] > 1e-2
            assert y[b].shape == ref.shape and relerr(y[b][ok], ref[ok]) < TOL, relerr(y[b][ok], ref[ok])


@pytest.mark.parametrize("n_fft,hop,L", [(2048, 512, 20000), (1024, 256, 9001), (256, 64, 3000), (64, 16, 777), (512, 191, 5000),
                                         (2048, 16, 6000)])
def test_tiled_radix16_kernels_write_inside_their_outputs(sp, n_fft, hop, L):
    """Guard regions around every output of the tiled power-of-two kernels (forward: complex / magnitude / phase / mask with a
    crop; inverse with the fused overlap-add, with `length`, and through the frame workspace for a very small hop); every
    element written, the sentinels intact, and the PCM tail on top."""
    B, G, SENT = 2, 4096, -12345.0
    x = torch.from_numpy(_noise(B, L, seed=n_fft + L)).cuda()
    plan = sp.get_plan(n_fft, hop, n_fft, "hann", True, "cuda:0")
    T, F = plan.num_frames(L), n_fft // 2 + 1

    def guarded(shape, dtype=torch.float32):
        n = int(np.prod(shape)) * (2 if dtype == torch.complex64 else 1)
        flat = torch.full((n + 2 * G,), SENT, dtype=torch.float32, device="cuda")
        body = flat[G:G + n]
        return flat, (torch.view_as_complex(body.view(*shape, 2)) if dtype == torch.complex64 else body.view(shape))

    def check(flat, view, what):
        torch.cuda.synchronize()
        ref = torch.full((G,), SENT, dtype=torch.float32, device="cuda")
        assert torch.equal(flat[:G], ref) and torch.equal(flat[-G:], ref), what
        v = torch.view_as_real(view) if view.is_complex() else view
        assert not bool(v.eq(SENT).any()), what

    for Tn in (T, T - 1):
        bufs = {k: guarded((B, F, Tn), torch.complex64 if k == "spec" else torch.float32) for k in ("spec", "mag", "phase", "mask")}
        frm = np.array([[1, 3], [Tn - 2, Tn]])
        sp.stft(x, plan, mag_kind=sp.MAG_LOG10_EPS, want_spec=True, want_phase=True, want_mask=True, mask_frames=frm,
                t_out=Tn, out={k: v[1] for k, v in bufs.items()})
        for k, (flat, view) in bufs.items():
            check(flat, view, ("fwd", k, Tn))
        S = bufs["spec"][1]
        for length in (None, plan.istft_length(Tn) - 5, plan.istft_length(Tn) + 7):
            n = plan.istft_length(Tn, length)
            flat, y = guarded((B, n))
            sp.istft(plan, spec=S, out=y, length=length)
            check(flat, y, ("inv", Tn, length))
            assert bool(torch.isfinite(y).all())
        q = sp.istft(plan, spec=S, normalize=True, pcm16=True)
        assert q.dtype == torch.int16 and int(q.to(torch.int32).abs().max()) >= 32767


# ------------------------------------------------------------------------------------------- size limits
def test_ten_minute_clip_and_seventy_thousand_clips(sp):
    """Index ranges: one 10-minute clip (9.6 M samples, 50 001 frames: row offsets beyond 2^31 bytes in the complex output) and
    70 000 short clips (more rows than a grid's y extent) through forward, inverse, normalisation and the PCM tail --
    round-trip SNR >= 100 dB on the long clip, every short clip against the oracle's first / last rows."""
    from ml_audio_inpainting_b200 import audio_io
    L = 9_600_000
    x = torch.from_numpy(_noise(1, L, seed=77)).cuda()
    for n_fft, hop, win in ((512, 192, 384), (2048, 512, 2048)):
        plan = sp.get_plan(n_fft, hop, win, "hann", True, "cuda:0")
        S = sp.stft(x, plan)["spec"]
        assert S.shape == (1, n_fft // 2 + 1, 1 + L // hop)
        y = sp.istft(plan, spec=S, length=L)
        err = (y - x).double()
        snr = 10 * torch.log10(x.double().pow(2).sum() / err.pow(2).sum()).item()
        assert snr >= 100.0, (n_fft, snr)
        # the far end of the rows against the oracle: the last 44 hops as a clip of their own (L is a multiple of both hops, so
        # its frames line up with the long clip's); all but its first n_fft / (2 hop) frames see the same samples
        ref = lr.stft(x[0, L - 44 * hop:].cpu().numpy(), n_fft=n_fft, hop_length=hop, win_length=win)
        assert ref.shape[1] == 45
        got = S[0, :, -40:].cpu().numpy()
        assert relerr(got, ref[:, -40:]) < TOL, relerr(got, ref[:, -40:])
        del S, y
    B, L2 = 70_000, 2048
    xs = torch.from_numpy(_noise(B, L2, seed=3)).cuda()
    plan = sp.get_plan(512, 192, 384, "hann", True, "cuda:0")
    out = sp.stft(xs, plan, mag_kind=sp.MAG_LOG10_EPS, want_spec=True)
    yb = sp.istft(plan, spec=out["spec"], normalize=True, length=L2)
    q = sp.wave_to_pcm16(xs, normalize=True)
    for b in (0, 65535, 65536, B - 1):
        xb = xs[b].cpu().numpy()
        ref = lr.stft(xb, n_fft=512, hop_length=192, win_length=384)
        assert relerr(out["spec"][b].cpu().numpy(), ref) < TOL
        assert np.abs(out["mag"][b].cpu().numpy() - np.log10(np.abs(ref) + 1e-9)).max() < 1e-3
        ry = up.peak_normalize(lr.istft(ref, hop_length=192, win_length=384, n_fft=512, length=L2))
        assert relerr(yb[b].cpu().numpy(), ry) < TOL
        assert np.array_equal(q[b].cpu().numpy(), audio_io._to_int16(up.peak_normalize(xb), 32768.0))


@pytest.mark.parametrize("n_fft,hop", [(2048, 512), (1024, 256), (256, 64)])
def test_griffinlim_on_the_tiled_kernels(sp, n_fft, hop):
    """Griffin-Lim at the reference's default transform size (utils.mel_spectrogram_to_audio: n_fft 2048 / hop 512, utils.py:335-341)
    and the other tiled sizes: injected initial phasors, value parity with the oracle, and the phase update fused into the tiled
    inverse's load against the separate update kernel (AIP_GL_UNFUSED=1)."""
    B, L = 2, 24000
    x = _noise(B, L, seed=n_fft)
    m = np.stack([np.abs(lr.stft(x[b], n_fft=n_fft, hop_length=hop)) for b in range(B)]).astype(np.float32)
    rng = np.random.default_rng(n_fft)
    ang = np.exp(2j * np.pi * rng.random(m.shape)).astype(np.complex64)
    plan = sp.get_plan(n_fft, hop, n_fft, "hann", True, "cuda:0")
    mag = torch.from_numpy(m).cuda()
    for n_iter, bound in ((0, 1e-4), (1, 1e-4), (2, 1e-4), (8, 1e-4), (32, 3e-3)):
        y = sp.griffinlim(plan, mag, n_iter=n_iter, init_angles=torch.from_numpy(ang).cuda()).cpu().numpy()
        with sp.experiment_env(AIP_GL_UNFUSED="1"):
            y_unf = sp.griffinlim(plan, mag, n_iter=n_iter, init_angles=torch.from_numpy(ang).cuda()).cpu().numpy()
        assert relerr(y, y_unf) < bound, (n_iter, relerr(y, y_unf))
        if n_iter <= 8:
            for b in range(B):
                ref = lr.griffinlim(m[b], n_iter=n_iter, hop_length=hop, win_length=n_fft, n_fft=n_fft, init_angles=ang[b])
                assert y[b].shape == ref.shape and relerr(y[b], ref) < bound, (n_iter, b, relerr(y[b], ref))


# ------------------------------------------------------------------------------------------- output combinations
def test_combined_outputs_equal_separate_launches(sp):
    """Every combination of outputs of aip_stft_fwd_f32 against the same outputs from single-purpose launches: the combinations
    with a straight-line variant of their own (spectrogram + |S| / log10 / log1p, log10 + phase, |S|**2) bit for bit, the rest
    (general emitter: run-time switches, |S|**p) to rounding; and the general emitter must not be the 10 - 20 x cliff it was."""
    B, L = 6, 20000
    x = torch.from_numpy(_noise(B, L, seed=9)).cuda()
    plan = sp.get_plan(512, 192, 384, "hann", True, "cuda:0")
    T = plan.num_frames(L)
    gaps = np.stack([np.arange(B) * 1000 + 500, np.arange(B) * 1000 + 3700], 1)
    spec = sp.stft(x, plan, gap_samples=gaps)["spec"]
    single = {k: sp.stft(x, plan, gap_samples=gaps, mag_kind=k, want_spec=False)["mag"]
              for k in (sp.MAG_ABS, sp.MAG_LOG10_EPS, sp.MAG_LOG1P_POW)}
    phase = sp.stft(x, plan, gap_samples=gaps, mag_kind=sp.MAG_ABS, want_spec=False, want_phase=True)["phase"]
    for k in single:                                            # own variants
        o = sp.stft(x, plan, gap_samples=gaps, mag_kind=k, want_spec=True)
        assert torch.equal(o["spec"], spec) and torch.equal(o["mag"], single[k]), k
    o = sp.stft(x, plan, gap_samples=gaps, mag_kind=sp.MAG_LOG10_EPS, want_spec=False, want_phase=True)
    assert torch.equal(o["mag"], single[sp.MAG_LOG10_EPS]) and torch.equal(o["phase"], phase)
    p2 = sp.stft(x, plan, gap_samples=gaps, mag_kind=sp.MAG_POW, power=2.0, want_spec=False)["mag"]
    assert torch.allclose(p2, single[sp.MAG_ABS] ** 2, rtol=2e-6, atol=0)
    # general emitter: everything at once, and a power the straight-line code does not know
    frm = np.stack([np.full(B, 10), np.full(B, 20)], 1)
    o = sp.stft(x, plan, gap_samples=gaps, mag_kind=sp.MAG_LOG10_EPS, want_spec=True, want_phase=True, want_mask=True,
                mask_frames=frm, t_out=T - 3)
    assert torch.equal(o["spec"], spec[:, :, :T - 3]) and torch.equal(o["phase"], phase[:, :, :T - 3])
    assert torch.allclose(o["mag"], single[sp.MAG_LOG10_EPS][:, :, :T - 3], rtol=0, atol=2e-6)
    assert float(o["mask"][:, :, 10:20].min()) == 1.0 and float(o["mask"].sum()) == B * 257 * 10
    p3 = sp.stft(x, plan, gap_samples=gaps, mag_kind=sp.MAG_POW, power=3.0, want_spec=True)
    assert torch.equal(p3["spec"], spec) and torch.allclose(p3["mag"], single[sp.MAG_ABS] ** 3, rtol=2e-5, atol=0)
    # timing guard (generous): the general emitter within 4 x of the two specialised launches it stands for
    xb = torch.from_numpy(_noise(512, 80000, seed=1)).cuda()

    def ms(fn):
        for _ in range(2):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / 5
    t_sep = ms(lambda: sp.stft(xb, plan)) + ms(lambda: sp.stft(xb, plan, mag_kind=sp.MAG_ABS, want_spec=False, want_phase=True))
    t_all = ms(lambda: sp.stft(xb, plan, mag_kind=sp.MAG_ABS, want_spec=True, want_phase=True, want_mask=True))
    assert t_all < 4 * t_sep, (t_all, t_sep)


# ------------------------------------------------------------------------------------------- the functions' own defaults on real speech
def test_default_parameters_on_the_reference_clips(utils, golden_clips):
    """utils.extract_spectrogram / spectrogram_to_audio / extract_mel_spectrogram called the way their signatures default
    (n_fft 2048, hop 512: the tiled radix-16 kernels) on the reference's nine test clips: value parity with the oracle at the
    1e-4 bound of SURVEY 8c, the reference suite's float64 round trip (tests/utils_test.py:780-809) as an SNR floor of
    100 dB, and the mel front-end against librosa's melspectrogram."""
    for name in sorted(golden_clips):
        x = golden_clips[name]
        S = utils.extract_spectrogram(x)
        ref = up.extract_spectrogram(x)
        assert S.shape == ref.shape == (1025, 157) and S.dtype == np.complex64
        assert relerr(S, ref) < TOL, (name, relerr(S, ref))
        y = utils.spectrogram_to_audio(S, phase_info=True, n_fft=2048, hop_length=512)
        ry = up.spectrogram_to_audio(ref, phase_info=True, n_fft=2048, hop_length=512)
        assert y.shape == ry.shape == (512 * 156,) and relerr(y, ry) < TOL, (name, relerr(y, ry))
        n = len(y)
        snr = 10 * np.log10(np.sum(x[:n].astype(np.float64) ** 2) / np.sum((y - x[:n]).astype(np.float64) ** 2))
        assert snr >= 100.0, (name, snr)
        m = utils.extract_mel_spectrogram(x)
        rm = lr.melspectrogram(y=x, sr=SR)
        assert m.shape == rm.shape == (128, 157) and relerr(m, rm) < TOL, (name, relerr(m, rm))
        # float64 in -> the reference's dtypes out (complex128 / float64), values from the same fp32 kernels
        S64 = utils.extract_spectrogram(x.astype(np.float64))
        assert S64.dtype == np.complex128 and relerr(S64, ref) < TOL


# ------------------------------------------------------------------------------------------- the C ABI's error behaviour
def test_c_abi_rejects_bad_arguments_without_touching_memory(sp):
    """include/aip_b200.h: 0 on success, negative for argument errors (ARG -1, UNSUPPORTED -2, WORKSPACE -4), nothing launched,
    nothing written -- called raw through ctypes the way a foreign binding would."""
    import ctypes as C
    from ml_audio_inpainting_b200 import _cabi
    lib = _cabi.load()
    B, L = 2, 4000
    x = torch.from_numpy(_noise(B, L, seed=1)).cuda()
    plan = sp.get_plan(512, 192, 384, "hann", True, "cuda:0")
    T = plan.num_frames(L)
    SENT = -777.0
    mag = torch.full((B, 257, T), SENT, device="cuda")
    spec = torch.full((B, 257, T, 2), SENT, device="cuda")
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    d = C.byref(plan.desc)
    win = plan.desc.window

    def fwd(desc=d, wave=x.data_ptr(), b=B, l=L, pitch=L, mk=_cabi.MAG_LOG10_EPS, t_out=T, spec_p=None, mag_p=mag.data_ptr()):
        return lib.aip_stft_fwd_f32(desc, wave, b, l, pitch, None, None, None, 1, mk, 1e-9, 1.0, t_out, spec_p, mag_p, None, None, st)

    assert fwd() == 0 and not bool(mag.eq(SENT).any())
    mag.fill_(SENT)
    assert fwd(wave=None) == -1 and fwd(desc=None) == -1
    assert fwd(mag_p=None) == -1                                   # a magnitude kind without its output
    assert fwd(mk=_cabi.MAG_NONE) == -1                            # ... and an output without a kind
    assert fwd(mk=9) == -1 and fwd(t_out=T + 1) == -1 and fwd(pitch=L - 1) == -1 and fwd(b=-1) == -1
    assert fwd(b=0) == 0 and fwd(t_out=0) == 0
    assert fwd(mk=_cabi.MAG_NONE, mag_p=None) == 0                 # no output at all: accepted, nothing launched
    big_desc = sp.get_plan(2048, 512, 2048, "hann", True, "cuda:0").desc
    assert lib.aip_stft_fwd_f32(C.byref(big_desc), x.data_ptr(), B, L, L, None, None, None, 1, 0, 0.0, 1.0, 1, None, None, None, None, st) == 0
    assert fwd(l=100) == -1                                        # 100 samples hold one frame, T_out asks for more
    bad = _cabi.StftDesc(500, 192, 1, 0, win)                     # not a power of two
    assert fwd(desc=C.byref(bad)) == -2
    assert fwd(desc=C.byref(_cabi.StftDesc(512, 0, 1, 0, win))) == -2
    assert fwd(desc=C.byref(_cabi.StftDesc(512, 192, 1, 0, None))) == -1
    torch.cuda.synchronize()
    assert bool(mag.eq(SENT).all())                                # none of the rejected calls wrote anything
    # inverse
    S = sp.stft(x, plan)["spec"]
    n = plan.istft_length(T)
    inv = plan.inv_wss(T)
    y = torch.full((B, n), SENT, device="cuda")

    def inverse(desc=d, spec_p=S.data_ptr(), mag_p=None, b=B, t=T, length=0, wss=inv.data_ptr(), out=y.data_ptr(), pitch=n, ws=None, wsb=0,
                dom=0):
        return lib.aip_istft_f32(desc, spec_p, mag_p, None, dom, None, b, t, length, wss, out, pitch, ws, wsb, st)

    assert inverse() == 0 and not bool(y.eq(SENT).any())
    y.fill_(SENT)
    assert inverse(spec_p=None) == -1 and inverse(out=None) == -1 and inverse(wss=None) == -1
    assert inverse(pitch=n - 1) == -1 and inverse(t=0) == -1 and inverse(length=-5) == -1 and inverse(dom=7) == -1
    assert inverse(desc=C.byref(bad)) == -2
    big = sp.get_plan(4096, 1024, 4096, "hann", True, "cuda:0")    # off every tiled kernel: needs the frame workspace
    assert lib.aip_istft_workspace_bytes(C.byref(big.desc), B, 8) == B * 8 * 4096 * 4
    S4 = torch.zeros((B, 2049, 8, 2), device="cuda")
    y4 = torch.empty((B, big.istft_length(8)), device="cuda")
    assert lib.aip_istft_f32(C.byref(big.desc), S4.data_ptr(), None, None, 0, None, B, 8, 0, big.inv_wss(8).data_ptr(), y4.data_ptr(),
                             y4.shape[1], None, 0, st) == -4
    q = torch.empty((B, L), dtype=torch.int16, device="cuda")
    assert lib.aip_wave_to_pcm16_f32(x.data_ptr(), L, q.data_ptr(), L, B, L, 5, None, st) == -1
    assert lib.aip_wave_to_pcm16_f32(x.data_ptr(), L, q.data_ptr(), L, B, L, 1, None, st) == -1      # normalise without a peaks array
    assert lib.aip_wave_to_pcm16_f32(x.data_ptr(), L, q.data_ptr(), L - 1, B, L, 0, None, st) == -1
    torch.cuda.synchronize()
    assert bool(y.eq(SENT).all())
    for code, word in ((0, b"ok"), (-1, b"argument"), (-2, b"unsupported"), (-3, b"no fallback"), (-4, b"workspace")):
        assert word in lib.aip_status_string(code)


def test_randomised_geometry_on_the_tiled_kernels(sp):
    """Seeded random geometries through the tiled power-of-two kernels (and n_fft 512 off its fused path): n_fft, hop (odd and even,
    tiny to larger than the window), win_length, clip length, centring, batch size, gaps -- forward against the oracle, inverse
    against the oracle where the overlap-add is well conditioned."""
    rng = np.random.default_rng(2026)
    for trial in range(24):
        n_fft = int(rng.choice([64, 128, 256, 512, 1024, 2048]))
        hop = int(rng.choice([max(1, n_fft // 16), n_fft // 8, n_fft // 4, n_fft // 4 + 1, n_fft // 2, n_fft // 3, n_fft, n_fft // 4 - 1]))
        if n_fft == 512 and hop % 2 == 0:
            hop += 1                                           # keep it off the fused n_fft = 512 kernels
        win = int(rng.choice([n_fft, n_fft, 3 * n_fft // 4, n_fft // 2 + 3]))
        center = bool(rng.integers(0, 2))
        B = int(rng.integers(1, 4))
        L = int(rng.integers(n_fft + 1, 12 * n_fft + 977))
        x = _noise(B, L, seed=trial)
        xd = torch.from_numpy(x).cuda()
        plan = sp.get_plan(n_fft, hop, win, "hann", center, "cuda:0")
        T = plan.num_frames(L)
        g0 = rng.integers(0, L - 1, size=B)
        gaps = np.stack([g0, np.minimum(L, g0 + rng.integers(1, L // 2 + 2, size=B))], 1)
        S = sp.stft(xd, plan, gap_samples=gaps)["spec"]
        what = dict(trial=trial, n_fft=n_fft, hop=hop, win=win, center=center, B=B, L=L)
        assert S.shape == (B, n_fft // 2 + 1, T), what
        for b in range(B):
            xg = x[b].copy()
            xg[gaps[b, 0]:gaps[b, 1]] = 0
            ref = lr.stft(xg, n_fft=n_fft, hop_length=hop, win_length=win, center=center)
            assert relerr(S[b].cpu().numpy(), ref) < TOL, (what, b, relerr(S[b].cpu().numpy(), ref))
        if hop <= win // 2:                                    # enough overlap for a conditioned inverse
            Sc = sp.stft(xd, plan)["spec"]
            y = sp.istft(plan, spec=Sc).cpu().numpy()
            wss = lr.window_sumsquare("hann", T, hop_length=hop, win_length=win, n_fft=n_fft, dtype=np.float32)
            wss = wss[n_fft // 2:] if center else wss
            for b in range(B):
                ref = lr.istft(Sc[b].cpu().numpy(), hop_length=hop, win_length=win, n_fft=n_fft, center=center)
                ok = wss[:len(ref)] > 1e-2 * wss.max()
                assert y[b].shape == ref.shape and relerr(y[b][ok], ref[ok]) < TOL, (what, b, relerr(y[b][ok], ref[ok]))


# ------------------------------------------------------------------------------------------- a compiled host on the C ABI
def test_cpp_host_drives_the_c_abi_without_python(tmp_path):
    """examples/abi_host_demo.cu: cudaMalloc'ed buffers, one stream, forward (log-magnitude of the gapped clips + complex
    spectrogram) and inverse through include/aip_b200.h alone; it checks the round-trip SNR (>= 100 dB) and the -9 floor of the
    frames inside the gap itself and returns non-zero otherwise."""
    import shutil
    import subprocess
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not Path(nvcc).exists():
        pytest.skip("no nvcc on this box")
    root = Path(__file__).resolve().parents[1]
    libdir = root / "ml_audio_inpainting_b200" / "lib"
    exe = tmp_path / "abi_host_demo"
    res = subprocess.run([nvcc, "-std=c++17", "-I", str(root / "include"), "-gencode", "arch=compute_100a,code=sm_100a", "-o", str(exe),
                          str(root / "examples" / "abi_host_demo.cu"), f"-L{libdir}", "-laip_b200", "-Xlinker", "-rpath", "-Xlinker",
                          str(libdir)], capture_output=True, text=True)
    assert res.returncode == 0, res.stderr
    run = subprocess.run([str(exe)], capture_output=True, text=True, timeout=120)
    assert run.returncode == 0, (run.returncode, run.stdout, run.stderr)
    assert "round-trip SNR" in run.stdout
    print(run.stdout.strip())


# ------------------------------------------------------------------------------------------- Griffin-Lim's random initial phases
def test_random_phasors_and_seeded_griffinlim(sp):
    """aip_random_phasors_f32: unit phasors with uniform, uncorrelated phases (Philox4x32-10 keyed by the seed); a seed fixes
    the stream, torch.manual_seed fixes griffinlim(init='random'), two calls in a row differ."""
    import ctypes as C
    from ml_audio_inpainting_b200 import _cabi
    lib = _cabi.load()
    n = 1_000_003                                               # not a multiple of four: the last quad is partial
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    a = torch.full((n + 8, 2), -5.0, device="cuda")
    b = torch.empty((n, 2), device="cuda")
    assert lib.aip_random_phasors_f32(a.data_ptr(), n, 1234, st) == 0
    assert lib.aip_random_phasors_f32(b.data_ptr(), n, 1234, st) == 0
    assert torch.equal(a[:n], b) and bool((a[n:] == -5.0).all())                  # same key, same stream; nothing past n
    assert lib.aip_random_phasors_f32(b.data_ptr(), n, 1235, st) == 0
    z, z2 = torch.view_as_complex(a[:n].contiguous()), torch.view_as_complex(b)
    assert float((z.abs() - 1).abs().max()) < 1e-6
    phi = torch.angle(z).double()
    hist = torch.histc(phi.float(), bins=64, min=-np.pi, max=np.pi).double()
    chi2 = float(((hist - n / 64) ** 2 / (n / 64)).sum())
    assert chi2 < 130, chi2                                     # 63 degrees of freedom: P(chi2 > 130) ~ 1e-6
    assert float(z.mean().abs()) < 5e-3 and float((z * z2.conj()).mean().abs()) < 5e-3     # no bias, keys independent
    assert float((z[1:] * z[:-1].conj()).mean().abs()) < 5e-3 and float((z[4:] * z[:-4].conj()).mean().abs()) < 5e-3
    plan = sp.get_plan(512, 192, 384, "hann", True, "cuda:0")
    mag = sp.stft(torch.from_numpy(_noise(2, 8000, seed=4)).cuda(), plan, mag_kind=sp.MAG_ABS, want_spec=False)["mag"]
    torch.manual_seed(7)
    y1 = sp.griffinlim(plan, mag, n_iter=4)
    y2 = sp.griffinlim(plan, mag, n_iter=4)
    torch.manual_seed(7)
    y3 = sp.griffinlim(plan, mag, n_iter=4)
    assert torch.equal(y1, y3) and not torch.equal(y1, y2) and bool(torch.isfinite(y1).all())
    g = torch.Generator(device="cuda").manual_seed(3)
    y4 = sp.griffinlim(plan, mag, n_iter=4, generator=g)
    y5 = sp.griffinlim(plan, mag, n_iter=4, generator=torch.Generator(device="cuda").manual_seed(3))
    assert torch.equal(y4, y5)


def test_expm1_back_end_on_the_device(sp):
    """AIP_DOM_EXPM1 (undoing the GAN front-end's log1p) with the device's ex2.approx-based expm1 against numpy.expm1 in float64:
    magnitudes spanning tiny to large, through both the fused n_fft = 512 inverse and the tiled one."""
    rng = np.random.default_rng(12)
    for n_fft, hop in ((512, 128), (1024, 256)):
        plan = sp.get_plan(n_fft, hop, n_fft, "hann", True, "cuda:0")
        F, T, B = n_fft // 2 + 1, 40, 2
        logm = np.concatenate([rng.uniform(0, 6, (B, F, T // 2)), 10.0 ** rng.uniform(-6, -1, (B, F, T - T // 2))], 2).astype(np.float32)
        ph = rng.uniform(-np.pi, np.pi, (B, F, T)).astype(np.float32)
        y = sp.istft(plan, mag=torch.from_numpy(logm).cuda(), phase=torch.from_numpy(ph).cuda(), mag_domain=sp.DOM_EXPM1).cpu().numpy()
        for b in range(B):
            S = (np.expm1(logm[b].astype(np.float64)) * np.exp(1j * ph[b].astype(np.float64))).astype(np.complex64)
            ref = lr.istft(S, hop_length=hop, win_length=n_fft, n_fft=n_fft)
            assert relerr(y[b], ref) < TOL, (n_fft, relerr(y[b], ref))


# ------------------------------------------------------------------------------------------- librosa < 0.10's centre padding
@pytest.mark.parametrize("n_fft,hop,win", [(512, 192, 384), (512, 128, 512), (512, 191, 384), (2048, 512, 2048), (1024, 255, 1024),
                                           (256, 64, 256), (4096, 1024, 4096)])
def test_reflect_padding_mode(sp, n_fft, hop, win):
    """get_plan(pad_mode="reflect") = librosa.stft's default before 0.10 (the reference pins librosa>=0.8.1 only): the fused
    n_fft = 512 kernels (bulk copy + fix-up), the tiled kernels (staged span and per-warp checked loads) and the one-frame-per-CTA
    kernel, gaps touching both clip ends, against the oracle's np.pad(mode="reflect") path; the inverse is unaffected and a
    clip no longer than the padding is refused like np.pad refuses it."""
    B, L = 3, 5 * n_fft + 333
    x = _noise(B, L, seed=n_fft + hop)
    xd = torch.from_numpy(x).cuda()
    plan = sp.get_plan(n_fft, hop, win, "hann", True, "cuda:0", pad_mode="reflect")
    gaps = np.array([[0, n_fft // 3], [L - n_fft // 4, L], [L // 2, L // 2 + 100]])
    out = sp.stft(xd, plan, gap_samples=gaps, mag_kind=sp.MAG_LOG10_EPS, want_spec=True)
    zero_plan = sp.get_plan(n_fft, hop, win, "hann", True, "cuda:0")
    assert zero_plan is not plan and plan.pad_mode == "reflect"
    for b in range(B):
        xg = x[b].copy()
        xg[gaps[b, 0]:gaps[b, 1]] = 0
        ref = lr.stft(xg, n_fft=n_fft, hop_length=hop, win_length=win, pad_mode="reflect")
        got = out["spec"][b].cpu().numpy()
        assert got.shape == ref.shape and relerr(got, ref) < TOL, (b, relerr(got, ref))
    S = sp.stft(xd, plan)["spec"]
    assert relerr(sp.stft(xd, zero_plan)["spec"].cpu().numpy(), S.cpu().numpy()) > 1e-2      # the modes do differ
    y = sp.istft(plan, spec=S, length=L).cpu().numpy()
    for b in range(B):
        ref = lr.istft(S[b].cpu().numpy(), hop_length=hop, win_length=win, n_fft=n_fft, length=L)
        assert relerr(y[b], ref) < TOL
    with pytest.raises(Exception):
        sp.stft(xd[:, :n_fft // 2], plan)
    with pytest.raises(NotImplementedError):
        sp.get_plan(n_fft, hop, win, "hann", True, "cuda:0", pad_mode="edge")


def test_dropin_follows_the_default_pad_mode(sp, utils, monkeypatch):
    """spectral.DEFAULT_PAD_MODE = "reflect" makes every caller that does not choose (the drop-in utils.py, the front-ends) behave
    like a librosa < 0.10 installation."""
    x = _noise(1, 9000, seed=5)[0]
    monkeypatch.setattr(sp, "DEFAULT_PAD_MODE", "reflect")
    S = utils.extract_spectrogram(x, n_fft=512, hop_length=192, win_length=384)
    assert relerr(S, lr.stft(x, n_fft=512, hop_length=192, win_length=384, pad_mode="reflect")) < TOL
    S2 = utils.extract_spectrogram(x)
    assert relerr(S2, lr.stft(x, n_fft=2048, hop_length=512, pad_mode="reflect")) < TOL
    monkeypatch.setattr(sp, "DEFAULT_PAD_MODE", "constant")
    S = utils.extract_spectrogram(x, n_fft=512, hop_length=192, win_length=384)
    assert relerr(S, lr.stft(x, n_fft=512, hop_length=192, win_length=384)) < TOL
