"""CPU tier: the kernels' own per-thread code (csrc/aip_tiles.cuh compiled by g++, replayed thread by thread)
against the oracle.  This checks the algorithm, the lane/warp maps, the exchange-buffer layout, the tile edge
handling and the overlap-add WITHOUT a GPU; the GPU tier (test_gpu_parity.py) checks the real launches."""
import numpy as np
import pytest

from oracle import callers_port as cp
from oracle import librosa_port as lr
from tests import emul

TOL = 1e-4


def relerr(a, b):
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))


def noise(B, L, seed):
    rng = np.random.default_rng(seed)
    return np.clip(0.1 * rng.standard_normal((B, L)), -1, 1).astype(np.float32)


def win(win_length, name="hann"):
    return lr.fft_window(name, win_length, 512).astype(np.float32)


def inv_wss(window_name, win_length, hop, T, out_len, center=True):
    wss = lr.window_sumsquare(window_name, T, hop_length=hop, win_length=win_length, n_fft=512, dtype=np.float32)
    start = 256 if center else 0
    wss = wss[start:start + out_len]
    if len(wss) < out_len:
        wss = np.pad(wss, (0, out_len - len(wss)))
    with np.errstate(divide="ignore"):
        return np.where(wss > np.finfo(np.float32).tiny, 1.0 / wss, 1.0).astype(np.float32)


@pytest.mark.parametrize("hop,wl", [(192, 384), (128, 512), (512, 512), (64, 256), (250, 500)])
@pytest.mark.parametrize("L", [512, 777, 6001, 16000])
def test_forward_complex(hop, wl, L):
    x = noise(2, L, seed=hop + L)
    S = emul.stft(x, hop, win(wl), win_length=wl)["spec"]
    for b in range(2):
        ref = lr.stft(x[b], n_fft=512, hop_length=hop, win_length=wl)
        assert S[b].shape == ref.shape
        assert relerr(S[b], ref) < TOL


@pytest.mark.parametrize("wl", [384, 256, 300])
def test_forward_zero_tap_pruning_matches_full_transform(wl):
    """win_length <= 384 lets stage 1 skip the 2 x 64 zero taps of the centre-padded window: same spectrum."""
    x = noise(2, 9000, seed=wl)
    full = emul.stft(x, 192, win(wl))["spec"]
    pruned = emul.stft(x, 192, win(wl), win_length=wl)["spec"]
    assert relerr(pruned, full) < 2e-6
    assert relerr(pruned[0], lr.stft(x[0], n_fft=512, hop_length=192, win_length=wl)) < TOL


@pytest.mark.parametrize("center", [True, False])
@pytest.mark.parametrize("vec_ok", [True, False])
def test_forward_center_and_scalar_path(center, vec_ok):
    x = noise(1, 5003, seed=3)
    S = emul.stft(x, 192, win(384), center=center, vec_ok=vec_ok, win_length=384)["spec"]
    ref = lr.stft(x[0], n_fft=512, hop_length=192, win_length=384, center=center)
    assert S[0].shape == ref.shape and relerr(S[0], ref) < TOL


@pytest.mark.parametrize("window", ["hann", "hamming", "blackman"])
def test_forward_windows(window):
    x = noise(1, 8000, seed=9)
    S = emul.stft(x, 128, win(400, window))["spec"]
    ref = lr.stft(x[0], n_fft=512, hop_length=128, win_length=400, window=window)
    assert relerr(S[0], ref) < TOL


def test_forward_gap_logmag_mask_crop():
    L, B = 16000, 4
    x = noise(B, L, seed=1)
    gaps = np.array([[0, 1600], [7000, 10200], [L - 1600, L], [5000, 5000]])
    frames = np.array([[0, 9], [36, 54], [75, 84], [10, 10]])
    out = emul.stft(x, 192, win(384), win_length=384, gap_samples=gaps, mask_frames=frames, mag_kind=2, want_spec=False,
                    want_mask=True, t_out=80)
    for b in range(B):
        xg = x[b].copy()
        xg[gaps[b, 0]:gaps[b, 1]] = 0
        ref = np.abs(lr.stft(xg, n_fft=512, hop_length=192, win_length=384))[:, :80]
        assert out["mag"][b].shape == (257, 80)
        assert relerr(10.0 ** out["mag"][b].astype(np.float64), ref + 1e-9) < TOL
        m = np.zeros((257, 80), np.float32)
        m[:, frames[b, 0]:frames[b, 1]] = 1
        assert np.array_equal(out["mask"][b], m)
    # frames that see only zeros are exactly log10(1e-9)
    assert np.all(out["mag"][1][:, 40:50] == np.float32(-9.0))


def test_forward_gan_epilogue_and_spec_gap():
    L = 16000
    x = noise(1, L, seed=5)
    ref = cp.gan_item(x[0], gap_start_s=0.4)
    s0, s1 = ref["gap_samples"]
    f0, f1 = ref["gap_frames"]
    o = emul.stft(x, 128, win(512), mag_kind=3, want_spec=False, want_phase=True, want_mask=True,
                  mask_frames=np.array([[f0, f1]]), mask_in_gap_is_one=False)
    assert np.array_equal(o["mask"][0], ref["mask"])
    assert relerr(np.expm1(o["mag"][0].astype(np.float64)), np.expm1(ref["original_magnitude"].astype(np.float64))) < TOL
    w = np.expm1(ref["original_magnitude"].astype(np.float64))
    assert float((np.abs(np.exp(1j * o["phase"][0]) - np.exp(1j * ref["original_phase"])) * w / w.max()).max()) < TOL
    imp = emul.stft(x, 128, win(512), gap_samples=np.array([[s0, s1]]), mag_kind=3, want_spec=False)
    assert relerr(np.expm1(imp["mag"][0].astype(np.float64)), np.expm1(ref["impaired_magnitude"].astype(np.float64))) < TOL
    # model_eval.py:154 spectrum-domain gap
    ev = cp.eval_frontend_cnnlstm(x[0], t0=0.3, t1=0.38)
    z = emul.stft(x, 192, win(384), win_length=384, zero_frames=np.array([ev["gap_frames"]]), mag_kind=2, want_spec=False)
    assert relerr(10.0 ** z["mag"][0].astype(np.float64), 10.0 ** ev["log_impaired_magnitude"].astype(np.float64)) < TOL
    assert np.all(z["mag"][0][:, ev["gap_frames"][0]:ev["gap_frames"][1]] == np.float32(-9.0))


@pytest.mark.parametrize("hop,wl", [(192, 384), (128, 512), (512, 512), (64, 256), (250, 500)])
@pytest.mark.parametrize("L", [2000, 16000])
def test_inverse_complex(hop, wl, L):
    x = noise(2, L, seed=7 * hop + L)
    S = np.stack([lr.stft(x[b], n_fft=512, hop_length=hop, win_length=wl) for b in range(2)])
    T = S.shape[2]
    ref = np.stack([lr.istft(S[b], hop_length=hop, win_length=wl, n_fft=512) for b in range(2)])
    y = emul.istft(hop, win(wl), inv_wss("hann", wl, hop, T, ref.shape[1]), spec=S, win_length=wl)
    good = np.ones(ref.shape[1], bool)
    if hop == wl:   # window-sum-square has (near-)zeros: compare where it is well conditioned
        good = inv_wss("hann", wl, hop, T, ref.shape[1]) < 100.0
    for b in range(2):
        assert relerr(y[b][good], ref[b][good]) < TOL


@pytest.mark.parametrize("hop,wl", [(192, 384), (192, 256), (128, 512), (128, 384)])
def test_inverse_specialised_overlap_add_matches_generic(hop, wl):
    """Interior tiles of hop 192 / 128 run the compile-time overlap-add (and, for win <= 384 at hop 192, the
    output-pruned stage B); win_length = 0 keeps the generic phases: same waveform, many tiles, ragged tail."""
    x = noise(2, 48000 + 77, seed=hop + wl)
    S = np.stack([lr.stft(x[b], n_fft=512, hop_length=hop, win_length=wl) for b in range(2)])
    T = S.shape[2]
    iw = inv_wss("hann", wl, hop, T, hop * (T - 1))
    fast = emul.istft(hop, win(wl), iw, spec=S, win_length=wl)
    generic = emul.istft(hop, win(wl), iw, spec=S, win_length=0) if hop == 192 else None
    ref = np.stack([lr.istft(S[b], hop_length=hop, win_length=wl, n_fft=512) for b in range(2)])
    assert relerr(fast, ref) < TOL
    if generic is not None:
        assert relerr(fast, generic) < 2e-6


@pytest.mark.parametrize("length", [3000, 15872, 16000, 20000])
def test_inverse_length_argument(length):
    x = noise(1, 16000, seed=2)
    S = lr.stft(x[0], n_fft=512, hop_length=192, win_length=384)
    ref = lr.istft(S, hop_length=192, win_length=384, n_fft=512, length=length)
    T = S.shape[1]
    n_frames = min(T, int(np.ceil((length + 512) / 192)))
    iw = inv_wss("hann", 384, 192, n_frames, length)
    y = emul.istft(192, win(384), iw, spec=S[None], length=length, win_length=384)
    assert y.shape[1] == len(ref) == length
    good = iw < 100.0        # past the last frame's centre the window sum decays to ~0: ill-conditioned in the reference too
    assert good.sum() >= min(length, 15936) - 8 and relerr(y[0][good], ref[good]) < TOL
    assert np.all(y[0][16384:] == 0) and np.all(ref[16384:] == 0)


def test_inverse_uncentered():
    x = noise(1, 9000, seed=4)
    S = lr.stft(x[0], n_fft=512, hop_length=128, win_length=512, center=False)
    ref = lr.istft(S, hop_length=128, win_length=512, n_fft=512, center=False)
    y = emul.istft(128, win(512), inv_wss("hann", 512, 128, S.shape[1], len(ref), center=False), spec=S[None], center=False)
    inner = slice(512, len(ref) - 512)          # the un-centred ends divide by tiny window sums
    assert relerr(y[0][inner], ref[inner]) < TOL


def test_inverse_mag_phase_and_domains():
    x = noise(1, 8000, seed=6)
    S = lr.stft(x[0], n_fft=512, hop_length=192, win_length=384)
    T = S.shape[1]
    ref = lr.istft((np.abs(S) * np.exp(1j * np.angle(S))).astype(np.complex64), hop_length=192, win_length=384, n_fft=512)
    iw = inv_wss("hann", 384, 192, T, len(ref))
    mag, ph = np.abs(S)[None], np.angle(S)[None]
    assert relerr(emul.istft(192, win(384), iw, mag=mag, phase=ph, win_length=384)[0], ref) < TOL
    assert relerr(emul.istft(192, win(384), iw, mag=np.log10(mag + 1e-12), phase=ph, mag_domain=1)[0], ref) < 2 * TOL
    db = 20 * np.log10(mag / mag.max() * 0.5 + 1e-12)
    ref_db = lr.istft((lr.db_to_amplitude(db[0]) * np.exp(1j * ph[0])).astype(np.complex64), hop_length=192, win_length=384, n_fft=512)
    assert relerr(emul.istft(192, win(384), iw, mag=db, phase=ph, mag_domain=2)[0], ref_db) < 2 * TOL
    assert relerr(emul.istft(192, win(384), iw, mag=db, phase=ph, db_flags=np.array([1]))[0], ref_db) < 2 * TOL
    assert relerr(emul.istft(192, win(384), iw, mag=np.log1p(mag), phase=ph, mag_domain=3)[0], ref) < 2 * TOL
    # imag(DC) / imag(Nyquist) are ignored, as scipy.fft.irfft does
    S2 = S.copy()
    S2[0] += 3j
    S2[-1] -= 2j
    assert relerr(emul.istft(192, win(384), iw, spec=S2[None])[0], lr.istft(S, hop_length=192, win_length=384, n_fft=512)) < TOL


def test_inverse_blend_prologue():
    """model.py:108 blend + 10** + phase reuse fused into the inverse prologue."""
    x = noise(1, 8000, seed=12)
    S = lr.stft(x[0], n_fft=512, hop_length=192, win_length=384)
    T = S.shape[1]
    rng = np.random.default_rng(0)
    log_in = np.log10(np.abs(S) + 1e-9).astype(np.float32)
    model_out = (log_in + 0.3 * rng.standard_normal(log_in.shape)).astype(np.float32)
    mask = np.zeros_like(log_in)
    mask[:, 10:17] = 1
    blended = model_out * mask + log_in * (1 - mask)
    ref = lr.istft(((10.0 ** blended) * np.exp(1j * np.angle(S))).astype(np.complex64), hop_length=192, win_length=384, n_fft=512)
    iw = inv_wss("hann", 384, 192, T, len(ref))
    y = emul.istft(192, win(384), iw, mag=model_out[None], phase=np.angle(S)[None], mag_domain=1,
                   blend_in=log_in[None], blend_mask=mask[None])
    assert relerr(y[0], ref) < 2 * TOL


def test_round_trip_snr(golden_clips):
    name = sorted(golden_clips)[0]
    x = golden_clips[name][:32000]
    S = emul.stft(x[None], 192, win(384), win_length=384)["spec"]
    T = S.shape[2]
    y = emul.istft(192, win(384), inv_wss("hann", 384, 192, T, 192 * (T - 1)), spec=S, win_length=384)[0]
    n = len(y)
    err = (x[512:n - 512] - y[512:n - 512]).astype(np.float64)
    snr = 10 * np.log10((x[512:n - 512].astype(np.float64) ** 2).sum() / (err ** 2).sum())
    assert snr >= 100.0, snr


@pytest.mark.parametrize("hop,wl", [(192, 384), (128, 512), (64, 256)])
def test_inverse_peak_tracking(hop, wl):
    """The overlap-add records max |y| per clip (fused peak normalisation, utils.py:84): equals the peak of the output."""
    x = noise(3, 20000 + 33, seed=hop)
    x[1] *= 3.0
    S = np.stack([lr.stft(x[b], n_fft=512, hop_length=hop, win_length=wl) for b in range(3)])
    T = S.shape[2]
    peaks = np.zeros(3, np.float32)
    y = emul.istft(hop, win(wl), inv_wss("hann", wl, hop, T, hop * (T - 1)), spec=S, win_length=wl, peaks=peaks)
    assert np.array_equal(peaks, np.abs(y).max(axis=1))


def test_randomised_geometry_forward_and_inverse():
    """Seeded sweep over clip lengths / hops / windows / `length=` around the tile and clip edges: the specialised paths
    (zero-tap pruning, compile-time overlap-add, output-pruned stage B) against the oracle."""
    rng = np.random.default_rng(20260101)
    for case in range(24):
        hop, wl = [(192, 384), (128, 512), (192, 320), (128, 384), (96, 384), (192, 512)][case % 6]
        L = int(rng.integers(512, 40000))
        x = noise(1, L, seed=1000 + case)
        ref_S = lr.stft(x[0], n_fft=512, hop_length=hop, win_length=wl)
        S = emul.stft(x, hop, win(wl), win_length=wl)["spec"]
        assert S[0].shape == ref_S.shape and relerr(S[0], ref_S) < TOL, (case, hop, wl, L)
        T = ref_S.shape[1]
        length = [0, 0, int(rng.integers(hop, hop * (T + 2))), L][case % 4]
        ref_y = lr.istft(ref_S, hop_length=hop, win_length=wl, n_fft=512, length=length or None)
        iw = inv_wss("hann", wl, hop, T, len(ref_y))
        y = emul.istft(hop, win(wl), iw, spec=ref_S[None], length=length, win_length=wl)
        good = iw < 100.0                      # window-sum-square (near) zero at a truncated tail: ill conditioned in the reference too
        assert relerr(y[0][good], ref_y[good]) < TOL, (case, hop, wl, L, length)


@pytest.mark.parametrize("L,hop,wl", [(6001, 192, 384), (16000, 192, 384), (9000, 128, 512), (20011, 250, 500)])
def test_forward_multi_clip_gaps_masks_and_unaligned_rows(L, hop, wl):
    """Several clips with per-clip gaps and frame masks; odd L makes the row pitch unaligned, so the waveform is staged
    by the threads instead of TMA -- the gap must still be exact (a separate zeroing pass used to race with that copy)."""
    B = 5
    x = noise(B, L, seed=L)
    rng = np.random.default_rng(L)
    g0 = rng.integers(0, L - 700, size=B)
    g0[1] = 0
    gaps_ = np.stack([g0, g0 + rng.integers(1, 700, size=B)], 1)
    gaps_[2] = [L - 3, L]                                 # the ragged tail of a clip
    T = 1 + L // hop
    frames = np.stack([rng.integers(0, T // 2, size=B), rng.integers(T // 2, T, size=B)], 1)
    a = emul.stft(x, hop, win(wl), gap_samples=gaps_, mask_frames=frames, mag_kind=2, want_spec=True, want_mask=True,
                  win_length=wl)
    for i in range(B):
        xg = x[i].copy()
        xg[gaps_[i, 0]:gaps_[i, 1]] = 0
        assert relerr(a["spec"][i], lr.stft(xg, n_fft=512, hop_length=hop, win_length=wl)) < TOL, (i, gaps_[i])
        m = np.zeros(T, np.float32)
        m[frames[i, 0]:frames[i, 1]] = 1
        assert np.array_equal(a["mask"][i], np.broadcast_to(m, (257, T)))


def test_epilogue_atan2_and_log1p_accuracy():
    """fast_atan2 / fast_log1p (phase and log1p outputs of the GAN front-end) against numpy on a dense sweep and on the
    special points: all four quadrants, the axes, signed zeros, tiny and huge arguments."""
    import ctypes as C
    rng = np.random.default_rng(5)
    n = 200000
    y = (rng.standard_normal(n) * 10.0 ** rng.uniform(-6, 3, n)).astype(np.float32)
    x = (rng.standard_normal(n) * 10.0 ** rng.uniform(-6, 3, n)).astype(np.float32)
    special = np.array([0.0, -0.0, 1.0, -1.0, 1e-30, -1e-30, 3e30, -3e30], np.float32)
    y[:64] = np.repeat(special, 8)
    x[:64] = np.tile(special, 8)
    a = np.empty(n, np.float32)
    l = np.empty(n, np.float32)
    FP = C.POINTER(C.c_float)
    emul.lib().emul_fast_math(n, y.ctypes.data_as(FP), x.ctypes.data_as(FP), a.ctypes.data_as(FP), l.ctypes.data_as(FP))
    ref_a = np.arctan2(y.astype(np.float64), x.astype(np.float64))
    assert np.abs(a - ref_a).max() < 5e-7                                   # ~1 ulp of pi
    assert np.array_equal(np.signbit(a[:64]), np.signbit(ref_a[:64].astype(np.float32)))
    ref_l = np.log1p(np.abs(x).astype(np.float64))
    assert (np.abs(l - ref_l) / np.maximum(ref_l, 1e-30)).max() < 1e-6


@pytest.mark.parametrize("hop,wl,L", [(192, 384, 6001), (128, 512, 3000), (192, 384, 300), (250, 500, 777)])
def test_forward_reflect_padding(hop, wl, L):
    """aip_stft_desc.center == 2: the centre padding mirrored about the first / last sample (np.pad mode="reflect", librosa < 0.10's
    default pad_mode) instead of zeros -- of the GAPPED clip, as the reference zeroes the gap before it calls librosa."""
    x = noise(3, L, seed=L)
    gaps = np.array([[0, min(L, 120)], [L - min(L, 150), L], [L // 3, L // 2]])
    out = emul.stft(x, hop, win(wl), center=2, win_length=wl, gap_samples=gaps, mag_kind=2, want_spec=True, vec_ok=(hop % 4 == 0))
    for b in range(3):
        xg = x[b].copy()
        xg[gaps[b, 0]:gaps[b, 1]] = 0
        ref = lr.stft(xg, n_fft=512, hop_length=hop, win_length=wl, pad_mode="reflect")
        assert out["spec"][b].shape == ref.shape and relerr(out["spec"][b], ref) < TOL, (b, relerr(out["spec"][b], ref))
        zero = lr.stft(xg, n_fft=512, hop_length=hop, win_length=wl)
        assert relerr(zero, ref) > 1e-2                          # the two paddings really differ on this signal


def test_fast_expm1_accuracy():
    """fast_expm1 (the GAN back-end's un-log, AIP_DOM_EXPM1) against numpy.expm1 in float64: the two branches and their seam."""
    import ctypes as C
    rng = np.random.default_rng(8)
    x = np.concatenate([rng.uniform(-20, 20, 100000), rng.uniform(-0.3, 0.3, 100000), 10.0 ** rng.uniform(-30, -1, 20000),
                        [0.0, -0.0, 0.25, -0.25, 0.2499999, 1e-38]]).astype(np.float32)
    out = np.empty_like(x)
    FP = C.POINTER(C.c_float)
    emul.lib().emul_fast_expm1(len(x), x.ctypes.data_as(FP), out.ctypes.data_as(FP))
    ref = np.expm1(x.astype(np.float64))
    nz = ref != 0
    assert (np.abs(out[nz] - ref[nz]) / np.abs(ref[nz])).max() < 1.5e-6
    assert np.array_equal(out[~nz], x[~nz])


@pytest.mark.parametrize("hop,wl", [(192, 384), (128, 512)])
def test_forward_power_spectrogram_variant_and_general_powers(hop, wl):
    """|S| ** p (the mel front-end's epilogue, utils.py:268-277): p = 2 is a straight-line variant (re^2 + im^2, no sqrt),
    any other p the general emitter; both against the oracle, with a crop and a gap."""
    x = noise(2, 9000, seed=hop)
    gaps = np.array([[1000, 2600], [8000, 9000]])
    for p in (2.0, 1.0, 0.5, 3.0):
        got = emul.stft(x, hop, win(wl), win_length=wl, gap_samples=gaps, mag_kind=4, power=p, want_spec=False, t_out=40)["mag"]
        for b in range(2):
            xg = x[b].copy()
            xg[gaps[b, 0]:gaps[b, 1]] = 0
            ref = np.abs(lr.stft(xg, n_fft=512, hop_length=hop, win_length=wl))[:, :40] ** p
            assert got[b].shape == ref.shape and relerr(got[b], ref) < (TOL if p <= 2 else 3 * TOL), (p, relerr(got[b], ref))


# ---- gap variants (aip_stft_gap_variants_f32; SURVEY 8f rank 3, models/CNNBLSTM/dataset.py:93-111) ----------------------
@pytest.mark.parametrize("hop,wl,L,g,t_out", [(192, 384, 16000, 3200, None), (192, 384, 6001, 700, None),
                                              (192, 384, 16000, 3200, 80), (128, 512, 9000, 1280, None),
                                              (64, 256, 6000, 500, 70), (250, 500, 20011, 1600, None),
                                              (192, 384, 16000, 9000, None)])
@pytest.mark.parametrize("vec_ok", [True, False])
def test_gap_variants_bit_identical_to_full_transforms(hop, wl, L, g, t_out, vec_ok):
    """Every variant = copy of the clean log-magnitude + re-transform of the frames the gap touches; that must equal the
    full transform of the gapped clip BIT FOR BIT (a missed frame would keep its clean value), for gaps at the clip start,
    the clip end, every phase against the tile grid, and of zero length."""
    N = 2
    x = noise(N, L, seed=hop + L)
    starts = np.unique(np.concatenate([np.arange(0, L - g, 211), [L - g - 1, L - g, 0, 1]]))
    G = len(starts)
    gaps = np.stack([starts, starts + g], 1)
    gaps[3] = [starts[3], starts[3]]          # a zero-length gap: the variant is the clean spectrogram
    gaps = np.concatenate([gaps, gaps[::-1]])  # row 1 takes them in the other order
    w = win(wl)
    var = emul.stft_variants(x, hop, w, G, gaps, mag_kind=2, t_out=t_out, vec_ok=vec_ok, win_length=wl, gap_len_max=g)
    full = emul.stft(np.repeat(x, G, 0), hop, w, gap_samples=gaps, mag_kind=2, t_out=t_out, want_spec=False,
                     vec_ok=vec_ok, win_length=wl)["mag"]
    assert not np.isnan(var["mag"]).any()
    assert np.array_equal(var["mag"], full)
    assert np.array_equal(var["mag"][3], var["clean_mag"][0])
    # and against the oracle for a few of them
    for v in (0, G // 2, G - 1, G + 1):
        xg = x[v // G].copy()
        xg[gaps[v, 0]:gaps[v, 1]] = 0
        ref = np.abs(lr.stft(xg, n_fft=512, hop_length=hop, win_length=wl))[:, :var["mag"].shape[2]]
        assert relerr(10.0 ** var["mag"][v].astype(np.float64), ref + 1e-9) < TOL


def test_gap_variants_other_magnitudes_and_getitem_shape():
    """|S| and log1p variants, and the dataset item against the oracle's restatement of __getitem__ (same np.random draws)."""
    L, G = 16000, 5
    x = noise(1, L, seed=9)
    w = win(384)
    np.random.seed(3)
    ref = cp.cnnblstm_getitem(x[0], gaps_per_audio=G, max_len_s=1.0)
    np.random.seed(3)
    starts = np.array([np.random.randint(0, L - 3200) for _ in range(G)])
    gaps = np.stack([starts, starts + 3200], 1)
    n_t = ref["spectrogram_gaps"].shape[2]
    for kind, fn in ((2, None), (1, lambda m: m), (3, np.log1p)):
        var = emul.stft_variants(x, 192, w, G, gaps, mag_kind=kind, t_out=n_t, win_length=384)["mag"]
        full = emul.stft(np.repeat(x, G, 0), 192, w, gap_samples=gaps, mag_kind=kind, t_out=n_t, want_spec=False,
                         win_length=384)["mag"]
        assert np.array_equal(var, full)
    var = emul.stft_variants(x, 192, w, G, gaps, mag_kind=2, t_out=n_t, win_length=384)["mag"]
    assert var.shape == ref["spectrogram_gaps"].shape
    lin_ref = 10.0 ** ref["spectrogram_gaps"].astype(np.float64)
    assert relerr(10.0 ** var.astype(np.float64), lin_ref) < TOL
