"""The reference's OWN caller source as the checker.

tests/golden/reference_callers.npz holds what the reference's unmodified ``LibriSpeechDataset.__getitem__``
(models/CNNBLSTM/dataset.py:74-121), ``SpeechInpaintingDataset.__getitem__`` (models/GAN/dataset.py:63-166),
``model_eval.inpaint`` (models/model_eval.py:48-194, both branches, random-init models with captured outputs) and the
``pre_process_dataset.py:36-41`` loop body produce on the test_samples clips (tests/golden/make_reference_callers.py; the
three absent third-party packages are stood in for by tests/refshim.py).

  CPU tier   * in the build container (where /root/reference exists) the reference source is RE-RUN and must reproduce
               the committed fixture bit for bit;
             * the oracle's restatement of those callers (oracle/callers_port.py) must reproduce it bit for bit too --
               that is what entitles the other tests to use the restatement as their checker.
  GPU tier   * the B200 front-ends / back-end (through the C ABI) against the fixture: integer facts bit-exact, spectra
               within 1e-4 relative max-abs, written PCM within 1 LSB.
"""
from pathlib import Path

import numpy as np
import pytest

from oracle import callers_port as cp
from oracle import utils_port as up
from tests import refshim
from tests.golden import make_reference_callers as mk

TOL = 1e-4
GOLD = Path(__file__).parent / "golden"


@pytest.fixture(scope="module")
def fx():
    z = np.load(GOLD / "reference_callers.npz")
    return {k: z[k] for k in z.files}


def sampled(fx, key, arr):
    """(fixture sample, the same positions of ``arr``, fixture |.|-sum, |arr|-sum)."""
    a = np.asarray(arr)
    assert list(a.shape) == fx[key + "/shape"].tolist(), key
    return fx[key + "/val"], a.reshape(-1)[fx[key + "/idx"]], float(fx[key + "/abssum"]), float(np.abs(a).astype(np.float64).sum())


def _pcm_close(pcm, want):
    """written PCM-16: never more than one LSB apart, and that on a small fraction of the samples only.  A sample flips to the
    neighbouring code when x * 32768 lies within the arithmetic's error of a rounding boundary: the reference computes in
    float64 (peak normalisation of a float64 array), the CUDA path in fp32 (relative error 6e-8 -> 2e-3 LSB at full scale ->
    ~0.4 % of the samples per rounding step).  Measured: oracle vs the reference's torch-float32 blend <= 0.2 %, CUDA 0.4 - 1.3 %."""
    d = np.abs(pcm.astype(np.int64) - want.astype(np.int64))
    assert d.max() <= 1 and (d > 0).mean() < 2.5e-2, (int(d.max()), float((d > 0).mean()))


def relerr(a, b):
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))


@pytest.mark.skipif(not refshim.reference_available(), reason="the reference checkout only exists in the build container")
def test_fixture_is_what_the_reference_source_produces(fx):
    again = mk.generate()
    assert sorted(again) == sorted(fx)
    for k in fx:
        assert np.array_equal(again[k], fx[k]), k


@pytest.mark.skipif(not refshim.reference_available(), reason="the reference checkout only exists in the build container")
def test_configs4_model_is_the_references_architecture():
    """tools/cnnblstm_model.py (the model between front- and back-end in bench.py's cnnblstm_e2e leg) takes the state_dict of the
    reference's own StackedBLSTMCNN(cnn_blstm.yaml) unchanged and computes the same reconstruct_spectrogram, bit for bit."""
    import torch
    from tools.cnnblstm_model import StandInBLSTMCNN
    with refshim.reference_modules() as ref:
        torch.manual_seed(0)
        theirs = ref.cnnblstm_model.StackedBLSTMCNN(str(refshim.REFERENCE / "models" / "CNNBLSTM" / "cnn_blstm.yaml")).eval()
        ours = StandInBLSTMCNN().eval()
        ours.load_state_dict(theirs.state_dict(), strict=True)
        x = torch.randn(2, 257, 48)
        mask = torch.zeros(2, 257, 48)
        mask[:, :, 10:17] = 1
        with torch.no_grad():
            assert torch.equal(theirs.reconstruct_spectrogram(x, mask), ours.reconstruct_spectrogram(x, mask))
            assert torch.equal(theirs(x.unsqueeze(1)), ours(x.unsqueeze(1)))


def test_oracle_caller_restatement_equals_the_reference_source(fx, golden_clips):
    """oracle/callers_port.py + utils_port.py against what the reference's own source produced (bit for bit: both sit on
    the same librosa restatement, so any difference would be a mistake in the restated caller logic)."""
    # CNNBLSTM dataset items
    np.random.seed(mk.CNN_SEED)
    for i, name in enumerate(fx["cnn/files"]):
        it = cp.cnnblstm_getitem(golden_clips[str(name)], gaps_per_audio=mk.CNN_GAPS)
        for key, arr in (("spectrogram_gaps", it["spectrogram_gaps"]), ("spectrogram_target_phases", it["spectrogram_target_phases"])):
            want, got, s0, s1 = sampled(fx, f"cnn/{i}/{key}", arr)
            assert np.array_equal(want, got) and s0 == s1
        assert np.array_equal(fx[f"cnn/{i}/gap_ints"], it["gap_ints"])
        assert np.array_equal(fx[f"cnn/{i}/gap_frames"], it["gap_frames"])
    assert int(fx["cnn/rng_after"]) == np.random.randint(0, 1 << 30)
    # GAN dataset items
    names = [str(n) for n in fx["names"]]
    np.random.seed(mk.GAN_SEED)
    for i in range(mk.GAN_FILES):
        it = cp.gan_item(golden_clips[names[i]])
        for key in ("original_magnitude", "impaired_magnitude", "original_phase"):
            want, got, s0, s1 = sampled(fx, f"gan/{i}/{key}", it[key])
            assert np.array_equal(want, got) and s0 == s1
        assert fx[f"gan/{i}/gap_frames"].tolist() == list(it["gap_frames"])
    assert int(fx["gan/rng_after"]) == np.random.randint(0, 1 << 30)
    # model_eval.inpaint, CNN-BLSTM branch: blend with the captured model output, 10 **, phase reuse, save_audio
    from ml_audio_inpainting_b200 import audio_io
    for i in range(mk.EVAL_CNN_CLIPS):
        ev = cp.eval_frontend_cnnlstm(golden_clips[names[i]])
        log_imp = ev["log_impaired_magnitude"].astype(np.float32)
        model_out = log_imp.copy()
        model_out[:, 166:173] = fx[f"eval_cnn/{i}/model_out_gap"]
        blended = model_out * ev["mask"] + log_imp * (1 - ev["mask"])            # models/CNNBLSTM/model.py:108
        y = cp.eval_backend((10 ** blended).astype(np.float32), ev["original_phase"])
        pcm = audio_io._to_int16(up.peak_normalize(y), 32768.0)
        _pcm_close(pcm, fx[f"eval_cnn/{i}/pcm"])      # the script blends and takes 10 ** in torch float32: last-bit differences
    # GAN branch: spectrogram_to_audio(generator output, phase=original)
    for i in range(mk.EVAL_GAN_CLIPS):
        ev = cp.eval_frontend_gan(golden_clips[names[i]])
        y = cp.eval_backend(fx[f"eval_gan/{i}/generator_out"], ev["original_phase"], hop_length=128, win_length=512)
        pcm = audio_io._to_int16(up.peak_normalize(y), 32768.0)
        _pcm_close(pcm, fx[f"eval_gan/{i}/pcm"])
    # pre_process_dataset.py loop body
    np.random.seed(mk.PRE_SEED)
    for i in range(mk.PRE_FILES):
        y, iv = up.add_random_gap_from_audio(golden_clips[names[i]], 0.1)
        assert np.array_equal(np.array(iv), fx[f"pre/{i}/gap_int_s"])
        assert np.array_equal(audio_io._to_int16(up.peak_normalize(y), 32768.0), fx[f"pre/{i}/pcm"])


# ------------------------------------------------------------------------------------------------------ GPU tier

@pytest.mark.gpu
def test_cuda_dataset_front_ends_match_the_reference_source(fx, golden_clips):
    import torch
    from ml_audio_inpainting_b200 import frontend
    names = [str(n) for n in fx["names"]]
    # LibriSpeechDataset.__getitem__: same np.random stream, same draws, same frames; spectra within 1e-4
    files = [str(n) for n in fx["cnn/files"]]
    wave = torch.from_numpy(np.stack([golden_clips[n] for n in files])).cuda()
    np.random.seed(mk.CNN_SEED)
    item = frontend.cnnblstm_dataset_batch(wave, gaps_per_audio=mk.CNN_GAPS)
    assert int(fx["cnn/rng_after"]) == np.random.randint(0, 1 << 30)
    for i in range(len(files)):
        want, got, s0, s1 = sampled(fx, f"cnn/{i}/spectrogram_gaps", item["spectrogram_gaps"][i].cpu().numpy())
        floor = want > -8.0                                                  # away from the eps floor compare linearly
        assert relerr(10.0 ** got[floor].astype(np.float64), 10.0 ** want[floor].astype(np.float64)) < TOL
        assert np.abs(got - want).max() < 0.35                               # at the floor: |S| ~ 1e-9 +- fp32 noise
        want, got, s0, s1 = sampled(fx, f"cnn/{i}/spectrogram_target_phases",
                                    item["spectrogram_target_phases"][i].cpu().numpy())
        assert relerr(got, want) < TOL and abs(s1 - s0) / s0 < TOL
        assert np.array_equal(fx[f"cnn/{i}/gap_ints"], item["gap_ints"][i])
        assert np.array_equal(fx[f"cnn/{i}/gap_frames"], item["gap_frames"][i])
        m = item["gap_masks"][i].cpu().numpy()
        for j, (f0, f1) in enumerate(fx[f"cnn/{i}/gap_frames"]):
            assert mk.mask_range(m[j], 1) == (f0, f1)
    # SpeechInpaintingDataset.__getitem__
    wave = torch.from_numpy(np.stack([golden_clips[n] for n in names[:mk.GAN_FILES]])).cuda()
    np.random.seed(mk.GAN_SEED)
    g = frontend.gan_batch(wave, gap_len_s=0.2)
    assert int(fx["gan/rng_after"]) == np.random.randint(0, 1 << 30)
    for i in range(mk.GAN_FILES):
        for key in ("original_magnitude", "impaired_magnitude"):
            want, got, s0, s1 = sampled(fx, f"gan/{i}/{key}", g[key][i].cpu().numpy())
            assert relerr(np.expm1(got.astype(np.float64)), np.expm1(want.astype(np.float64))) < TOL
        want, got, _, _ = sampled(fx, f"gan/{i}/original_phase", g["original_phase"][i].cpu().numpy())
        wm, _, _, _ = sampled(fx, f"gan/{i}/original_magnitude", g["original_magnitude"][i].cpu().numpy())
        w = np.expm1(wm.astype(np.float64))                                      # phase as a magnitude-weighted phasor
        assert np.abs(w * (np.exp(1j * got) - np.exp(1j * want))).max() / w.max() < TOL
        assert mk.mask_range(g["mask"][i].cpu().numpy(), 0) == tuple(fx[f"gan/{i}/gap_frames"])


@pytest.mark.gpu
def test_cuda_model_eval_and_bulk_loop_match_the_reference_source(fx, golden_clips):
    """model_eval.inpaint's two branches and the pre_process_dataset loop body, fed the captured model outputs, down to the
    PCM the reference's save_audio wrote."""
    import torch
    from ml_audio_inpainting_b200 import audio_io, frontend, preprocess, spectral as sp
    names = [str(n) for n in fx["names"]]
    n = mk.EVAL_CNN_CLIPS
    wave = torch.from_numpy(np.stack([golden_clips[k] for k in names[:n]])).cuda()
    ev = frontend.eval_cnnlstm_batch(wave)
    model_out = ev["log_impaired_magnitude"].clone()
    model_out[:, :, 166:173] = torch.from_numpy(np.stack([fx[f"eval_cnn/{i}/model_out_gap"] for i in range(n)])).cuda()
    plan = sp.get_plan(512, 192, 384, "hann", True, wave.device)
    y = sp.istft_blend(plan, model_out, ev["log_impaired_magnitude"], ev["mask"], ev["original_phase"],
                       mag_domain=sp.DOM_POW10, normalize=True).cpu().numpy()
    for i in range(n):
        _pcm_close(audio_io._to_int16(y[i], 32768.0), fx[f"eval_cnn/{i}/pcm"])
    # GAN branch (model_eval.py:118-140): the generator output goes to spectrogram_to_audio as it is, original phase
    n = mk.EVAL_GAN_CLIPS
    wave = torch.from_numpy(np.stack([golden_clips[k] for k in names[:n]])).cuda()
    eg = frontend.eval_gan_batch(wave)
    gen = torch.from_numpy(np.stack([fx[f"eval_gan/{i}/generator_out"] for i in range(n)])).cuda()
    plan2 = sp.get_plan(512, 128, 512, "hann", True, wave.device)
    y = sp.istft(plan2, mag=gen, phase=eg["original_phase"], db_auto=True, normalize=True).cpu().numpy()
    for i in range(n):
        _pcm_close(audio_io._to_int16(y[i], 32768.0), fx[f"eval_gan/{i}/pcm"])
    # pre_process_dataset.py:36-41
    n = mk.PRE_FILES
    wave = torch.from_numpy(np.stack([golden_clips[k] for k in names[:n]])).cuda()
    np.random.seed(mk.PRE_SEED)
    res = preprocess.preprocess_batch(wave, gap_len=0.1)
    out = res["audio_gap_normalized"].cpu().numpy()
    for i in range(n):
        assert np.array_equal(res["gap_int_s"][i], fx[f"pre/{i}/gap_int_s"])
        _pcm_close(audio_io._to_int16(out[i], 32768.0), fx[f"pre/{i}/pcm"])
