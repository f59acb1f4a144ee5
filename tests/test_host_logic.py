"""CPU tier: host-side product logic (index arithmetic, windows, sharding, codec, layout rules)."""
import ast
import re
from pathlib import Path

import numpy as np
import pytest

from ml_audio_inpainting_b200 import audio_io, gaps, sharding
from oracle import callers_port as cp
from oracle import librosa_port as lr
from oracle import utils_port as up

ROOT = Path(__file__).resolve().parents[1]
PKG = ROOT / "ml_audio_inpainting_b200"


def test_gap_indices_bit_exact_for_every_start():
    """Every possible gap start of a 5 s clip: frame ranges equal the oracle's float64 path (P1 and P2)."""
    L, sr = 80000, 16000
    for gap_s in (0.2, 0.08, 0.1):
        g = gaps.gap_len_samples(gap_s, sr)
        assert g == int(gap_s * sr)
        starts = np.arange(0, L - g + 1)
        f0, f1 = gaps.cnnblstm_frame_range(starts, g, sr, 192)
        t0, t1 = starts / sr, (starts + g) / sr
        assert np.array_equal(f0, lr.time_to_frames(t0, sr=sr, hop_length=192))
        assert np.array_equal(f1, lr.time_to_frames(t1, sr=sr, hop_length=192))
        a0, a1 = gaps.gan_frame_range(starts, starts + g, 128, 626)
        for s in starts[::997]:
            assert (int(a0[s]), int(a1[s])) == cp.gan_frame_mask_range(int(s), int(s + g), 128, 626)
    assert gaps.cnnblstm_frame_range(64320, 3200, 16000, 192) != (64320 // 192, (64320 + 3200) // 192)


def test_rng_consumption_matches_oracle():
    np.random.seed(3)
    a = gaps.draw_starts_exclusive(80000, 3200, 7)
    np.random.seed(3)
    b = [up.add_random_gap_from_audio(np.zeros(80000, np.float32), 0.2)[1][0] for _ in range(7)]
    assert np.array_equal(a / 16000, np.array(b))
    np.random.seed(4)
    c = gaps.draw_starts_inclusive(80000, 3200, 5)
    np.random.seed(4)
    d = [up.create_gap_mask(80000, 0.2, 16000)[1][0] for _ in range(5)]
    assert c.tolist() == d
    for args in [(1000, 0.0, 16000, None), (1000, 1.0, 16000, None), (80000, 0.08, 16000, 2.0)]:
        np.random.seed(0)
        s, e, kind = gaps.gap_mask_interval(*args)
        np.random.seed(0)
        assert (s, e) == up.create_gap_mask(*args)[1]


def test_frame_and_length_arithmetic():
    assert gaps.n_frames(80000, 512, 192) == 417 and gaps.n_frames(160000, 512, 192) == 834
    assert gaps.n_frames(80000, 512, 128) == 626 and gaps.n_frames(512, 512, 192, center=False) == 1
    assert gaps.istft_length(417, 512, 192) == 79872 and gaps.istft_length(626, 512, 128) == 80000
    assert gaps.istft_length(417, 512, 192, length=80000) == 80000
    assert gaps.cnnblstm_crop_frames(16000, 5.0, 192) == 417
    with pytest.raises(ValueError):
        gaps.n_frames(100, 512, 192, center=False)


def test_fft_window_matches_oracle():
    pytest.importorskip("torch")
    from ml_audio_inpainting_b200.spectral import fft_window
    for name, wl in [("hann", 384), ("hann", 512), ("hamming", 400), ("blackman", 512)]:
        assert np.array_equal(fft_window(name, wl, 512), lr.fft_window(name, wl, 512))
    with pytest.raises(ValueError):
        fft_window("hann", 1024, 512)


def test_shard_bounds_partition():
    for n in (0, 1, 7, 72000, 4096):
        for ws in (1, 2, 3, 4, 8):
            spans = [sharding.shard_bounds(n, r, ws) for r in range(ws)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(ws - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        sharding.shard_bounds(10, 2, 2)


def test_flac_and_wav_codec_round_trip(tmp_path, golden_clips):
    x = golden_clips[sorted(golden_clips)[0]][:20000]
    pcm = np.round(x * 32768).astype(np.int16)
    blob = audio_io.encode_flac(pcm, 16000)
    dec, info = audio_io.decode_flac(blob, verify_md5=True)
    assert info.sample_rate == 16000 and info.channels == 1 and np.array_equal(np.asarray(dec).reshape(-1), pcm)
    audio_io.write_audio(tmp_path / "a.flac", x, 16000, "flac")
    y, sr = audio_io.read_audio(tmp_path / "a.flac")
    assert sr == 16000 and np.abs(y.reshape(-1) - x).max() <= 1.0 / 32767
    audio_io.write_audio(tmp_path / "a.wav", x, 16000, "wav")
    z, sr = audio_io.read_audio(tmp_path / "a.wav")
    assert sr == 16000 and np.abs(z.reshape(-1) - x).max() <= 1.0 / 32767
    with pytest.raises(Exception):
        audio_io.read_audio(tmp_path / "missing.flac")


def test_product_never_imports_the_oracle():
    """The oracle is test infrastructure: nothing under the package (nor the drop-in) may import it."""
    for py in PKG.rglob("*.py"):
        tree = ast.parse(py.read_text())
        for node in ast.walk(tree):
            names = []
            if isinstance(node, ast.Import):
                names = [a.name for a in node.names]
            elif isinstance(node, ast.ImportFrom):
                names = [node.module or ""]
            assert not any(n == "oracle" or n.startswith("oracle.") for n in names), py
    for src in [q for q in (PKG / "csrc").glob("*") if q.is_file()]:
        assert "oracle" not in src.read_text(errors="ignore").replace("the oracle", ""), src
    assert "TEST INFRASTRUCTURE ONLY" in (ROOT / "oracle" / "librosa_port.py").read_text()
    hdr = (ROOT / "oracle" / "__init__.py").read_text()
    assert "TEST INFRASTRUCTURE ONLY" in hdr and "PINNED TO VALUES THE REFERENCE ITSELF PRODUCED" in hdr and "unpinned" in hdr


def test_bench_and_entry_only_use_oracle_as_checker():
    bench = (ROOT / "bench.py").read_text()
    # the cpu_baseline / reference arm only: the two CPU workers (the port, and the torch.stft baseline's window table)
    assert bench.count("from oracle") == 2 and "_cpu_worker" in bench and "_torch_cpu_worker" in bench
    for m in re.finditer(r"from oracle", bench):
        head = bench[:m.start()]
        assert head.rfind("def _cpu_worker") > head.rfind("def main") or head.rfind("def _torch_cpu_worker") > head.rfind("def main")
    entry = (ROOT / "__graft_entry__.py").read_text()
    assert entry.count("from oracle") == 1 and "def smoke" in entry and "def build" in entry
    assert re.search(r"/root/reference", bench) is None and re.search(r"/root/reference", entry) is None


def test_mel_bands_cover_every_nonzero_weight():
    """aip_mel_project_f32 sums each mel row over a host-computed bin range only: the range must hold every non-zero weight."""
    from ml_audio_inpainting_b200.spectral import mel_bands, mel_basis
    for args in ((16000, 2048, 128), (16000, 512, 128), (22050, 1024, 40)):
        w = mel_basis(*args)
        b = mel_bands(w)
        assert b.dtype == np.int32 and b.shape == (args[2], 2)
        for m in range(args[2]):
            inside = np.zeros(w.shape[1], bool)
            inside[b[m, 0]:b[m, 1]] = True
            assert np.all(w[m][~inside] == 0)
        # a bin belongs to at most two neighbouring triangles: the banded contraction reads the spectrogram about twice
        assert int(((w != 0).sum(0)).max()) <= 2 and int((b[:, 1] - b[:, 0]).sum()) <= 2 * w.shape[1] + args[2]
    z = np.zeros((3, 9), np.float32)
    z[1, 4] = 1
    assert mel_bands(z).tolist() == [[0, 0], [4, 5], [0, 0]]


def test_experiment_env_sets_reloads_and_restores(monkeypatch):
    """AIP_* switches are read once when the library loads; experiment_env sets them, makes the library re-read, and restores."""
    import os
    from ml_audio_inpainting_b200 import spectral
    monkeypatch.setenv("AIP_FWD_CHUNK", "7")
    monkeypatch.delenv("AIP_INV_TMA", raising=False)
    with spectral.experiment_env(AIP_FWD_CHUNK=None, AIP_INV_TMA="0"):
        assert "AIP_FWD_CHUNK" not in os.environ and os.environ["AIP_INV_TMA"] == "0"
    assert os.environ["AIP_FWD_CHUNK"] == "7" and "AIP_INV_TMA" not in os.environ


def test_write_audio_takes_device_quantised_int16(tmp_path):
    """An int16 array is the file's samples as they are (quantised by aip_wave_to_pcm16_f32 on the device), for both containers;
    a float array goes through the host conversion and gives the same file when it holds the same samples."""
    from ml_audio_inpainting_b200 import audio_io
    rng = np.random.default_rng(0)
    x = np.clip(0.4 * rng.standard_normal(5000), -1, 1).astype(np.float32)
    q = audio_io._to_int16(x, 32768.0)
    audio_io.write_audio(tmp_path / "a.flac", q, 16000, "flac")
    audio_io.write_audio(tmp_path / "b.flac", x, 16000, "flac")
    assert (tmp_path / "a.flac").read_bytes() == (tmp_path / "b.flac").read_bytes()
    back, sr = audio_io.read_audio(tmp_path / "a.flac")
    assert sr == 16000 and np.array_equal(np.rint(back[:, 0] * 32768).astype(np.int16) if back.ndim == 2 else
                                          np.rint(back * 32768).astype(np.int16), q)
    audio_io.write_audio(tmp_path / "c.wav", q, 16000, "wav")
    back, _ = audio_io.read_audio(tmp_path / "c.wav")
    assert np.array_equal(np.rint(np.asarray(back).reshape(-1) * 32768).astype(np.int16), q)


def test_default_pad_mode_switch(monkeypatch):
    """spectral.DEFAULT_PAD_MODE: "constant" unless AIP_LIBROSA_PAD_MODE says otherwise at import time."""
    import importlib
    import ml_audio_inpainting_b200.spectral as spmod
    assert spmod.DEFAULT_PAD_MODE == "constant"
    monkeypatch.setenv("AIP_LIBROSA_PAD_MODE", "reflect")
    importlib.reload(spmod)
    assert spmod.DEFAULT_PAD_MODE == "reflect"
    monkeypatch.delenv("AIP_LIBROSA_PAD_MODE")
    importlib.reload(spmod)
    assert spmod.DEFAULT_PAD_MODE == "constant"
