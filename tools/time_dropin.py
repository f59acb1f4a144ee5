"""Per-call latency of the drop-in utils.py (one 5-s clip per call, numpy in / numpy out, as the reference's scripts use it)
next to the CPU oracle (numpy + scipy.fft = what librosa runs) on the same host.  Wall clock, median of N calls."""
import statistics
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "ml_audio_inpainting_b200" / "dropin"))
import utils                                    # noqa: E402  the drop-in
from oracle import librosa_port as lr           # noqa: E402  timing baseline only
from oracle import utils_port as up             # noqa: E402


def med(fn, n=30, warm=3):
    for _ in range(warm):
        fn()
    ts = []
    for _ in range(n):
        t0 = time.perf_counter()
        fn()
        ts.append(time.perf_counter() - t0)
    return 1e3 * statistics.median(ts)


rng = np.random.default_rng(0)
x = np.clip(0.1 * rng.standard_normal(80000), -1, 1).astype(np.float32)
kw = dict(n_fft=512, hop_length=192, win_length=384)
S = utils.extract_spectrogram(x, **kw)
mag, ph = np.abs(S), np.angle(S)
rows = [
    ("extract_spectrogram", lambda: utils.extract_spectrogram(x, **kw), lambda: up.extract_spectrogram(x, **kw)),
    ("spectrogram_to_audio(complex)", lambda: utils.spectrogram_to_audio(S, phase_info=True, **kw),
     lambda: up.spectrogram_to_audio(S, phase_info=True, **kw)),
    ("spectrogram_to_audio(mag, phase)", lambda: utils.spectrogram_to_audio(mag, phase=ph, **kw),
     lambda: up.spectrogram_to_audio(mag, phase=ph, **kw)),
    ("spectrogram_to_audio(griffinlim 32)", lambda: utils.spectrogram_to_audio(mag, n_iter=32, **kw),
     lambda: up.spectrogram_to_audio(mag, n_iter=32, **kw)),
    ("extract_mel_spectrogram", lambda: utils.extract_mel_spectrogram(x, 16000, 512, 192, 128),
     lambda: lr.melspectrogram(y=x, sr=16000, n_fft=512, hop_length=192, n_mels=128)),
]
Sd = utils.extract_spectrogram(x)                      # the functions' own defaults: n_fft 2048, hop 512
rows += [
    ("extract_spectrogram (defaults 2048/512)", lambda: utils.extract_spectrogram(x), lambda: up.extract_spectrogram(x)),
    ("spectrogram_to_audio(complex, 2048/512)", lambda: utils.spectrogram_to_audio(Sd, phase_info=True, n_fft=2048, hop_length=512),
     lambda: up.spectrogram_to_audio(Sd, phase_info=True, n_fft=2048, hop_length=512)),
    ("extract_mel_spectrogram (defaults)", lambda: utils.extract_mel_spectrogram(x), lambda: lr.melspectrogram(y=x, sr=16000)),
]
import tempfile
from ml_audio_inpainting_b200 import audio_io
with tempfile.TemporaryDirectory() as td:
    f = Path(td) / "clip.flac"
    t10 = np.arange(160000)
    audio_io.write_audio(f, (0.3 * np.sin(0.05 * t10) * np.sin(3e-4 * t10) + 0.01 * rng.standard_normal(160000)).astype(np.float32), 16000)
    print(f"{'load_audio (10 s FLAC -> 5 s)':38s} drop-in {med(lambda: utils.load_audio(f)):8.3f} ms", flush=True)
    print(f"{'save_audio (5 s, normalize, FLAC)':38s} drop-in {med(lambda: utils.save_audio(x, Path(td) / 'o.flac')):8.3f} ms", flush=True)
for name, gpu, cpu in rows:
    n = 10 if "griffin" in name else 30
    print(f"{name:38s} drop-in {med(gpu, n):8.3f} ms   oracle (1 core) {med(cpu, max(3, n // 3), 1):8.3f} ms", flush=True)
