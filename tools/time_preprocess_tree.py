"""File-to-file throughput of the bulk preprocessing loop (reference pre_process_dataset.py:19-43) on a synthetic LibriSpeech-shaped
tree: N FLAC files of 6 - 15 s (16 kHz, mono, 16-bit) -> decode, pad / truncate to 5 s, random 0.1 s gap, peak-normalise, 16-bit
FLAC out.  Host decode / encode (native codec, one file per thread) around one device batch per 256 files."""
import os
import sys
import tempfile
import time
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from ml_audio_inpainting_b200 import audio_io, preprocess   # noqa: E402

n_files = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
rng = np.random.default_rng(0)
with tempfile.TemporaryDirectory() as td:
    src, dst = Path(td) / "in", Path(td) / "out"
    t0 = time.perf_counter()
    for i in range(n_files):
        d = src / f"{i % 16:03d}" / f"{i % 7:02d}"
        d.mkdir(parents=True, exist_ok=True)
        n = int(rng.integers(6 * 16000, 15 * 16000))
        t = np.arange(n)
        x = 6000 * np.sin(0.02 * t * (1 + i % 5)) * np.sin(1e-4 * t) + 300 * rng.standard_normal(n)
        (d / f"{i:06d}.flac").write_bytes(audio_io.encode_flac(np.rint(x).astype(np.int16), 16000))
    print(f"wrote {n_files} synthetic files in {time.perf_counter() - t0:.1f} s", flush=True)
    torch.zeros(1, device="cuda")
    for threads in (1, os.cpu_count() or 1):
        np.random.seed(0)
        t0 = time.perf_counter()
        n = preprocess.preprocess_tree(src, dst, progress=False, io_threads=threads)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        print(f"preprocess_tree, {threads:2d} I/O thread(s): {n} files in {dt:.2f} s = {n / dt:7.0f} files/s = {5 * n / dt:8.0f} audio-s/s "
              f"(the 72 000 files of configs[3]: {72000 / (n / dt) / 60:.1f} min)", flush=True)
