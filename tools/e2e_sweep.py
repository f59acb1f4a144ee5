"""Sweep of the host-buffer pipeline's chunk size / stream count (experiment; bench.py uses chunk 256, 3 streams)."""
import sys
import time
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from ml_audio_inpainting_b200 import frontend, spectral as sp  # noqa: E402

B, L = 4096, 160000
dev = torch.device("cuda:0")
plan = sp.get_plan(512, 192, 384, "hann", True, dev)
T, F = plan.num_frames(L), plan.n_bins
h_wave = torch.empty((B, L), dtype=torch.float32, pin_memory=True)
h_wave.normal_(0, 0.1)
h_out = torch.empty((B, F, T), dtype=torch.float32, pin_memory=True)
starts = np.random.RandomState(0).randint(0, L - 3200, size=B)
gaps_np = np.stack([starts, starts + 3200], 1).astype(np.int32)
for chunk, ns in [(256, 3), (128, 3), (512, 3), (1024, 3), (256, 2), (256, 4), (512, 4), (64, 4)]:
    pipe = frontend.HostPipeline(plan, B, L, chunk=chunk, n_streams=ns)
    pipe.logmag_gap(h_wave, gaps_np, h_out)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(3):
        pipe.logmag_gap(h_wave, gaps_np, h_out)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / 3
    print(f"chunk {chunk:5d} streams {ns}: {dt * 1e3:7.2f} ms  {B * 10 / dt / 1e3:7.1f} k audio-s/s  D2H {B * F * T * 4 / dt / 1e9:5.1f} GB/s", flush=True)
    del pipe
    torch.cuda.empty_cache()
