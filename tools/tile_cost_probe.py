"""Per-tile cost of the forward kernel when every tile starts a new clip: B short clips of exactly one 32-frame tile
(plain kernel, with / without a gap) against the same number of tiles cut from long clips.  usage: tile_cost_probe.py [tiles]"""
import sys
from pathlib import Path
import numpy as np, torch
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from ml_audio_inpainting_b200 import spectral as sp


def timeit(fn, n=20, w=5):
    for _ in range(w): fn()
    torch.cuda.synchronize(); e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


n_tiles = int(sys.argv[1]) if len(sys.argv) > 1 else 6400
plan = sp.get_plan(512, 192, 384)
for name, L, tiles_per_clip in (("one-tile clips (L = 5952, T = 32)", 31 * 192, 1), ("two-tile clips (T = 64)", 63 * 192, 2),
                                ("5 s clips (T = 417, 14 tiles)", 80000, 14)):
    B = max(1, n_tiles // tiles_per_clip)
    x = (0.1 * torch.randn(B, L, device="cuda")).clamp_(-1, 1)
    T = plan.num_frames(L)
    out = {"mag": torch.empty((B, 257, T), device="cuda")}
    st = np.random.RandomState(0).randint(0, max(1, L - 3200), size=B)
    gaps = torch.as_tensor(np.stack([st, st + 3200], 1).astype(np.int32), device="cuda")
    t0 = timeit(lambda: sp.stft(x, plan, mag_kind=sp.MAG_LOG10_EPS, want_spec=False, out=out))
    t1 = timeit(lambda: sp.stft(x, plan, gap_samples=gaps, mag_kind=sp.MAG_LOG10_EPS, want_spec=False, out=out))
    nt = B * tiles_per_clip
    print(f"{name:38s} {nt:6d} tiles: no gap {t0 * 1e3:7.1f} us ({t0 * 1e3 * 148 / nt:5.2f} us/tile/SM)   with gap {t1 * 1e3:7.1f} us "
          f"({t1 * 1e3 * 148 / nt:5.2f} us/tile/SM)")
