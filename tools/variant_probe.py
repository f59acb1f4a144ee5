"""Where the gap-variant transform kernel's per-tile time goes: the same 6400 tiles with random / sorted gap starts, one variant
per file, and the copy pass switched off (AIP_VAR_NO_FILL=1, profiling switch).  usage: variant_probe.py"""
import os, sys
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


L, g, T = 80000, 3200, 417
plan = sp.get_plan(512, 192, 384)
_keep = sp.experiment_env(AIP_VAR_NO_FILL="1")      # keep a reference: the switch lasts as long as the context object
_keep.__enter__()
for name, N, G, mode in (("256 files x 25, random starts", 256, 25, "random"), ("256 files x 25, sorted starts", 256, 25, "sorted"),
                         ("256 files x 25, all at 2.0 s", 256, 25, "fixed"), ("6400 files x 1, random", 6400, 1, "random"),
                         ("6400 files x 1, fixed", 6400, 1, "fixed")):
    x = (0.1 * torch.randn(N, L, device="cuda")).clamp_(-1, 1)
    rs = np.random.RandomState(0)
    if mode == "random":
        st = rs.randint(0, L - g, size=(N, G))
    elif mode == "sorted":
        st = np.sort(rs.randint(0, L - g, size=(N, G)), axis=1)
    else:
        st = np.full((N, G), 32000)
    st = st.reshape(-1)
    gaps = torch.as_tensor(np.stack([st, st + g], 1).astype(np.int32), device="cuda")
    clean = sp.stft(x, plan, mag_kind=sp.MAG_LOG10_EPS, t_out=T, want_spec=False)["mag"]
    out = torch.empty((N * G, 257, T), device="cuda")
    t = timeit(lambda: sp.stft_gap_variants(x, plan, gaps, G, t_out=T, clean_mag=clean, out=out, gap_len_max=g))
    print(f"{name:36s} {t * 1e3:7.1f} us  ({t * 1e3 * 148 / (N * G):5.2f} us/tile/SM)")
    del x, clean, out
