"""Times the dataset-shaped collate (models/CNNBLSTM/dataset.py:74-121, gaps_per_audio = 25) on one GPU:
G full gapped transforms per file (what the reference's loop amounts to) against aip_stft_gap_variants_f32
(one clean transform per file, streaming copy, re-transform of the frames each gap touches).
usage: time_variants.py [N files] [G]"""
import sys
from pathlib import Path
import numpy as np, torch
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from ml_audio_inpainting_b200 import frontend, spectral as sp


def timeit(fn, n=10, w=3):
    for _ in range(w): fn()
    torch.cuda.synchronize(); e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


N = int(sys.argv[1]) if len(sys.argv) > 1 else 256
G = int(sys.argv[2]) if len(sys.argv) > 2 else 25
L, g = 80000, 3200
x = (0.1 * torch.randn(N, L, device="cuda")).clamp_(-1, 1)
np.random.seed(0)
starts = np.random.randint(0, L - g, size=N * G)
sam = np.stack([starts, starts + g], 1)
gaps = torch.as_tensor(sam.astype(np.int32), device="cuda")
plan = sp.get_plan(512, 192, 384)
T, F = 417, 257
xr = x.repeat_interleave(G, 0)
full_out = {"mag": torch.empty((N * G, F, T), device="cuda")}
var_out = torch.empty((N * G, F, T), device="cuda")
clean = sp.stft(x, plan, mag_kind=sp.MAG_LOG10_EPS, t_out=T, want_spec=False)["mag"]
out_bytes = N * G * F * T * 4
t_full = timeit(lambda: sp.stft(xr, plan, gap_samples=gaps, mag_kind=sp.MAG_LOG10_EPS, t_out=T, want_spec=False, out=full_out))
t_var = timeit(lambda: sp.stft_gap_variants(x, plan, gaps, G, t_out=T, clean_mag=clean, out=var_out, gap_len_max=g))
t_var_all = timeit(lambda: sp.stft_gap_variants(x, plan, gaps, G, t_out=T, out=var_out, gap_len_max=g))
assert torch.equal(full_out["mag"], var_out)
print(f"N={N} files x G={G} gaps, 5 s clips, P1, log10 magnitudes [{N * G}, {F}, {T}] = {out_bytes / 1e9:.3f} GB written")
print(f"G full gapped transforms (pre-repeated waves)      {t_full:8.3f} ms   {out_bytes / t_full / 1e6:8.1f} GB/s of output")
print(f"variants (clean magnitudes given)                  {t_var:8.3f} ms   {out_bytes / t_var / 1e6:8.1f} GB/s of output   x{t_full / t_var:.2f}")
print(f"variants incl. the clean transform                 {t_var_all:8.3f} ms   {out_bytes / t_var_all / 1e6:8.1f} GB/s of output   x{t_full / t_var_all:.2f}")
t_item = timeit(lambda: frontend.cnnblstm_dataset_batch(x, gaps_per_audio=G, starts=starts), n=5)
t_naive = timeit(lambda: frontend.cnnblstm_batch(x.repeat_interleave(G, 0), starts=starts, want_target=True), n=5)
print(f"frontend.cnnblstm_dataset_batch (item: mags, masks, target view)   {t_item:8.3f} ms")
print(f"frontend.cnnblstm_batch on G repeated rows (mags, masks, G targets) {t_naive:8.3f} ms   x{t_naive / t_item:.2f}")
print(f"files/s: {N / t_item * 1e3:,.0f}   (reference CPU loop: 2 decodes + 2 STFTs per gap)")
# per-kernel times of one item call (CUPTI through torch.profiler; shares only, the numbers above are the clean timings)
try:
    from torch.profiler import profile, ProfilerActivity
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for _ in range(3):
            frontend.cnnblstm_dataset_batch(x, gaps_per_audio=G, starts=starts)
        torch.cuda.synchronize()
    for ev in sorted(prof.key_averages(), key=lambda e: -e.device_time_total)[:8]:
        print(f"  {ev.key[:90]:90s} {ev.device_time_total / ev.count / 1e3:8.3f} ms x{ev.count}")
except Exception as exc:      # profiling is optional
    print("profiler unavailable:", exc)
