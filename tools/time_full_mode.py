"""The general-epilogue forward variant (FWD_FULL: any combination of outputs that has no straight-line instantiation) against
the same outputs produced by separate specialised launches."""
import sys
from pathlib import Path
import numpy as np, torch
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from ml_audio_inpainting_b200 import spectral as sp

def timeit(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n

B, L = 1024, 160000
x = (0.1 * torch.randn(B, L, device="cuda")).clamp_(-1, 1)
plan = sp.get_plan(512, 192, 384)
T = plan.num_frames(L)
o_spec = {"spec": torch.empty((B, 257, T), dtype=torch.complex64, device="cuda")}
o_mag = {"mag": torch.empty((B, 257, T), device="cuda")}
o_both = {**o_spec, **o_mag}
t_spec = timeit(lambda: sp.stft(x, plan, out=o_spec))
t_mag = timeit(lambda: sp.stft(x, plan, mag_kind=sp.MAG_LOG10_EPS, want_spec=False, out=o_mag))
t_both = timeit(lambda: sp.stft(x, plan, mag_kind=sp.MAG_LOG10_EPS, want_spec=True, out=o_both))
t_pow3 = timeit(lambda: sp.stft(x, plan, mag_kind=sp.MAG_POW, power=3.0, want_spec=False, out=o_mag))
t_pow2 = timeit(lambda: sp.stft(x, plan, mag_kind=sp.MAG_POW, power=2.0, want_spec=False, out=o_mag))
print(f"complex {t_spec:.3f} ms, log10 {t_mag:.3f} ms, both in one FWD_FULL launch {t_both:.3f} ms (two launches {t_spec + t_mag:.3f}); "
      f"|S|^3 (FWD_FULL) {t_pow3:.3f} ms, |S|^2 {t_pow2:.3f} ms")
