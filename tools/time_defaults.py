"""utils.spectrogram_to_audio's own default geometry (n_fft 512, hop_length 512: frames that do not overlap, utils.py:279-284)
next to the geometry the reference's scripts pass (hop 192 / win 384): inverse and Griffin-Lim 64, batch of 5-s clips."""
import sys
from pathlib import Path
import torch
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

B, L = 1024, 80000
x = (0.1 * torch.randn(B, L, device="cuda")).clamp_(-1, 1)
for hop, win in ((512, 512), (192, 384)):
    plan = sp.get_plan(512, hop, win)
    S = sp.stft(x, plan)["spec"]
    T = S.shape[2]
    y = torch.empty((B, plan.istft_length(T)), device="cuda")
    ti = timeit(lambda: sp.istft(plan, spec=S, out=y))
    mag = S.abs()
    tg = timeit(lambda: sp.griffinlim(plan, mag, n_iter=64), n=3)
    by = B * (8 * 257 * T + 4 * y.shape[1])
    print(f"hop {hop} win {win}: T {T}  istft {ti:7.3f} ms ({by / ti / 1e6:6.0f} GB/s)   griffinlim 64 it {tg:8.2f} ms", flush=True)
