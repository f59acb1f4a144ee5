"""Throughput of the generic (n_fft != 512) forward / inverse path against its algorithmic bytes."""
import sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from ml_audio_inpainting_b200 import _cabi, spectral as sp
if "--lib" in sys.argv:          # A/B of a variant build (tools/build_variant.sh): experiment tooling only
    _cabi.LIB_PATH = Path(sys.argv[sys.argv.index("--lib") + 1]).resolve()
    print("library:", _cabi.LIB_PATH)

def timeit(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n

B, L = 512, 160000
x = (0.1 * torch.randn(B, L, device="cuda")).clamp_(-1, 1)
import os
print("AIP_POW2 =", os.environ.get("AIP_POW2", "1 (tiled radix-16 kernels)"))
for n_fft, hop, win in ((2048, 512, 2048), (1024, 256, 1024), (256, 64, 256), (512, 192, 384)):
    plan = sp.get_plan(n_fft, hop, win)
    T = plan.num_frames(L)
    F = n_fft // 2 + 1
    out = {"spec": torch.empty((B, F, T), dtype=torch.complex64, device="cuda")}
    t = timeit(lambda: sp.stft(x, plan, out=out))
    by = B * (4 * L + 8 * F * T)
    S = out["spec"]
    y = torch.empty((B, plan.istft_length(T)), device="cuda")
    ti = timeit(lambda: sp.istft(plan, spec=S, out=y))
    byi = B * (8 * F * T + 4 * y.shape[1])
    print(f"n_fft {n_fft} hop {hop}: stft(complex) {t:7.3f} ms {by / t / 1e6:7.1f} GB/s   istft {ti:7.3f} ms {byi / ti / 1e6:7.1f} GB/s", flush=True)
