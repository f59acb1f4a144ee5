"""One warm launch pair of aip_stft_gap_variants_f32 (fill + variant transform), for ncu captures.  usage: variant_once.py [N] [G] [reps]"""
import sys
from pathlib import Path
import numpy as np, torch
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from ml_audio_inpainting_b200 import spectral as sp
N = int(sys.argv[1]) if len(sys.argv) > 1 else 256
G = int(sys.argv[2]) if len(sys.argv) > 2 else 25
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
L, g, T = 80000, 3200, 417
x = (0.1 * torch.randn(N, L, device="cuda")).clamp_(-1, 1)
np.random.seed(0)
starts = np.random.randint(0, L - g, size=N * G)
gaps = torch.as_tensor(np.stack([starts, starts + g], 1).astype(np.int32), device="cuda")
plan = sp.get_plan(512, 192, 384)
clean = sp.stft(x, plan, mag_kind=sp.MAG_LOG10_EPS, t_out=T, want_spec=False)["mag"]
out = torch.empty((N * G, 257, T), device="cuda")
for _ in range(reps):
    sp.stft_gap_variants(x, plan, gaps, G, t_out=T, clean_mag=clean, out=out, gap_len_max=g)
torch.cuda.synchronize()
print("ok")
