"""A few warm launches of the tiled radix-16 kernels at the reference's default n_fft 2048 / hop 512 (for ncu)."""
import sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from ml_audio_inpainting_b200 import spectral as sp
n_fft, hop = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (2048, 512)
B, L = 512, 160000
x = (0.1 * torch.randn(B, L, device="cuda")).clamp_(-1, 1)
plan = sp.get_plan(n_fft, hop, n_fft)
for _ in range(3):
    S = sp.stft(x, plan)["spec"]
    y = sp.istft(plan, spec=S)
torch.cuda.synchronize()
print(S.shape, y.shape)
