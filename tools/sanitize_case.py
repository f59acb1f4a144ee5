"""Small end-to-end case for compute-sanitizer (memcheck / racecheck): forward (all emitters), inverse (both
loaders, edge + interior tiles), Griffin-Lim, generic path, small kernels.  Sizes kept tiny."""
import sys
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from ml_audio_inpainting_b200 import frontend, preprocess, spectral as sp   # noqa: E402

torch.manual_seed(0)
x = (0.1 * torch.randn(5, 16000, device="cuda")).clamp_(-1, 1)
gaps = np.array([[100, 1700], [0, 320], [15000, 16000], [4000, 4000], [7000, 9000]])
for hop, win in ((192, 384), (128, 512)):
    plan = sp.get_plan(512, hop, win, "hann", True, "cuda:0")
    a = sp.stft(x, plan, gap_samples=gaps, mag_kind=sp.MAG_LOG10_EPS, want_spec=False)
    b = sp.stft(x, plan, mag_kind=sp.MAG_ABS, want_spec=True, want_phase=True, want_mask=True,
                mask_frames=np.array([[3, 9]] * 5), zero_frames=np.array([[3, 9]] * 5))
    c = sp.stft(x, plan)
    y = sp.istft(plan, spec=c["spec"])
    y2 = sp.istft(plan, mag=b["mag"], phase=b["phase"], db_auto=True)
    y3 = sp.istft(plan, spec=c["spec"], length=12345)
    g = sp.griffinlim(plan, b["mag"][:2], n_iter=2)
frontend.cnnblstm_batch(x[:, :8000].repeat(1, 10))
frontend.gan_batch(x[:, :8000].repeat(1, 10))
preprocess.preprocess_batch(x, 0.1, want_logmag=True)
plan = sp.get_plan(1024, 256, 1024, "hann", True, "cuda:0")
s = sp.stft(x, plan)["spec"]
sp.istft(plan, spec=s)
torch.cuda.synchronize()
print("sanitize case done", float(y.abs().max()), float(g.abs().max()))
