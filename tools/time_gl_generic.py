"""Griffin-Lim 32 iterations at the reference's default n_fft 2048 / hop 512 (the mel back-end's transform): phase update fused
into the tiled inverse's load against the separate update kernel."""
import sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from ml_audio_inpainting_b200 import spectral as sp
B, L = 256, 160000
plan = sp.get_plan(2048, 512, 2048)
x = (0.1 * torch.randn(B, L, device="cuda")).clamp_(-1, 1)
mag = sp.stft(x, plan, mag_kind=sp.MAG_ABS, want_spec=False)["mag"]
ang = torch.polar(torch.ones_like(mag), 6.2831853 * torch.rand_like(mag))
for name, env in (("fused", None), ("update kernel", "1"), ("fused", None), ("radix-2 kernels", "P")):
    kw = {"AIP_POW2": "0"} if env == "P" else {"AIP_GL_UNFUSED": env}
    with sp.experiment_env(**kw):
        for _ in range(2): sp.griffinlim(plan, mag, n_iter=32, init_angles=ang)
        torch.cuda.synchronize()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3): sp.griffinlim(plan, mag, n_iter=32, init_angles=ang)
        e1.record(); torch.cuda.synchronize()
    print(f"griffinlim 32 it, {B} x 10 s, n_fft 2048: {name:16s} {e0.elapsed_time(e1) / 3:8.2f} ms", flush=True)
