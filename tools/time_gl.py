"""Griffin-Lim, 32 iterations x 1024 clips x 10 s (bench leg griffinlim32): fused phase update (INV_GL) against AIP_GL_UNFUSED=1."""
import os, sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from ml_audio_inpainting_b200 import spectral as sp
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
plan = sp.get_plan(512, 192, 384)
x = (0.1 * torch.randn(B, 160000, device="cuda")).clamp_(-1, 1)
mag = sp.stft(x, plan, mag_kind=sp.MAG_ABS, want_spec=False)["mag"]
del x
gen = torch.Generator(device="cuda").manual_seed(99)
ang = torch.polar(torch.ones_like(mag), 6.2831853 * torch.rand(mag.shape, device="cuda", generator=gen))
res = {}
for name, env, tma in (("fused-tma", None, None), ("fused", None, "0"), ("unfused", "1", "0"), ("fused-tma", None, None)):
    with sp.experiment_env(AIP_GL_UNFUSED=env, AIP_INV_TMA=tma):
        y = sp.griffinlim(plan, mag, n_iter=32, init_angles=ang)
        torch.cuda.synchronize()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3): y = sp.griffinlim(plan, mag, n_iter=32, init_angles=ang)
        e1.record(); torch.cuda.synchronize()
    t = e0.elapsed_time(e1) / 3
    res[name] = y
    print(f"{name:10s} {t:8.2f} ms  {B * 10.0 / t * 1e3 / 1e3:8.1f} k audio-s/s", flush=True)
d = (res["fused"] - res["unfused"]).abs().max().item() / res["unfused"].abs().max().item()
print(f"fused vs unfused after 32 iterations: relative max-abs difference {d:.3e}")
print("fused-tma bit-identical to fused:", bool(torch.equal(res["fused-tma"], res["fused"])))
