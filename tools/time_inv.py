"""iSTFT of B clips (bench leg istft, configs[2]): direct-load kernel against the TMA-staged one (AIP_INV_TMA), with and without
the L2 row requests of stage A (AIP_INV_L2_PREFETCH).   python tools/time_inv.py [B] [L]"""
import sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from ml_audio_inpainting_b200 import spectral as sp
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
L = int(sys.argv[2]) if len(sys.argv) > 2 else 160000
for hop, win in ((192, 384), (128, 512)):
    plan = sp.get_plan(512, hop, win)
    x = (0.1 * torch.randn(B, L, device="cuda")).clamp_(-1, 1)
    r = sp.stft(x, plan, mag_kind=sp.MAG_ABS, want_spec=True, want_phase=True)
    S, mag, ph = r["spec"], r["mag"], r["phase"]
    del x, r
    out = torch.empty((B, plan.istft_length(S.shape[2])), device="cuda")
    res = {}
    cases = [("direct", dict(AIP_INV_TMA="0", AIP_INV_L2_PREFETCH="0")), ("direct+l2", dict(AIP_INV_TMA="0")),
             ("default", dict()), ("direct", dict(AIP_INV_TMA="0", AIP_INV_L2_PREFETCH="0")), ("default", dict())]
    for name, env in cases:
        for kind in ("spec", "mag+phase"):
            kw = dict(spec=S) if kind == "spec" else dict(mag=mag, phase=ph)
            with sp.experiment_env(**env):
                for _ in range(5): sp.istft(plan, out=out, **kw)
                torch.cuda.synchronize()
                e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(20): sp.istft(plan, out=out, **kw)
                e1.record(); torch.cuda.synchronize()
                t = e0.elapsed_time(e1) / 20
            res[(name, kind)] = out.clone()
            gb = B * (S.shape[1] * S.shape[2] * 8 + out.shape[1] * 4) / 1e9
            print(f"hop {hop} T {S.shape[2]} {kind:9s} {name:10s} {t:7.4f} ms  {gb / t * 1e3:7.1f} GB/s", flush=True)
    print("bit-identical:", all(bool(torch.equal(res[("direct", k)], res[(n, k)])) for n in ("direct+l2", "default") for k in ("spec", "mag+phase")))
    del S, mag, ph, out, res
    torch.cuda.empty_cache()
