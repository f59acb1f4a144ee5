"""iSTFT of B x 10 s clips (bench leg istft, configs[2]): direct-load kernel against the TMA-staged one (AIP_INV_TMA=1)."""
import sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from ml_audio_inpainting_b200 import spectral as sp
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
for hop, win in ((192, 384), (128, 512)):
    plan = sp.get_plan(512, hop, win)
    x = (0.1 * torch.randn(B, 160000, device="cuda")).clamp_(-1, 1)
    S = sp.stft(x, plan)["spec"]
    if S.shape[2] % 2:
        S = S[:, :, :-1].contiguous()
    del x
    out = torch.empty((B, plan.istft_length(S.shape[2])), device="cuda")
    res = {}
    for name, env in (("direct", "0"), ("tma", None), ("direct", "0"), ("tma", None)):
        with sp.experiment_env(AIP_INV_TMA=env):
            for _ in range(5): sp.istft(plan, spec=S, out=out)
            torch.cuda.synchronize()
            e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(20): sp.istft(plan, spec=S, out=out)
            e1.record(); torch.cuda.synchronize()
            t = e0.elapsed_time(e1) / 20
        res[name] = out.clone()
        gb = B * (S.shape[1] * S.shape[2] * 8 + out.shape[1] * 4) / 1e9
        print(f"hop {hop} T {S.shape[2]} {name:7s} {t:7.4f} ms  {gb / t * 1e3:7.1f} GB/s", flush=True)
    print("bit-identical:", bool(torch.equal(res["direct"], res["tma"])))
    del S, out, res
    torch.cuda.empty_cache()
