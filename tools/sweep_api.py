"""Throughput sweep over the option combinations of the batched API (1024 x 5 s clips, n_fft 512 / hop 192 / win 384 and the GAN
geometry): looks for performance cliffs -- a combination that costs far more than the outputs it writes."""
import itertools
import sys
from pathlib import Path
import numpy as np, torch
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from ml_audio_inpainting_b200 import spectral as sp

def timeit(fn, n=5):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n

B, L = 1024, 80000
x = (0.1 * torch.randn(B, L, device="cuda")).clamp_(-1, 1)
gaps = np.stack([np.full(B, 30000), np.full(B, 33200)], 1)
frm = np.stack([np.full(B, 100), np.full(B, 120)], 1)
for hop, win in ((192, 384), (128, 512)):
    plan = sp.get_plan(512, hop, win)
    T = plan.num_frames(L)
    print(f"== forward, hop {hop} win {win}, T {T}")
    for mk, spec, phase, mask, zero in itertools.product((sp.MAG_NONE, sp.MAG_ABS, sp.MAG_LOG10_EPS, sp.MAG_LOG1P_POW, sp.MAG_POW),
                                                        (False, True), (False, True), (False, True), (False, True)):
        if mk == sp.MAG_NONE and not spec:
            continue
        kw = dict(mag_kind=mk, want_spec=spec, want_phase=phase, want_mask=mask, gap_samples=gaps)
        if mk == sp.MAG_POW: kw["power"] = 2.0
        if mask: kw["mask_frames"] = frm
        if zero: kw["zero_frames"] = frm
        t = timeit(lambda: sp.stft(x, plan, **kw))
        by = B * (4 * L + 257 * T * (4 * (mk != 0) + 8 * spec + 4 * phase + 4 * mask))
        flag = "   <-- " if by / t / 1e6 < 1500 else ""
        print(f"  mag {mk} spec {int(spec)} phase {int(phase)} mask {int(mask)} zero {int(zero)}: {t:7.3f} ms {by / t / 1e6:7.0f} GB/s{flag}")
    S = sp.stft(x, plan)["spec"]
    mag, ph = S.abs(), S.angle()
    print(f"== inverse, hop {hop} win {win}")
    for name, fn in (("complex", lambda: sp.istft(plan, spec=S)), ("complex + length", lambda: sp.istft(plan, spec=S, length=L)),
                     ("complex normalised", lambda: sp.istft(plan, spec=S, normalize=True)),
                     ("complex -> pcm16", lambda: sp.istft(plan, spec=S, normalize=True, pcm16=True)),
                     ("mag", lambda: sp.istft(plan, mag=mag)), ("mag + phase", lambda: sp.istft(plan, mag=mag, phase=ph)),
                     ("mag + phase, 10**", lambda: sp.istft(plan, mag=mag, phase=ph, mag_domain=sp.DOM_POW10)),
                     ("mag + phase, dB auto", lambda: sp.istft(plan, mag=mag, phase=ph, db_auto=True)),
                     ("mag + phase, expm1", lambda: sp.istft(plan, mag=mag, phase=ph, mag_domain=sp.DOM_EXPM1)),
                     ("hand-off CNN", lambda: sp.istft_blend(plan, mag, mag, mag, ph)),
                     ("hand-off GAN normalised", lambda: sp.istft_blend(plan, mag, mag, mag, ph, mag_domain=sp.DOM_LINEAR, mask_keeps_input=True, normalize=True))):
        t = timeit(fn)
        print(f"  {name:28s} {t:7.3f} ms")
