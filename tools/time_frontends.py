"""Times the batched caller front-ends / back-end on one GPU."""
import sys, time
from pathlib import Path
import numpy as np, torch
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from ml_audio_inpainting_b200 import frontend, spectral as sp

def timeit(fn, n=10, w=3):
    for _ in range(w): fn()
    torch.cuda.synchronize(); e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n

B, L = 1024, 80000
x = (0.1 * torch.randn(B, L, device="cuda")).clamp_(-1, 1)
np.random.seed(0)
secs = B * 5.0
t = timeit(lambda: frontend.cnnblstm_batch(x)); print(f"cnnblstm_batch (logmag+mask+target) {t:8.3f} ms  {secs / t * 1e3 / 1e6:8.2f} M audio-s/s")
t = timeit(lambda: frontend.cnnblstm_batch(x, want_target=False)); print(f"cnnblstm_batch (logmag+mask)        {t:8.3f} ms  {secs / t * 1e3 / 1e6:8.2f} M audio-s/s")
t = timeit(lambda: frontend.gan_batch(x)); print(f"gan_batch (4 outputs, P2)            {t:8.3f} ms  {secs / t * 1e3 / 1e6:8.2f} M audio-s/s")
t = timeit(lambda: frontend.eval_cnnlstm_batch(x)); print(f"eval_cnnlstm_batch                   {t:8.3f} ms  {secs / t * 1e3 / 1e6:8.2f} M audio-s/s")
ev = frontend.eval_cnnlstm_batch(x)
t = timeit(lambda: frontend.backend_batch(ev["log_impaired_magnitude"], ev["original_phase"], mag_domain=sp.DOM_POW10)); print(f"backend_batch (10**, mag+phase)       {t:8.3f} ms  {secs / t * 1e3 / 1e6:8.2f} M audio-s/s")
plan = sp.get_plan(512, 192, 384)
t = timeit(lambda: sp.istft(plan, spec=ev["original_spectrogram"])); print(f"istft complex 5 s clips              {t:8.3f} ms  {secs / t * 1e3 / 1e6:8.2f} M audio-s/s")
