"""Manual check (needs a B200 AND a checkout of the reference): run the reference's UNMODIFIED caller source --
models/CNNBLSTM/dataset.py, models/GAN/dataset.py, models/model_eval.py (both branches), the pre_process_dataset.py loop
body -- on the B200 drop-in ``utils`` and compare with what the same source produces on the reference's own utils.py
(tests/golden/reference_callers.npz).

    AIP_REFERENCE_ROOT=/path/to/ml-audio-inpainting python tools/run_reference_callers_on_dropin.py

The driver's GPU boxes have no reference checkout, so this is not part of the test tiers; the log of a run is kept under
profiles/.  Nothing is copied into this repository.
"""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from tests.golden import make_reference_callers as mk       # noqa: E402


def main():
    fx = dict(np.load(ROOT / "tests" / "golden" / "reference_callers.npz"))
    got = mk.generate("dropin")
    assert sorted(got) == sorted(fx), "different set of outputs"
    worst = {}
    for k in sorted(fx):
        a, b = got[k], fx[k]
        if k.endswith(("/idx", "/shape")) or k in ("names", "cnn/files") or k.endswith(("gap_frames", "gap_ints", "gap_int_s", "rng_after")):
            assert np.array_equal(a, b), k                                   # integer facts, draws, frame ranges: exact
        elif k.endswith("/pcm"):
            d = np.abs(a.astype(np.int64) - b.astype(np.int64))
            assert d.max() <= 1 and (d > 0).mean() < 2.5e-2, (k, int(d.max()), float((d > 0).mean()))
            worst[k] = f"max {int(d.max())} LSB on {100 * (d > 0).mean():.2f} % of the samples"
        elif k.endswith("/val"):
            if "spectrogram_gaps" in k:                                      # log10 domain: compare linearly away from the floor
                e = np.abs(10.0 ** a.astype(np.float64) - 10.0 ** b.astype(np.float64)).max() / (10.0 ** b.astype(np.float64)).max()
            elif "magnitude" in k:                                           # log1p domain
                e = np.abs(np.expm1(a.astype(np.float64)) - np.expm1(b.astype(np.float64))).max() / np.expm1(b.astype(np.float64)).max()
            elif "phase" in k and "target" not in k:                         # angles: compare as phasors, weighted by the magnitude sample
                w = np.expm1(fx[k.replace("original_phase", "original_magnitude")].astype(np.float64))
                e = np.abs(w * (np.exp(1j * a) - np.exp(1j * b))).max() / w.max()
            else:
                e = np.abs(a - b).max() / np.abs(b).max()
            assert e < 1e-4, (k, e)
            worst[k] = f"rel. max-abs {e:.2e}"
        elif k.endswith("/abssum"):
            pass
        elif "model_out_gap" in k or "generator_out" in k:
            e = np.abs(a - b).max() / np.abs(b).max()                        # the random-init models see fp32 spectra: small drift
            worst[k] = f"model output rel. max-abs {e:.2e} (not asserted: depends on the torch model, not on the path)"
        else:
            raise AssertionError(f"unclassified key {k}")
    for k, v in worst.items():
        print(f"{k:48s} {v}")
    print("OK: the reference's caller source on the B200 drop-in matches its output on the reference's own utils.py")


if __name__ == "__main__":
    main()
