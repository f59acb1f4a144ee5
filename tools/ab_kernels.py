"""A/B timing of kernel variants: each argument is a libaip_b200 build (tools/build_variant.sh); the same
headline launches are timed with CUDA events for each.  Experiment tooling, not a product path.

    python tools/ab_kernels.py build/ab/a.so build/ab/b.so ...
"""
import ctypes as C
import json
import sys
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from ml_audio_inpainting_b200 import _cabi  # noqa: E402  (signatures only)

SR, N_FFT = 16000, 512
NO_PRUNE = "--no-prune" in sys.argv
CLIP_L = next((int(a.split("=")[1]) for a in sys.argv if a.startswith("--L=")), 160000)   # samples per clip


def load(path):
    lib = C.CDLL(str(path))
    for name, (res, args) in _cabi.SIGNATURES.items():
        if not hasattr(lib, name):
            continue
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    return lib


def hann_padded(win):
    n = np.arange(win)
    w = 0.5 - 0.5 * np.cos(2 * np.pi * n / win)
    out = np.zeros(N_FFT, np.float32)
    lo = (N_FFT - win) // 2
    out[lo:lo + win] = w
    return out


def time_it(fn, reps=20, warm=5):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(reps + 1)]
    ev[0].record()
    for i in range(reps):
        fn()
        ev[i + 1].record()
    torch.cuda.synchronize()
    t = [ev[i].elapsed_time(ev[i + 1]) for i in range(reps)]
    return round(float(np.mean(t)), 4), round(float(np.min(t)), 4)


def main():
    dev = torch.device("cuda:0")
    st = torch.cuda.current_stream().cuda_stream
    g = torch.Generator(device=dev).manual_seed(1234)
    res = {}
    for win, hop, B, tag in ((384, 192, 4096, "p1"), (512, 128, 1024, "p2")):
        L = CLIP_L
        wave = (0.1 * torch.randn(B, L, device=dev, generator=g)).clamp_(-1, 1)
        window = torch.from_numpy(hann_padded(win)).to(dev)
        T = 1 + L // hop
        rs = np.random.RandomState(0)
        gl = int(0.2 * SR)
        s0 = rs.randint(0, L - gl, size=B)
        gaps = torch.from_numpy(np.stack([s0, s0 + gl], 1).astype(np.int32)).to(dev)
        mag = torch.empty(B, 257, T, device=dev)
        Bi = min(B, 1024)
        spec = torch.empty(Bi, 257, T, 2, device=dev)
        out_len = hop * (T - 1)
        wav_out = torch.empty(Bi, out_len, device=dev)
        inv_wss = torch.empty(out_len, device=dev)
        for path in [a for a in sys.argv[1:] if not a.startswith("--")]:
            lib = load(path)
            # win_length in the descriptor enables the zero-tap pruning of newer builds (older builds ignore the field)
            desc = _cabi.StftDesc(N_FFT, hop, 1, 0 if NO_PRUNE else win, window.data_ptr())
            d = C.byref(desc)

            def fwd():
                rc = lib.aip_stft_fwd_f32(d, wave.data_ptr(), B, L, L, gaps.data_ptr(), None, None, 0, 2, 1e-9, 1.0, T,
                                          None, mag.data_ptr(), None, None, st)
                assert rc == 0, rc

            def fwd_spec():
                rc = lib.aip_stft_fwd_f32(d, wave.data_ptr(), Bi, L, L, gaps.data_ptr(), None, None, 0, 0, 0.0, 1.0, T,
                                          spec.data_ptr(), None, None, None, st)
                assert rc == 0, rc

            rc = lib.aip_inv_window_sumsquare_f32(d, T, 0, inv_wss.data_ptr(), out_len, st)
            assert rc == 0, rc

            def inv():
                rc = lib.aip_istft_f32(d, spec.data_ptr(), None, None, 0, None, Bi, T, 0, inv_wss.data_ptr(),
                                       wav_out.data_ptr(), out_len, None, 0, st)
                assert rc == 0, rc

            name = Path(path).stem
            r = res.setdefault(name, {})
            r[f"fwd_log10_{tag}"] = time_it(fwd)
            r[f"fwd_spec_{tag}"] = time_it(fwd_spec)
            r[f"inv_spec_{tag}"] = time_it(inv)
            # checksums so that variants can be compared for gross errors
            r[f"chk_{tag}"] = [round(float(mag[:64].double().mean()), 6), round(float(wav_out[:64].double().abs().mean()), 8)]
        del wave, mag, spec, wav_out
        torch.cuda.empty_cache()
    for name, r in res.items():
        print(json.dumps({"variant": name, **r}))


if __name__ == "__main__":
    main()
