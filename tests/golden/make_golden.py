"""Regenerates tests/golden/*.npz from the reference checkout (run in the build container only).

  clips_int16.npz   the 9 /root/reference/test_samples/*.flac clips after utils.load_audio's truncation
                    (first 80 000 samples = 5 s at 16 kHz; utils.py:38-41), kept as int16 PCM; decoded with
                    the package's own FLAC reader, each file verified against its STREAMINFO MD5.
  anchors.json      per clip, oracle known-answer scalars at P1 (n_fft 512 / win 384 / hop 192):
                    shape, max|S|, max log10(|S_gap|+1e-9), gap frames, round-trip SNR, iSTFT length,
                    and a float64 checksum of |S|; cross-checked against torch.stft on generation.

The reference holds no golden vectors of its own for this path and librosa cannot
be imported here, so these anchors pin the ORACLE (regression) and the decode, not librosa itself.
"""
import json
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
from ml_audio_inpainting_b200 import audio_io                      # noqa: E402
from oracle import callers_port as cp, librosa_port as lr, utils_port as up   # noqa: E402

REF = Path("/root/reference/test_samples")
OUT = Path(__file__).resolve().parent


def main():
    clips, anchors = {}, {}
    for f in sorted(REF.glob("*.flac")):
        data = f.read_bytes()
        pcm, info = audio_io.decode_flac(data, verify_md5=True)
        assert info.sample_rate == 16000 and info.channels == 1 and info.bits_per_sample == 16
        pcm = np.asarray(pcm).reshape(-1)
        clip = pcm[:80000].astype(np.int16)
        clips[f.stem] = clip
        x = clip.astype(np.float32) / np.float32(32768.0)
        S = lr.stft(x, n_fft=512, hop_length=192, win_length=384)
        w = torch.from_numpy(lr.fft_window("hann", 384, 512).astype(np.float32))
        St = torch.stft(torch.from_numpy(x), 512, 192, 512, window=w, center=True, pad_mode="constant",
                        return_complex=True).numpy()
        assert np.abs(S - St).max() / np.abs(S).max() < 1e-5
        mask, (s0, s1) = up.create_gap_mask(len(x), 0.08, 16000, gap_start_s=2.0)
        Sg = lr.stft(x * mask, n_fft=512, hop_length=192, win_length=384)
        ev = cp.eval_frontend_cnnlstm(x)
        y = lr.istft(S, hop_length=192, win_length=384, n_fft=512)
        n = len(y)
        snr = 10 * np.log10((x[:n].astype(np.float64) ** 2)[512:n - 512].sum()
                            / ((x[:n] - y).astype(np.float64) ** 2)[512:n - 512].sum())
        anchors[f.stem] = {
            "total_samples": int(info.total_samples), "shape": list(S.shape), "max_abs_S": float(np.abs(S).max()),
            "max_log10_gap": float(np.log10(np.abs(Sg) + 1e-9).max()), "min_log10_specgap": float(ev["log_impaired_magnitude"].min()),
            "gap_samples": [int(s0), int(s1)], "cnnlstm_gap_frames": [int(v) for v in ev["gap_frames"]],
            "gan_gap_frames": [int(v) for v in cp.gan_frame_mask_range(s0, s1, 128, 626)],
            "istft_len": int(n), "roundtrip_snr_db": float(snr), "sum_abs_S": float(np.abs(S).astype(np.float64).sum()),
        }
        print(f.stem, anchors[f.stem])
    np.savez_compressed(OUT / "clips_int16.npz", **clips)
    (OUT / "anchors.json").write_text(json.dumps(anchors, indent=1))


if __name__ == "__main__":
    main()
