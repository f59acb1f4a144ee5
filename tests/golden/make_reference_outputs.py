"""Regenerates tests/golden/reference_cnnlstm_inpainted_int16.npz from the reference checkout (build container only).

The reference ships the OUTPUT of its own evaluation script: /root/reference/test_samples_reconstructed/
<clip>_cnnlstm_inpainted.flac, written by models/model_eval.py:179-192 =
``utils.save_audio(utils.spectrogram_to_audio(inpainted, phase=angle(S_orig), n_fft=512, hop_length=192, win_length=384))``
with the real librosa / soundfile on the author's machine.  ``inpainted = 10 ** model.reconstruct_spectrogram(...)`` and
``reconstruct_spectrogram`` (models/CNNBLSTM/model.py:108) returns ``model_out * mask + log10(|S (1 - mask)| + 1e-9) *
(1 - mask)``: outside the gap frames [166, 173) the file therefore holds the reference's OWN
STFT -> log10 -> 10** -> phase reuse -> iSTFT -> peak-normalise -> PCM-16 round trip of the matching test_samples clip,
whatever the (missing) checkpoint produced inside the gap.  These nine files are the only values on this path that the
reference itself computed; they pin the oracle and the CUDA path to the reference (tests/test_reference_outputs.py).

Stored as the raw int16 PCM of each file (decoded with the package's FLAC reader, STREAMINFO MD5 verified).
"""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
from ml_audio_inpainting_b200 import audio_io                      # noqa: E402

REF = Path("/root/reference/test_samples_reconstructed")
OUT = Path(__file__).resolve().parent


def main():
    out = {}
    for f in sorted(REF.glob("*_cnnlstm_inpainted.flac")):
        pcm, info = audio_io.decode_flac(f.read_bytes(), verify_md5=True)
        assert info.sample_rate == 16000 and info.channels == 1 and info.bits_per_sample == 16
        pcm = np.asarray(pcm).reshape(-1)
        assert len(pcm) == 79872 == 192 * (417 - 1)                 # hop * (T - 1): librosa.istft's natural length
        out[f.stem.replace("_cnnlstm_inpainted", "")] = pcm.astype(np.int16)
        print(f.stem, len(pcm), int(np.abs(pcm.astype(np.int32)).max()))
    assert len(out) == 9
    np.savez_compressed(OUT / "reference_cnnlstm_inpainted_int16.npz", **out)


if __name__ == "__main__":
    main()
