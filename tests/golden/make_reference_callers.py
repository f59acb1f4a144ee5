"""Regenerates tests/golden/reference_callers.npz by RUNNING THE REFERENCE'S OWN CALLER SOURCE (build container only).

What runs, unmodified, from /root/reference (tests/refshim.py supplies stand-ins for the three absent third-party packages
only -- librosa := oracle/librosa_port.py, soundfile := the package's FLAC codec, matplotlib := stubs):

  utils.py                                   every function the callers below reach
  models/CNNBLSTM/dataset.py:24-121          LibriSpeechDataset.__init__ / __getitem__   (root = /root/reference, split = test_samples)
  models/GAN/dataset.py:12-166               SpeechInpaintingDataset.__getitem__
  models/model_eval.py:48-194                inpaint() with a randomly initialised StackedBLSTMCNN (cnn_blstm.yaml) and PConvUNet
                                             (GAN/config.yaml) -- the checkpoints are missing blobs -- whose raw outputs are
                                             captured with forward hooks and stored, so that the B200 path can be fed the
                                             very same "model output"
  pre_process_dataset.py:19-43               the bulk loop body: add_random_gap(path, 0.1) + save_audio (through utils)

Everything is seeded (np.random.seed / torch.manual_seed).  Large arrays are stored as exact integer facts (gap samples,
frame ranges, PCM) plus a seeded sample of 8192 values per array and float64 sums; the tests compare the drop-in front-ends
/ back-end (GPU tier, through the C ABI) and the oracle's caller restatement (CPU tier) with them.
"""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
from tests import refshim                                            # noqa: E402

OUT = Path(__file__).resolve().parent
N_SAMPLE = 8192
CNN_FILES, CNN_GAPS, CNN_SEED = 3, 4, 1234
GAN_FILES, GAN_SEED = 3, 4321
EVAL_CNN_CLIPS, EVAL_GAN_CLIPS = 3, 1
PRE_FILES, PRE_SEED = 3, 99


def sample_idx(shape, tag):
    """The positions (flat indices) at which a big array is stored."""
    n = int(np.prod(shape))
    return np.random.default_rng(sum(map(ord, tag))).choice(n, size=min(N_SAMPLE, n), replace=False)


def pack(out, key, arr):
    """Store a seeded sample + the float64 sum of |arr| (arrays are too big to commit whole)."""
    a = np.asarray(arr)
    idx = sample_idx(a.shape, key)
    out[key + "/shape"] = np.array(a.shape, np.int64)
    out[key + "/idx"] = idx.astype(np.int64)
    out[key + "/val"] = a.reshape(-1)[idx]
    out[key + "/abssum"] = np.array(np.abs(a).astype(np.float64).sum())


def mask_range(mask2d, inside):
    """The frame range [f0, f1) a dense [F, T] frame mask encodes (asserting that it is one frame range)."""
    cols = np.flatnonzero(np.all(mask2d == inside, axis=0))
    other = 1 - inside
    if len(cols) == 0:
        assert np.all(mask2d == other)
        return 0, 0
    f0, f1 = int(cols[0]), int(cols[-1]) + 1
    assert np.array_equal(cols, np.arange(f0, f1)) and np.all(np.delete(mask2d, cols, axis=1) == other)
    return f0, f1


def generate(utils_impl: str = "reference"):
    """``utils_impl='dropin'`` runs the SAME reference caller source on the B200 drop-in utils (needs a GPU and the reference
    checkout: tools/run_reference_callers_on_dropin.py)."""
    import torch
    import yaml
    from ml_audio_inpainting_b200 import audio_io

    out = {}
    with refshim.reference_modules(utils_impl) as ref:
        REF = refshim.REFERENCE
        names = sorted(p.stem for p in (REF / "test_samples").glob("*.flac"))
        out["names"] = np.array(names)

        # ---- LibriSpeechDataset.__getitem__ (models/CNNBLSTM/dataset.py:74-121) ----
        cfg = yaml.safe_load((REF / "models" / "CNNBLSTM" / "cnn_blstm.yaml").read_text())
        cfg["data"].update(root_path=str(REF), test_path="test_samples", n_files=CNN_FILES, gaps_per_audio=CNN_GAPS)
        cfg_path = ref.tmp / "cnn_blstm_fixture.yaml"
        cfg_path.write_text(yaml.safe_dump(cfg))
        ds = ref.cnnblstm_dataset.LibriSpeechDataset(str(cfg_path), "test")
        # dataset.py:62-71 keeps the first n_files in os.walk order, THEN sorts: which clips those are depends on the file system
        out["cnn/files"] = np.array([Path(p).stem for p in ds.file_paths])
        assert len(ds) == CNN_FILES and set(out["cnn/files"]) <= set(names)
        np.random.seed(CNN_SEED)
        for i in range(len(ds)):
            gaps_, ints, masks, targets = ds[i]
            assert gaps_.dtype == torch.float32 and targets.dtype == torch.complex64 and ints.dtype == torch.float32
            pack(out, f"cnn/{i}/spectrogram_gaps", gaps_.numpy())
            pack(out, f"cnn/{i}/spectrogram_target_phases", targets.numpy())
            out[f"cnn/{i}/gap_ints"] = ints.numpy()
            out[f"cnn/{i}/gap_frames"] = np.array([mask_range(m, 1) for m in masks.numpy()], np.int64)
        out["cnn/rng_after"] = np.array(np.random.randint(0, 1 << 30))          # the stream position after the items

        # ---- SpeechInpaintingDataset.__getitem__ (models/GAN/dataset.py:63-166) ----
        gcfg = yaml.safe_load((REF / "models" / "GAN" / "config.yaml").read_text())
        gcfg["data"].update(root_path=str(REF), test_path="test_samples")
        gds = ref.gan_dataset.SpeechInpaintingDataset(gcfg, "test")
        assert [p.stem for p in gds.file_paths] == names              # GAN/dataset.py:53-56: rglob + sort
        np.random.seed(GAN_SEED)
        for i in range(GAN_FILES):
            item = gds[i]
            for k in ("original_magnitude", "impaired_magnitude", "original_phase"):
                assert item[k].dtype == torch.float32 and tuple(item[k].shape) == (1, 257, 626)
                pack(out, f"gan/{i}/{k}", item[k][0].numpy())
            out[f"gan/{i}/gap_frames"] = np.array(mask_range(item["mask"][0].numpy(), 0), np.int64)
        out["gan/rng_after"] = np.array(np.random.randint(0, 1 << 30))

        # ---- model_eval.inpaint (models/model_eval.py:48-194), CNN-BLSTM branch ----
        me = ref.model_eval
        torch.manual_seed(0)
        model = me.StackedBLSTMCNN(str(REF / "models" / "CNNBLSTM" / "cnn_blstm.yaml")).eval()
        grabbed = {}
        model.register_forward_hook(lambda mod, args, res: grabbed.__setitem__("out", res.detach().clone()))
        for i in range(EVAL_CNN_CLIPS):
            dst = ref.tmp / f"{names[i]}_cnnlstm.flac"
            me.inpaint(model, str(REF / "models" / "CNNBLSTM" / "cnn_blstm.yaml"), str(REF / "test_samples" / (names[i] + ".flac")),
                       str(dst), torch.device("cpu"))
            pcm, info = audio_io.decode_flac(dst.read_bytes(), verify_md5=True)
            raw = grabbed["out"][0].numpy()                                       # [257, 417] log10-domain model output
            out[f"eval_cnn/{i}/model_out_gap"] = raw[:, 166:173].astype(np.float32)
            out[f"eval_cnn/{i}/pcm"] = np.asarray(pcm).reshape(-1).astype(np.int16)

        # ---- model_eval.inpaint, GAN branch ----
        ecfg = me.load_config(str(REF / "models" / "GAN" / "config.yaml"))["model"]["generator"]
        torch.manual_seed(1)
        gen = me.PConvUNet(input_channels=ecfg["input_channels"], mask_channels=ecfg["mask_channels"],
                           output_channels=ecfg["output_channels"]).eval()
        gen.register_forward_hook(lambda mod, args, res: grabbed.__setitem__("gan", res.detach().clone()))
        for i in range(EVAL_GAN_CLIPS):
            dst = ref.tmp / f"{names[i]}_gan.flac"
            me.inpaint(gen, str(REF / "models" / "GAN" / "config.yaml"), str(REF / "test_samples" / (names[i] + ".flac")),
                       str(dst), torch.device("cpu"))
            pcm, info = audio_io.decode_flac(dst.read_bytes(), verify_md5=True)
            out[f"eval_gan/{i}/generator_out"] = grabbed["gan"][0, 0].numpy().astype(np.float32)     # [257, 626]
            out[f"eval_gan/{i}/pcm"] = np.asarray(pcm).reshape(-1).astype(np.int16)

        # ---- the bulk loop body (pre_process_dataset.py:36-41): add_random_gap(path, 0.1) -> save_audio ----
        np.random.seed(PRE_SEED)
        for i in range(PRE_FILES):
            src = REF / "test_samples" / (names[i] + ".flac")
            audio_new, gap_int = ref.utils.add_random_gap(src, 0.1)
            assert audio_new.dtype == np.float64 and audio_new.shape == (80000,)     # the float64 quirk of utils.py:180-183
            dst = ref.tmp / f"{names[i]}_pre.flac"
            ref.utils.save_audio(audio_new, dst)
            pcm, info = audio_io.decode_flac(dst.read_bytes(), verify_md5=True)
            out[f"pre/{i}/gap_int_s"] = np.array(gap_int, np.float64)
            out[f"pre/{i}/pcm"] = np.asarray(pcm).reshape(-1).astype(np.int16)
    return out


def main():
    out = generate()
    np.savez_compressed(OUT / "reference_callers.npz", **out)
    print({k: (v.shape, v.dtype) for k, v in out.items() if not k.endswith(("/idx", "/shape"))})


if __name__ == "__main__":
    main()
