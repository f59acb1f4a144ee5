"""Test support: run the REFERENCE'S OWN Python source in this container (tests and fixture generation only).

The reference (``/root/reference``, read-only, absent on the GPU box) cannot be imported as it is: ``utils.py:2-6``
imports librosa, soundfile and matplotlib, none of which is installed or installable here.  ``reference_modules()``
puts stand-ins for exactly those three THIRD-PARTY packages into ``sys.modules``

    librosa     -> oracle/librosa_port.py   (the restatement of librosa >= 0.10 that the tests already pin)
    soundfile   -> ml_audio_inpainting_b200/audio_io.py (FLAC / WAV codec)
    matplotlib  -> inert stubs (plotting is out of scope)

and then loads the reference's files unmodified, by path: ``utils.py``, ``add_gaps.py``, ``models/CNNBLSTM/dataset.py``,
``models/CNNBLSTM/model.py``, ``models/GAN/dataset.py``, ``models/model_eval.py``.  ``config`` resolves to the drop-in ``config.py`` (same constants)
because the reference's creates ``<its own directory>/output`` on import and /root/reference must not be written to.

``utils_impl='reference'`` runs the reference's callers on the reference's utils (fixture generation, oracle checks);
``utils_impl='dropin'`` runs the very same caller source on the B200 drop-in ``utils`` (needs a GPU *and* the reference
checkout, which never coincide on the driver's boxes: that combination is a manual check).
Nothing under the product package imports this module.
"""
from __future__ import annotations

import contextlib
import importlib.util
import os
import sys
import tempfile
import types
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
REFERENCE = Path(os.environ.get("AIP_REFERENCE_ROOT", "/root/reference"))


def reference_available() -> bool:
    return (REFERENCE / "utils.py").exists()


# --------------------------------------------------------------------------------------------- third-party stand-ins
def _librosa_module() -> types.ModuleType:
    from oracle import librosa_port as lr
    from ml_audio_inpainting_b200 import audio_io

    m = types.ModuleType("librosa")
    m.__version__ = "0.10-oracle-port"

    def load(path, sr=22050, mono=True, **_):
        """librosa.load for files already at the target rate: libsndfile -> float32, mean over channels."""
        pcm, file_sr = audio_io.read_audio(path)
        y = pcm.astype(np.float32)
        y = y.mean(axis=1, dtype=np.float32) if (mono and y.ndim == 2) else y.T.squeeze()
        if sr is not None and file_sr != sr:
            import math
            import scipy.signal
            g = math.gcd(int(file_sr), int(sr))
            y = scipy.signal.resample_poly(y, sr // g, file_sr // g).astype(np.float32)
            file_sr = sr
        return y, file_sr

    def chirp(*, fmin, fmax, sr=22050, length=None, duration=None, linear=False, phi=None):
        import scipy.signal
        period = 1.0 / sr
        duration = duration if duration is not None else period * length
        phi = -np.pi * 0.5 if phi is None else phi
        t = np.arange(0, duration, period)
        return scipy.signal.chirp(t, fmin, duration, fmax, method="linear" if linear else "logarithmic",
                                  phi=phi / np.pi * 180)

    def amplitude_to_db(S, ref=1.0, amin=1e-5, top_db=80.0):
        mag = np.abs(np.asarray(S))
        ref_v = ref(mag) if callable(ref) else np.abs(ref)
        return power_to_db(mag ** 2, ref=ref_v ** 2, amin=amin ** 2, top_db=top_db)

    def power_to_db(S, ref=1.0, amin=1e-10, top_db=80.0):
        mag = np.abs(np.asarray(S))
        ref_v = ref(mag) if callable(ref) else np.abs(ref)
        log_spec = 10.0 * np.log10(np.maximum(amin, mag)) - 10.0 * np.log10(np.maximum(amin, ref_v))
        return np.maximum(log_spec, log_spec.max() - top_db) if top_db is not None else log_spec

    for name in ("stft", "istft", "griffinlim", "time_to_frames", "time_to_samples", "samples_to_frames",
                 "db_to_amplitude", "db_to_power", "hz_to_mel", "mel_to_hz", "mel_frequencies"):
        setattr(m, name, getattr(lr, name))
    m.load, m.chirp, m.amplitude_to_db, m.power_to_db = load, chirp, amplitude_to_db, power_to_db
    util = types.ModuleType("librosa.util")
    util.normalize, util.pad_center, util.tiny, util.phasor = lr.normalize, lr.pad_center, lr.tiny, lr.phasor
    feature = types.ModuleType("librosa.feature")
    feature.melspectrogram = lr.melspectrogram
    filters = types.ModuleType("librosa.filters")
    filters.mel, filters.get_window, filters.window_sumsquare = lr.mel, lr.get_window, lr.window_sumsquare
    display = types.ModuleType("librosa.display")
    display.specshow = lambda *a, **k: None
    m.util, m.feature, m.filters, m.display = util, feature, filters, display
    return m


def _soundfile_module() -> types.ModuleType:
    from ml_audio_inpainting_b200 import audio_io
    m = types.ModuleType("soundfile")

    def write(file, data, samplerate, subtype=None, endian=None, format=None, closefd=True):
        fmt = (format or Path(str(file)).suffix.lstrip(".") or "wav").lower()
        audio_io.write_audio(file, np.asarray(data), int(samplerate), fmt)

    def read(file, dtype="float64", **_):
        pcm, sr = audio_io.read_audio(file)
        pcm = pcm[:, 0] if (pcm.ndim == 2 and pcm.shape[1] == 1) else pcm
        return pcm.astype(dtype), sr

    m.write, m.read = write, read
    return m


def _matplotlib_modules() -> dict:
    class _Anything:
        def __getattr__(self, name):
            return _Anything()

        def __call__(self, *a, **k):
            return _Anything()

        def __iter__(self):
            return iter(())

    mpl = types.ModuleType("matplotlib")
    pyplot = types.ModuleType("matplotlib.pyplot")
    figure = types.ModuleType("matplotlib.figure")

    def _attr(name):
        if name.startswith("__"):
            raise AttributeError(name)
        return _Anything()

    pyplot.__getattr__ = _attr
    figure.Figure = _Anything
    mpl.pyplot, mpl.figure = pyplot, figure
    mpl.use = lambda *a, **k: None
    return {"matplotlib": mpl, "matplotlib.pyplot": pyplot, "matplotlib.figure": figure}


def _load(name: str, path: Path) -> types.ModuleType:
    spec = importlib.util.spec_from_file_location(name, str(path))
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


@contextlib.contextmanager
def reference_modules(utils_impl: str = "reference"):
    """Yields a namespace with the reference's modules loaded from /root/reference (unmodified source):
    ``utils``, ``add_gaps``, ``cnnblstm_dataset``, ``cnnblstm_model``, ``gan_dataset``, ``model_eval``; sys.modules / sys.path are restored."""
    if not reference_available():
        raise RuntimeError(f"{REFERENCE} is not present (the reference checkout only exists in the build container)")
    names = ["librosa", "librosa.util", "librosa.feature", "librosa.filters", "librosa.display", "soundfile",
             "matplotlib", "matplotlib.pyplot", "matplotlib.figure", "utils", "config", "add_gaps",
             "ref_cnnblstm_dataset", "ref_cnnblstm_model", "ref_gan_dataset"]
    saved = {k: sys.modules.get(k) for k in names}
    before = set(sys.modules)
    saved_path = list(sys.path)
    saved_env = os.environ.get("AIP_OUTPUT_DIR")
    tmp = tempfile.mkdtemp(prefix="aip_refshim_")
    try:
        for k in names:
            sys.modules.pop(k, None)
        lib = _librosa_module()
        sys.modules.update({"librosa": lib, "librosa.util": lib.util, "librosa.feature": lib.feature,
                            "librosa.filters": lib.filters, "librosa.display": lib.display,
                            "soundfile": _soundfile_module(), **_matplotlib_modules()})
        os.environ["AIP_OUTPUT_DIR"] = tmp
        dropin = ROOT / "ml_audio_inpainting_b200" / "dropin"
        _load("config", dropin / "config.py")
        ns = types.SimpleNamespace()
        ns.utils = _load("utils", (REFERENCE if utils_impl == "reference" else dropin) / "utils.py")
        ns.add_gaps = _load("add_gaps", REFERENCE / "add_gaps.py")
        ns.cnnblstm_dataset = _load("ref_cnnblstm_dataset", REFERENCE / "models" / "CNNBLSTM" / "dataset.py")
        ns.cnnblstm_model = _load("ref_cnnblstm_model", REFERENCE / "models" / "CNNBLSTM" / "model.py")
        ns.gan_dataset = _load("ref_gan_dataset", REFERENCE / "models" / "GAN" / "dataset.py")
        # models/model_eval.py does `from GAN.train import load_config`, `from GAN.networks import ...`,
        # `from CNNBLSTM.model import ...` (model_eval.py:16-21): its own directory has to be importable
        sys.path.insert(0, str(REFERENCE / "models"))
        ns.model_eval = _load("ref_model_eval", REFERENCE / "models" / "model_eval.py")
        ns.tmp = Path(tmp)
        yield ns
    finally:
        import shutil
        shutil.rmtree(tmp, ignore_errors=True)
        for k in set(sys.modules) - before:             # GAN.*, CNNBLSTM.*, loss, dataset, networks, ref_*
            f = getattr(sys.modules[k], "__file__", None) or ""
            if k.startswith("ref_") or str(f).startswith(str(REFERENCE)):
                sys.modules.pop(k, None)
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
        sys.path[:] = saved_path
        if saved_env is None:
            os.environ.pop("AIP_OUTPUT_DIR", None)
        else:
            os.environ["AIP_OUTPUT_DIR"] = saved_env
