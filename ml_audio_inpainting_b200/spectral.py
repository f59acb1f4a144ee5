"""Batched, device-resident STFT / iSTFT / Griffin-Lim on B200 (host side of the C ABI).

Tensors in, tensors out, all on one CUDA device; torch is used for memory and streams only --
every number is produced by the sm_100a kernels of ``csrc/aip_kernels.cu`` through
``include/aip_b200.h``.  Semantics follow librosa >= 0.10 as the reference calls it:

  stft        librosa.stft      via utils.extract_spectrogram   (reference utils.py:192-234)
  istft       librosa.istft     via utils.spectrogram_to_audio  (reference utils.py:316-327)
  griffinlim  librosa.griffinlim                                (reference utils.py:328-332)

Layout: waveforms ``[B, L]`` float32, spectrograms ``[B, F, T]`` (T contiguous, like librosa's
``[F, T]`` per clip), complex spectrograms as ``torch.complex64``.
"""
from __future__ import annotations

import contextlib
import ctypes as C
import os
from dataclasses import dataclass, field
from typing import Dict, Optional, Tuple

import numpy as np
import scipy.signal
import torch

from . import _cabi
from ._cabi import (DOM_DB, DOM_EXPM1, DOM_LINEAR, DOM_POW10, MAG_ABS, MAG_LOG10_EPS, MAG_LOG1P_POW,
                    MAG_NONE, MAG_POW, StftDesc, check)

__all__ = ["StftPlan", "get_plan", "stft", "stft_gap_variants", "istft", "istft_blend", "griffinlim", "db_heuristic", "fft_window",
           "wave_to_pcm16", "PCM_RAW", "PCM_NORMALIZE", "PCM_PEAKS_GIVEN", "mel_basis", "mel_project", "mel_inverse", "experiment_env",
           "MAG_NONE", "MAG_ABS", "MAG_LOG10_EPS", "MAG_LOG1P_POW", "MAG_POW",
           "DOM_LINEAR", "DOM_POW10", "DOM_DB", "DOM_EXPM1"]


def fft_window(window, win_length: int, n_fft: int) -> np.ndarray:
    """librosa.filters.get_window(window, win_length, fftbins=True) centre-padded to n_fft (float64)."""
    if callable(window):
        w = np.asarray(window(win_length), dtype=np.float64)
    elif isinstance(window, (str, tuple)) or np.isscalar(window):
        w = np.asarray(scipy.signal.get_window(window, win_length, fftbins=True), dtype=np.float64)
    else:
        w = np.asarray(window, dtype=np.float64)
        if w.shape[0] != win_length:
            raise ValueError(f"Window size mismatch: {w.shape[0]} != {win_length}")
    if win_length > n_fft:
        raise ValueError(f"Target size ({n_fft}) must be at least input size ({win_length})")
    lpad = (n_fft - win_length) // 2
    return np.pad(w, (lpad, n_fft - win_length - lpad))


@contextlib.contextmanager
def experiment_env(**env):
    """A/B experiments and tests of alternative kernels: set ``AIP_*`` switches (``None`` unsets one), make the library
    re-read them (it reads its environment once, at load: ``aip_debug_reload_env``), restore both on exit."""
    lib = _cabi.load()
    saved = {k: os.environ.get(k) for k in env}
    try:
        for k, v in env.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = str(v)
        lib.aip_debug_reload_env()
        yield
    finally:
        for k, v in saved.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v
        lib.aip_debug_reload_env()


def _require_cuda(t: torch.Tensor, name: str) -> None:
    if not t.is_cuda:
        raise RuntimeError(f"{name} must be a CUDA tensor: ml_audio_inpainting_b200 has no CPU path")


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else C.c_void_p(t.data_ptr())


def _stream() -> C.c_void_p:
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _pairs(x, B: int, device) -> Optional[torch.Tensor]:
    """[B,2] int32 device tensor from None / numpy / tensor / a single (a, b) pair."""
    if x is None:
        return None
    if isinstance(x, torch.Tensor):
        t = x.to(device=device, dtype=torch.int32)
    else:
        t = torch.as_tensor(np.asarray(x, dtype=np.int64).astype(np.int32), device=device)
    if t.ndim == 1:
        t = t.reshape(1, 2).expand(B, 2)
    if tuple(t.shape) != (B, 2):
        raise ValueError(f"expected a [{B}, 2] index array, got {tuple(t.shape)}")
    return t.contiguous()


@dataclass
class StftPlan:
    """Transform parameters + the device-resident window (and cached 1/window-sum-square tables)."""
    n_fft: int
    hop_length: int
    win_length: int
    window: object
    center: bool
    device: torch.device
    pad_mode: str = "constant"
    window_dev: torch.Tensor = field(repr=False, default=None)
    desc: StftDesc = field(repr=False, default=None)
    _inv_wss: Dict[Tuple[int, int], Tuple[torch.Tensor, "torch.cuda.Event"]] = field(repr=False, default_factory=dict)

    @property
    def n_bins(self) -> int:
        return self.n_fft // 2 + 1

    def num_frames(self, n_samples: int) -> int:
        T = _cabi.load().aip_num_frames(n_samples, self.n_fft, self.hop_length, int(self.center))
        if T < 1:
            raise ValueError(f"n_fft={self.n_fft} is too large for input signal of length={n_samples}")
        return int(T)

    def istft_length(self, n_frames: int, length: Optional[int] = None) -> int:
        return int(_cabi.load().aip_istft_length(n_frames, self.n_fft, self.hop_length, int(self.center),
                                                 int(length or 0)))

    def inv_wss(self, n_frames: int, length: Optional[int] = None) -> torch.Tensor:
        key = (int(n_frames), int(length or 0))
        hit = self._inv_wss.get(key)
        cur = torch.cuda.current_stream(self.device)
        if hit is None:
            out_len = self.istft_length(n_frames, length)
            t = torch.empty(out_len, dtype=torch.float32, device=self.device)
            with torch.cuda.device(self.device):
                check(_cabi.load().aip_inv_window_sumsquare_f32(C.byref(self.desc), key[0], key[1], _ptr(t),
                                                                out_len, C.c_void_p(cur.cuda_stream)),
                      "aip_inv_window_sumsquare_f32")
                ready = torch.cuda.Event()
                ready.record(cur)
            if len(self._inv_wss) > 64:
                self._inv_wss.clear()
            self._inv_wss[key] = (t, ready)
            return t
        t, ready = hit
        cur.wait_event(ready)      # the table was filled on whichever stream asked first: order later readers after that fill
        return t


_plans: Dict[tuple, StftPlan] = {}

# librosa.stft's centre padding when a caller does not say: "constant" (librosa >= 0.10, the oracle's choice) or
# "reflect" (librosa < 0.10) -- the reference never passes pad_mode and pins librosa>=0.8.1 only, so which one ITS results carry
# depends on the installed librosa; set AIP_LIBROSA_PAD_MODE=reflect (or this variable) to reproduce an old installation.
DEFAULT_PAD_MODE = os.environ.get("AIP_LIBROSA_PAD_MODE", "constant")


def get_plan(n_fft: int, hop_length: Optional[int] = None, win_length: Optional[int] = None, window="hann",
             center: bool = True, device=None, pad_mode: Optional[str] = None) -> StftPlan:
    """Cached plan.  Defaults follow librosa >= 0.10: win_length = n_fft, hop_length = win_length // 4, zero centre padding;
    ``pad_mode="reflect"`` is librosa < 0.10's default (the reference pins ``librosa>=0.8.1`` only) and affects the forward
    transform alone."""
    pad_mode = DEFAULT_PAD_MODE if pad_mode is None else pad_mode
    if pad_mode not in ("constant", "reflect"):
        raise NotImplementedError(f"pad_mode={pad_mode!r}: 'constant' and 'reflect' are implemented")
    lib = _cabi.load()
    device = torch.device(device if device is not None else f"cuda:{torch.cuda.current_device()}")
    if device.type != "cuda":
        raise RuntimeError("ml_audio_inpainting_b200 runs on CUDA devices only (no CPU fallback)")
    if device.index is None:
        device = torch.device("cuda", torch.cuda.current_device())
    n_fft = int(n_fft)
    win_length = int(win_length) if win_length is not None else n_fft
    hop_length = int(hop_length) if hop_length is not None else win_length // 4
    if hop_length <= 0:
        raise ValueError(f"hop_length={hop_length} must be a positive integer")
    if n_fft < 32 or n_fft > 4096 or (n_fft & (n_fft - 1)):
        raise NotImplementedError(f"n_fft={n_fft}: the CUDA kernels cover powers of two in [32, 4096]")
    wkey = window if isinstance(window, (str, tuple, float, int)) else ("array", np.asarray(window).tobytes())
    key = (n_fft, hop_length, win_length, wkey, bool(center), device.index, pad_mode)
    plan = _plans.get(key)
    if plan is None:
        with torch.cuda.device(device):
            if not lib.aip_device_supported():
                raise RuntimeError(f"{torch.cuda.get_device_name(device)} is not an sm_100 (B200) device; "
                                   "the kernels are built for sm_100a only and there is no fallback")
        w = fft_window(window, win_length, n_fft).astype(np.float32)
        wd = torch.from_numpy(w).to(device)
        desc = StftDesc(n_fft, hop_length, (2 if pad_mode == "reflect" else 1) if center else 0, win_length, wd.data_ptr())
        plan = StftPlan(n_fft, hop_length, win_length, window, bool(center), device, pad_mode, wd, desc)
        _plans[key] = plan
    return plan


def stft(wave: torch.Tensor, plan: StftPlan, *, gap_samples=None, zero_frames=None, mask_frames=None,
         mask_in_gap_is_one: bool = True, mag_kind: int = MAG_NONE, eps: float = 1e-9, power: float = 1.0,
         t_out: Optional[int] = None, want_spec: bool = True, want_phase: bool = False,
         want_mask: bool = False, out: Optional[dict] = None) -> dict:
    """Forward transform with fused epilogues.  Returns a dict with the requested outputs among
    ``spec`` (complex64 [B,F,T]), ``mag``, ``phase``, ``mask`` (float32 [B,F,T]).

    ``out`` may carry pre-allocated tensors under the same keys (benchmarks re-use buffers)."""
    _require_cuda(wave, "wave")
    if wave.dtype != torch.float32:
        wave = wave.to(torch.float32)
    squeeze = wave.ndim == 1
    if squeeze:
        wave = wave.unsqueeze(0)
    if wave.ndim != 2:
        raise ValueError("wave must be [L] or [B, L]")
    if wave.stride(1) != 1:
        wave = wave.contiguous()
    B, L = wave.shape
    T = plan.num_frames(L)
    t_out = T if t_out is None else min(int(t_out), T)
    F = plan.n_bins
    dev = wave.device
    out = dict(out) if out else {}

    def buf(name, dtype):
        t = out.get(name)
        if t is None:
            t = torch.empty((B, F, t_out), dtype=dtype, device=dev)
        elif tuple(t.shape) != (B, F, t_out) or t.dtype != dtype or not t.is_contiguous():
            raise ValueError(f"out[{name!r}] has the wrong shape/dtype/layout")
        return t

    spec = buf("spec", torch.complex64) if want_spec else None
    mag = buf("mag", torch.float32) if mag_kind != MAG_NONE else None
    phase = buf("phase", torch.float32) if want_phase else None
    mask = buf("mask", torch.float32) if want_mask else None
    gaps = _pairs(gap_samples, B, dev)
    zf = _pairs(zero_frames, B, dev)
    mf = _pairs(mask_frames, B, dev)
    lib = _cabi.load()
    with torch.cuda.device(dev):
        # The dense frame mask depends on the frame index only.  The forward kernel can emit it from its epilogue
        # (mask_out of aip_stft_fwd_f32), but its stage-2 warps are the critical role: 0.561 ms with the mask fused against
        # 0.21 ms + 0.07 ms for the streaming aip_frame_mask_f32 kernel (1024 x 5 s clips), so it is written separately.
        check(lib.aip_stft_fwd_f32(
            C.byref(plan.desc), _ptr(wave), B, L, wave.stride(0), _ptr(gaps), _ptr(zf), None,
            int(bool(mask_in_gap_is_one)), int(mag_kind), float(eps), float(power), t_out,
            _ptr(spec), _ptr(mag), _ptr(phase), None, _stream()), "aip_stft_fwd_f32")
        if mask is not None:
            check(lib.aip_frame_mask_f32(_ptr(mask), B, F, t_out, _ptr(mf), int(bool(mask_in_gap_is_one)), _stream()),
                  "aip_frame_mask_f32")
    res = {}
    for name, t in (("spec", spec), ("mag", mag), ("phase", phase), ("mask", mask)):
        if t is not None:
            res[name] = t[0] if squeeze else t
    return res


def stft_gap_variants(wave: torch.Tensor, plan: StftPlan, gap_samples, variants_per_row: int, *,
                      mag_kind: int = MAG_LOG10_EPS, eps: float = 1e-9, t_out: Optional[int] = None,
                      clean_mag: Optional[torch.Tensor] = None, out: Optional[torch.Tensor] = None,
                      gap_len_max: Optional[int] = None) -> dict:
    """``variants_per_row`` gapped magnitude spectrograms per row of ``wave`` [N, L] (the gaps_per_audio loop of
    models/CNNBLSTM/dataset.py:93-111) without ``variants_per_row`` full transforms: one clean transform per row, a
    streaming copy, and a re-transform of the few frames each gap touches (``aip_stft_gap_variants_f32``).

    ``gap_samples``: int [N * variants_per_row, 2] sample ranges (host array or CUDA int32 tensor), variant
    ``i * variants_per_row + j`` belongs to row ``i``.  ``gap_len_max``: an upper bound of the gap lengths (computed from
    ``gap_samples`` when omitted -- for a CUDA tensor that costs a device synchronisation).  Returns ``mag`` f32
    [N * variants_per_row, F, t_out] and ``clean_mag`` f32 [N, F, t_out].  Bit-identical to
    ``stft(wave.repeat_interleave(G, 0), gap_samples=...)``."""
    _require_cuda(wave, "wave")
    if plan.n_fft != 512:
        raise NotImplementedError("gap variants run on the n_fft = 512 register-FFT kernels only")
    if wave.dtype != torch.float32:
        wave = wave.to(torch.float32)
    if wave.ndim == 1:
        wave = wave.unsqueeze(0)
    if wave.stride(1) != 1:
        wave = wave.contiguous()
    N, L = wave.shape
    G = int(variants_per_row)
    if G < 1:
        raise ValueError("variants_per_row must be >= 1")
    T = plan.num_frames(L)
    t_out = T if t_out is None else min(int(t_out), T)
    F = plan.n_bins
    dev = wave.device
    if gap_len_max is not None:
        gmax = int(gap_len_max)
    elif isinstance(gap_samples, torch.Tensor):
        gmax = int((gap_samples[:, 1] - gap_samples[:, 0]).max().item()) if gap_samples.numel() else 0
    else:
        g_np = np.asarray(gap_samples, dtype=np.int64).reshape(-1, 2)
        gmax = int((g_np[:, 1] - g_np[:, 0]).max()) if g_np.size else 0
    gaps = _pairs(gap_samples, N * G, dev)
    if clean_mag is None:
        clean_mag = stft(wave, plan, mag_kind=mag_kind, eps=eps, t_out=t_out, want_spec=False)["mag"]
    elif tuple(clean_mag.shape) != (N, F, t_out) or clean_mag.dtype != torch.float32 or not clean_mag.is_contiguous():
        raise ValueError("clean_mag has the wrong shape/dtype/layout")
    if out is None:
        out = torch.empty((N * G, F, t_out), dtype=torch.float32, device=dev)
    elif tuple(out.shape) != (N * G, F, t_out) or out.dtype != torch.float32 or not out.is_contiguous():
        raise ValueError("out has the wrong shape/dtype/layout")
    lib = _cabi.load()
    ws_bytes = int(lib.aip_stft_gap_variants_workspace_bytes(N, G))
    ws = torch.empty((max(ws_bytes, 16) + 3) // 4, dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        check(lib.aip_stft_gap_variants_f32(
            C.byref(plan.desc), _ptr(wave), N, L, wave.stride(0), G, _ptr(gaps), max(gmax, 0), int(mag_kind), float(eps),
            t_out, _ptr(clean_mag), _ptr(out), _ptr(ws), ws_bytes, _stream()), "aip_stft_gap_variants_f32")
    return {"mag": out, "clean_mag": clean_mag}


def db_heuristic(x: torch.Tensor) -> torch.Tensor:
    """Per clip: (max < 0 and mean < 0), the test utils.spectrogram_to_audio applies (utils.py:313-314)."""
    _require_cuda(x, "x")
    x = x.contiguous().to(torch.float32)
    B = x.shape[0]
    flags = torch.empty(B, dtype=torch.int32, device=x.device)
    with torch.cuda.device(x.device):
        check(_cabi.load().aip_db_heuristic_f32(_ptr(x), B, x[0].numel(), _ptr(flags), _stream()),
              "aip_db_heuristic_f32")
    return flags


def istft(plan: StftPlan, spec: Optional[torch.Tensor] = None, mag: Optional[torch.Tensor] = None,
          phase: Optional[torch.Tensor] = None, mag_domain: int = DOM_LINEAR, db_auto: bool = False,
          length: Optional[int] = None, out: Optional[torch.Tensor] = None, normalize: bool = False,
          peaks_out: Optional[torch.Tensor] = None, pcm16: bool = False) -> torch.Tensor:
    """Inverse transform: complex ``spec`` or ``mag`` (+ ``phase``) -> waveform [B, out_len] float32.

    ``db_auto`` applies the reference's per-clip dB test to ``mag`` on the device (no host sync).
    ``normalize`` additionally peak-normalises every clip like ``librosa.util.normalize`` in the reference's
    ``save_audio`` (utils.py:84), with the per-clip peak taken inside the inverse kernel's overlap-add
    (``aip_istft_normalized_f32``); ``peaks_out`` [B] float32 receives the peaks before scaling.
    ``pcm16`` returns int16 [B, out_len] instead: the 16-bit samples ``save_audio`` writes into its FLAC (utils.py:87), produced
    by the pass that would have scaled the waveform (``out`` then keeps the un-normalised float waveform)."""
    src = spec if spec is not None else mag
    if src is None:
        raise ValueError("istft needs spec or mag")
    _require_cuda(src, "spectrogram")
    squeeze = src.ndim == 2
    lib = _cabi.load()

    def prep(t, dtype):
        if t is None:
            return None
        if t.ndim == 2:
            t = t.unsqueeze(0)
        return t.to(dtype).contiguous()

    spec = prep(spec, torch.complex64)
    mag = prep(mag, torch.float32) if spec is None else None
    phase = prep(phase, torch.float32) if spec is None else None
    ref = spec if spec is not None else mag
    B, F, T = ref.shape
    if F != plan.n_bins:
        raise ValueError(f"expected {plan.n_bins} frequency bins for n_fft={plan.n_fft}, got {F}")
    dev = ref.device
    out_len = plan.istft_length(T, length)
    if out is None:
        out = torch.empty((B, out_len), dtype=torch.float32, device=dev)
    elif tuple(out.shape) != (B, out_len) or out.dtype != torch.float32 or out.stride(1) != 1:
        raise ValueError("out has the wrong shape/dtype/layout")
    inv = plan.inv_wss(T, length)
    flags = db_heuristic(mag) if (db_auto and mag is not None) else None
    ws_bytes = int(lib.aip_istft_workspace_bytes(C.byref(plan.desc), B, T))
    ws = torch.empty(max(ws_bytes, 4) // 4, dtype=torch.float32, device=dev) if ws_bytes else None
    pcm = torch.empty((B, out_len), dtype=torch.int16, device=dev) if pcm16 else None
    with torch.cuda.device(dev):
        if normalize:
            if peaks_out is None:
                peaks_out = torch.empty(B, dtype=torch.float32, device=dev)
            elif tuple(peaks_out.shape) != (B,) or peaks_out.dtype != torch.float32 or not peaks_out.is_contiguous():
                raise ValueError("peaks_out must be a contiguous float32 [B] tensor")
            check(lib.aip_istft_normalized_f32(C.byref(plan.desc),
                                               _ptr(spec.view(torch.float32) if spec is not None else None),
                                               _ptr(mag), _ptr(phase), int(mag_domain), _ptr(flags), B, T,
                                               int(length or 0), _ptr(inv), _ptr(out), out.stride(0), _ptr(peaks_out),
                                               _ptr(pcm), out_len, _ptr(ws), ws_bytes, _stream()),
                  "aip_istft_normalized_f32")
        else:
            check(lib.aip_istft_f32(C.byref(plan.desc), _ptr(spec.view(torch.float32) if spec is not None else None),
                                    _ptr(mag), _ptr(phase), int(mag_domain), _ptr(flags), B, T, int(length or 0),
                                    _ptr(inv), _ptr(out), out.stride(0), _ptr(ws), ws_bytes, _stream()),
                  "aip_istft_f32")
            if pcm16 and out_len > 0:
                check(lib.aip_wave_to_pcm16_f32(_ptr(out), out.stride(0), _ptr(pcm), out_len, B, out_len, PCM_RAW, None,
                                                _stream()), "aip_wave_to_pcm16_f32")
    res = pcm if pcm16 else out
    return res[0] if squeeze else res


PCM_RAW, PCM_NORMALIZE, PCM_PEAKS_GIVEN = 0, 1, 2


def wave_to_pcm16(wave: torch.Tensor, normalize: bool = True, peaks: Optional[torch.Tensor] = None,
                  peaks_out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """``save_audio``'s tail on the device (utils.py:83-87): ``librosa.util.normalize`` (``normalize``) and soundfile's
    float -> PCM_16 conversion for FLAC (x * 32768, round half to even, clip) -> int16 of the same shape.  ``peaks`` [B]: the
    clips' max|x| when already known (skips the peak pass); ``peaks_out`` receives the peaks taken here."""
    _require_cuda(wave, "wave")
    squeeze = wave.ndim == 1
    w = (wave.unsqueeze(0) if squeeze else wave).to(torch.float32)
    if w.ndim != 2:
        raise ValueError("wave must be [L] or [B, L]")
    if w.stride(1) != 1:
        w = w.contiguous()
    B, L = w.shape
    pcm = torch.empty((B, L), dtype=torch.int16, device=w.device)
    mode, pk = PCM_RAW, None
    if peaks is not None:
        if tuple(peaks.shape) != (B,) or peaks.dtype != torch.float32 or not peaks.is_contiguous() or peaks.device != w.device:
            raise ValueError("peaks must be a contiguous float32 [B] tensor on the waveform's device")
        mode, pk = PCM_PEAKS_GIVEN, peaks
    elif normalize:
        pk = peaks_out if peaks_out is not None else torch.empty(B, dtype=torch.float32, device=w.device)
        if tuple(pk.shape) != (B,) or pk.dtype != torch.float32 or not pk.is_contiguous():
            raise ValueError("peaks_out must be a contiguous float32 [B] tensor")
        mode = PCM_NORMALIZE
    with torch.cuda.device(w.device):
        check(_cabi.load().aip_wave_to_pcm16_f32(_ptr(w), w.stride(0), _ptr(pcm), L, B, L, mode, _ptr(pk), _stream()),
              "aip_wave_to_pcm16_f32")
    return pcm[0] if squeeze else pcm


def istft_blend(plan: StftPlan, model_out: torch.Tensor, blend_in: torch.Tensor, blend_mask: torch.Tensor,
                phase: torch.Tensor, mag_domain: int = DOM_POW10, length: Optional[int] = None,
                mask_keeps_input: bool = False, normalize: bool = False,
                peaks_out: Optional[torch.Tensor] = None, out: Optional[torch.Tensor] = None,
                pcm16: bool = False) -> torch.Tensor:
    """The model hand-off in one kernel (``aip_istft_handoff_f32``): blend, un-log, phase reuse, istft.

    ``mask_keeps_input=False``: ``m = model_out * mask + blend_in * (1 - mask)`` -- the CNN-BLSTM convention, mask 1 inside
    the gap (reference models/CNNBLSTM/model.py:108; ``mag_domain=DOM_POW10`` adds models/model_eval.py:163).
    ``mask_keeps_input=True``: ``m = model_out * (1 - mask) + blend_in * mask`` -- the GAN convention, mask 1 outside the gap
    (``combined_log_mag``, models/GAN/train.py:473; the reference then hands the log1p-domain blend to
    ``spectrogram_to_audio`` as it is: ``mag_domain=DOM_LINEAR``; ``DOM_EXPM1`` undoes the log1p instead).
    ``normalize`` adds ``save_audio``'s peak normalisation (utils.py:84); ``pcm16`` returns the int16 samples of its 16-bit
    FLAC (utils.py:87) instead of the float waveform (see ``istft``)."""
    ts = []
    for name, t in (("model_out", model_out), ("blend_in", blend_in), ("blend_mask", blend_mask), ("phase", phase)):
        _require_cuda(t, name)
        t = t.to(torch.float32)
        ts.append((t.unsqueeze(0) if t.ndim == 2 else t).contiguous())
    squeeze = model_out.ndim == 2
    B, F, T = ts[0].shape
    if any(tuple(t.shape) != (B, F, T) for t in ts):
        raise ValueError("model_out, blend_in, blend_mask and phase must have the same [B, F, T] shape")
    if F != plan.n_bins:
        raise ValueError(f"expected {plan.n_bins} frequency bins for n_fft={plan.n_fft}, got {F}")
    dev = ts[0].device
    lib = _cabi.load()
    out = _out_buf(out, (B, plan.istft_length(T, length)), dev)
    inv = plan.inv_wss(T, length)
    ws_bytes = int(lib.aip_istft_workspace_bytes(C.byref(plan.desc), B, T))
    ws = torch.empty(max(ws_bytes, 4) // 4, dtype=torch.float32, device=dev) if ws_bytes else None
    peaks = None
    if normalize:
        peaks = peaks_out if peaks_out is not None else torch.empty(B, dtype=torch.float32, device=dev)
        if tuple(peaks.shape) != (B,) or peaks.dtype != torch.float32 or not peaks.is_contiguous():
            raise ValueError("peaks_out must be a contiguous float32 [B] tensor")
    pcm = torch.empty((B, out.shape[1]), dtype=torch.int16, device=dev) if pcm16 else None
    with torch.cuda.device(dev):
        check(lib.aip_istft_handoff_f32(C.byref(plan.desc), _ptr(ts[0]), _ptr(ts[1]), _ptr(ts[2]), int(bool(mask_keeps_input)),
                                        _ptr(ts[3]), int(mag_domain), B, T, int(length or 0), _ptr(inv), _ptr(out),
                                        out.stride(0), _ptr(peaks), _ptr(pcm), out.shape[1], _ptr(ws), ws_bytes, _stream()),
              "aip_istft_handoff_f32")
    res = pcm if pcm16 else out
    return res[0] if squeeze else res


def griffinlim(plan: StftPlan, mag: torch.Tensor, n_iter: int = 32, momentum: float = 0.99,
               init_angles: Optional[torch.Tensor] = None, generator: Optional[torch.Generator] = None,
               init: Optional[str] = "random") -> torch.Tensor:
    """librosa.griffinlim: ``init_angles`` (unit phasors, complex64 [B,F,T]) makes it deterministic;
    otherwise uniform random phases are drawn on the device (``init='random'``) or all ones (``None``)."""
    _require_cuda(mag, "mag")
    if momentum < 0:
        raise ValueError(f"griffinlim() called with momentum={momentum} < 0")
    squeeze = mag.ndim == 2
    if squeeze:
        mag = mag.unsqueeze(0)
    is_cplx = torch.is_complex(mag)      # librosa multiplies by whatever it is handed (tests/utils_test.py:624-645)
    mag = mag.to(torch.complex64 if is_cplx else torch.float32).contiguous()
    B, F, T = mag.shape
    dev = mag.device
    if init_angles is not None:
        ang = init_angles.to(device=dev, dtype=torch.complex64).reshape(B, F, T).contiguous().clone()
    elif init == "random":
        # one 62-bit key per call from torch's generator (the given one, else the global CPU generator: torch.manual_seed makes the
        # call reproducible), the phases themselves from the device's counter RNG -- no [B, F, T] temporaries, 8 bytes per bin written
        gen_dev = generator.device if generator is not None else "cpu"
        seed = int(torch.randint(0, 2 ** 62, (1,), generator=generator, device=gen_dev).item())
        ang = torch.empty((B, F, T), dtype=torch.complex64, device=dev)
        with torch.cuda.device(dev):
            check(_cabi.load().aip_random_phasors_f32(_ptr(ang.view(torch.float32)), B * F * T, seed, _stream()),
                  "aip_random_phasors_f32")
    elif init is None:
        ang = torch.ones((B, F, T), dtype=torch.complex64, device=dev)
    else:
        raise ValueError(f"init={init} must either None or 'random'")
    tprev = torch.empty_like(ang)
    out_len = plan.istft_length(T)
    out = torch.empty((B, out_len), dtype=torch.float32, device=dev)
    inv = plan.inv_wss(T)
    lib = _cabi.load()
    # istft workspace (generic n_fft only) + one more complex [B,F,T] array: rebuilt spectra ping-pong, no tprev copy
    ws_bytes = ((int(lib.aip_istft_workspace_bytes(C.byref(plan.desc), B, T)) + 15) // 16) * 16 + B * F * T * 8
    ws = torch.empty(ws_bytes // 4, dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        if is_cplx:
            check(lib.aip_griffinlim_c64_f32(C.byref(plan.desc), _ptr(torch.view_as_real(mag)), _ptr(ang.view(torch.float32)),
                                             _ptr(tprev.view(torch.float32)), B, T, int(n_iter), float(momentum),
                                             _ptr(inv), _ptr(out), out.stride(0), _ptr(ws), ws_bytes, _stream()),
                  "aip_griffinlim_c64_f32")
        else:
            check(lib.aip_griffinlim_f32(C.byref(plan.desc), _ptr(mag), _ptr(ang.view(torch.float32)),
                                         _ptr(tprev.view(torch.float32)), B, T, int(n_iter), float(momentum),
                                         _ptr(inv), _ptr(out), out.stride(0), _ptr(ws), ws_bytes, _stream()),
                  "aip_griffinlim_f32")
    return out[0] if squeeze else out


# --- mel (librosa.filters.mel / feature.melspectrogram; reference utils.py:236-277, :335-393) -----------------------

_mel_cache: Dict[tuple, tuple] = {}


def mel_basis(sr: float, n_fft: int, n_mels: int = 128, fmin: float = 0.0, fmax: Optional[float] = None) -> np.ndarray:
    """librosa.filters.mel(sr, n_fft, n_mels, fmin, fmax) with its defaults (Slaney scale, norm='slaney', float32):
    a host-side parameter table like the window [n_mels, 1 + n_fft // 2]."""
    if fmax is None:
        fmax = float(sr) / 2

    f_sp, min_log_hz, logstep = 200.0 / 3, 1000.0, np.log(6.4) / 27.0
    min_log_mel = min_log_hz / f_sp          # librosa's expression: 14.999999999999998, not 15.0

    def hz_to_mel(f):
        f = np.asanyarray(f, dtype=np.float64)
        return np.where(f >= min_log_hz, min_log_mel + np.log(np.maximum(f, 1e-300) / min_log_hz) / logstep, f / f_sp)

    def mel_to_hz(m):
        m = np.asanyarray(m, dtype=np.float64)
        return np.where(m >= min_log_mel, min_log_hz * np.exp(logstep * (m - min_log_mel)), f_sp * m)

    n_bins = 1 + n_fft // 2
    fftfreqs = np.fft.rfftfreq(n=n_fft, d=1.0 / sr)
    mel_f = mel_to_hz(np.linspace(hz_to_mel(fmin), hz_to_mel(fmax), int(n_mels) + 2))
    fdiff = np.diff(mel_f)
    ramps = np.subtract.outer(mel_f, fftfreqs)
    weights = np.zeros((int(n_mels), n_bins), dtype=np.float32)          # float32 from the start, as librosa builds it
    for i in range(int(n_mels)):
        weights[i] = np.maximum(0, np.minimum(-ramps[i] / fdiff[i], ramps[i + 2] / fdiff[i + 1]))
    weights *= (2.0 / (mel_f[2:int(n_mels) + 2] - mel_f[:int(n_mels)]))[:, np.newaxis]
    return weights


def mel_bands(basis: np.ndarray) -> np.ndarray:
    """Per row of a filter bank the bin range [f0, f1) that holds all of its non-zero entries (int32 [n_mels, 2];
    an all-zero row gets the empty range (0, 0)): what ``aip_mel_project_f32`` loops over instead of all F bins."""
    nz = np.asarray(basis) != 0
    any_ = nz.any(1)
    f0 = np.where(any_, nz.argmax(1), 0)
    f1 = np.where(any_, basis.shape[1] - nz[:, ::-1].argmax(1), 0)
    return np.stack([f0, f1], 1).astype(np.int32)


def _mel_tables(sr, n_fft, n_mels, fmin, fmax, device):
    key = (float(sr), int(n_fft), int(n_mels), float(fmin), None if fmax is None else float(fmax), device.index)
    hit = _mel_cache.get(key)
    if hit is None:
        w = mel_basis(sr, n_fft, n_mels, fmin, fmax)
        bands = mel_bands(w)
        inv = np.linalg.pinv(w).astype(np.float32)                       # utils.py:375 (host: a parameter table)
        hit = (torch.from_numpy(w).to(device), torch.from_numpy(bands).to(device), torch.from_numpy(np.ascontiguousarray(inv)).to(device))
        if len(_mel_cache) > 16:
            _mel_cache.clear()
        _mel_cache[key] = hit
    return hit


def _out_buf(out: Optional[torch.Tensor], shape, device) -> torch.Tensor:
    if out is None:
        return torch.empty(shape, dtype=torch.float32, device=device)
    if tuple(out.shape) != tuple(shape) or out.dtype != torch.float32 or not out.is_contiguous() or out.device != device:
        raise ValueError("out has the wrong shape/dtype/layout/device")
    return out


def mel_project(spec_pow: torch.Tensor, sr: float, n_fft: int, n_mels: int = 128, fmin: float = 0.0,
                fmax: Optional[float] = None, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """mel[b, m, t] = sum_f basis[m, f] * spec_pow[b, f, t] (``aip_mel_project_f32``): the contraction of
    librosa.feature.melspectrogram (utils.py:268-277) on ``|stft| ** power`` [B, F, T] -> [B, n_mels, T]."""
    _require_cuda(spec_pow, "spec_pow")
    squeeze = spec_pow.ndim == 2
    s = (spec_pow.unsqueeze(0) if squeeze else spec_pow).to(torch.float32).contiguous()
    B, F, T = s.shape
    if F != 1 + n_fft // 2:
        raise ValueError(f"expected {1 + n_fft // 2} frequency bins for n_fft={n_fft}, got {F}")
    basis, bands, _ = _mel_tables(sr, n_fft, n_mels, fmin, fmax, s.device)
    out = _out_buf(out, (B, int(n_mels), T), s.device)
    with torch.cuda.device(s.device):
        check(_cabi.load().aip_mel_project_f32(_ptr(basis), _ptr(bands), _ptr(s), B, F, T, int(n_mels), _ptr(out), _stream()),
              "aip_mel_project_f32")
    return out[0] if squeeze else out


def mel_inverse(mel: torch.Tensor, sr: float, n_fft: int, n_mels: int = 128, fmin: float = 0.0,
                fmax: Optional[float] = None, take_sqrt: bool = False, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """pinv(mel basis) @ mel (+ sqrt for power spectrograms): utils.py:375-383 on the device (``aip_mel_inverse_f32``).
    Negative projections become NaN under the square root as ``np.sqrt`` makes them in the reference -- except those below
    the noise floor of an fp32 power spectrogram (<= 1e-9 of the clip's largest projection), which are taken as 0."""
    _require_cuda(mel, "mel")
    squeeze = mel.ndim == 2
    m = (mel.unsqueeze(0) if squeeze else mel).to(torch.float32).contiguous()
    B, M, T = m.shape
    if M != int(n_mels):
        raise ValueError(f"shapes ({1 + n_fft // 2},{int(n_mels)}) and ({M},{T}) not aligned")
    F = 1 + n_fft // 2
    _, _, inv = _mel_tables(sr, n_fft, n_mels, fmin, fmax, m.device)
    out = _out_buf(out, (B, F, T), m.device)
    peaks = torch.empty(B, dtype=torch.float32, device=m.device) if take_sqrt else None
    with torch.cuda.device(m.device):
        check(_cabi.load().aip_mel_inverse_f32(_ptr(inv), _ptr(m), B, F, T, int(n_mels), int(bool(take_sqrt)), _ptr(out),
                                               _ptr(peaks), _stream()), "aip_mel_inverse_f32")
    return out[0] if squeeze else out
