"""Batched front-ends / back-end: the per-item numpy epilogues of the reference's callers, fused
into the transform kernels and run for a whole batch of clips per launch.

  cnnblstm_batch      models/CNNBLSTM/dataset.py:89-119  (log10(|S_gap|+1e-9), complex target, frame mask)
  cnnblstm_dataset_batch  models/CNNBLSTM/dataset.py:74-121  (a whole __getitem__: gaps_per_audio gaps per file from ONE clean
                      transform; only the frames a gap touches are re-transformed)
  gan_batch           models/GAN/dataset.py:104-166      (log1p magnitudes, phase, 0-in-gap mask)
  eval_cnnlstm_batch  models/model_eval.py:146-154       (spectrum-domain gap)
  eval_gan_batch      models/model_eval.py:61-111
  backend_batch       models/model_eval.py:131-140, :180-189 -> utils.spectrogram_to_audio(phase=...)
  HostPipeline        host-buffer entry (pinned in / pinned out), H2D | kernel | D2H overlapped on 4 streams

All random draws come from the GLOBAL ``np.random`` stream in the reference's order (one draw per
item; a vectorised ``randint(size=B)`` yields the same numbers as B scalar draws), and every index
(gap samples, frame ranges) is computed on the host in float64 / int64 exactly as the reference does
(see ``gaps.py``); only index pairs travel to the device.
"""
from __future__ import annotations

from typing import Optional

import numpy as np
import torch

from . import gaps, spectral as sp

__all__ = ["cnnblstm_batch", "cnnblstm_dataset_batch", "gan_batch", "eval_cnnlstm_batch", "eval_gan_batch", "backend_batch", "cnnblstm_backend_batch",
           "HostPipeline"]


def _as_batch(wave: torch.Tensor) -> torch.Tensor:
    if not wave.is_cuda:
        raise RuntimeError("wave must be a CUDA tensor: there is no CPU path")
    return wave.unsqueeze(0) if wave.ndim == 1 else wave


def cnnblstm_batch(wave: torch.Tensor, gap_len_s: float = 0.2, sample_rate: int = 16000, max_len_s: float = 5.0,
                   n_fft: int = 512, hop_len: int = 192, win_len: int = 384, starts: Optional[np.ndarray] = None,
                   want_target: bool = True, want_mask: bool = True) -> dict:
    """One gap per row of ``wave`` [B, L] (rows may repeat a clip ``gaps_per_audio`` times).

    Returns ``spectrogram_gap`` f32 [B,F,Tc], ``gap_int_s`` f32 [B,2] (host), ``gap_mask`` f32 [B,F,Tc],
    ``spectrogram_target_phase`` c64 [B,F,Tc], plus the exact integer ranges used."""
    wave = _as_batch(wave)
    B, L = wave.shape
    g = gaps.gap_len_samples(gap_len_s, sample_rate)
    if g >= L:
        raise ValueError(f"Gap length ({g}s) exceeds audio length ({L / sample_rate}s)")   # utils.py:175-176
    if starts is None:
        starts = gaps.draw_starts_exclusive(L, g, B)                                         # utils.py:179
    starts = np.asarray(starts, dtype=np.int64)
    t0, t1 = gaps.seconds_interval(starts, g, sample_rate)                                  # utils.py:186
    f0 = gaps.time_to_frames(t0, sample_rate, hop_len)                                      # dataset.py:116
    f1 = gaps.time_to_frames(t1, sample_rate, hop_len)                                      # dataset.py:117
    plan = sp.get_plan(n_fft, hop_len, win_len, "hann", True, wave.device)
    t_crop = gaps.cnnblstm_crop_frames(sample_rate, max_len_s, hop_len)                     # dataset.py:89
    sam = np.stack([starts, starts + g], 1)
    frm = np.stack([np.atleast_1d(f0), np.atleast_1d(f1)], 1)
    res = sp.stft(wave, plan, gap_samples=sam, mask_frames=frm, mask_in_gap_is_one=True,
                  mag_kind=sp.MAG_LOG10_EPS, eps=1e-9, t_out=t_crop, want_spec=False, want_mask=want_mask)
    out = {"spectrogram_gap": res["mag"],
           "gap_int_s": np.stack([t0, t1], 1).astype(np.float32),                           # dataset.py:112
           "gap_samples": sam, "gap_frames": frm}
    if want_mask:
        out["gap_mask"] = res["mask"]
    if want_target:
        out["spectrogram_target_phase"] = sp.stft(wave, plan, t_out=t_crop)["spec"]        # dataset.py:102,110
    return out


def cnnblstm_dataset_batch(wave: torch.Tensor, gaps_per_audio: int = 25, gap_len_s: float = 0.2, sample_rate: int = 16000,
                           max_len_s: float = 5.0, n_fft: int = 512, hop_len: int = 192, win_len: int = 384,
                           starts: Optional[np.ndarray] = None, want_mask: bool = True) -> dict:
    """``LibriSpeechDataset.__getitem__`` (models/CNNBLSTM/dataset.py:74-121) for a batch of files: row ``i`` of ``wave``
    [N, L] is one decoded file (``utils.load_audio`` output), and the item holds ``gaps_per_audio`` random gaps of it.

    The reference decodes the file twice and runs two full STFTs per gap (dataset.py:93-103).  Here the clean transform
    runs ONCE per file (complex target + clean log-magnitude); a gap changes only the ~(gap + n_fft) / hop frames whose
    span meets it, so each variant is a streaming copy of the clean log-magnitude plus a re-transform of those frames
    (``spectral.stft_gap_variants``).  Draw order: one ``np.random.randint(0, L - g)`` per (file, gap) in the reference's
    order -- file-major, gap-minor (utils.py:179).

    Returns ``spectrogram_gaps`` f32 [N, G, F, Tc], ``gap_ints`` f32 [N, G, 2] (host), ``gap_masks`` f32 [N, G, F, Tc],
    ``spectrogram_target_phases`` c64 [N, G, F, Tc] (an expanded VIEW of the one clean spectrogram per file: the reference
    stores G identical copies, dataset.py:110), plus ``gap_samples`` / ``gap_frames`` int64 [N, G, 2]."""
    wave = _as_batch(wave)
    N, L = wave.shape
    G = int(gaps_per_audio)
    g = gaps.gap_len_samples(gap_len_s, sample_rate)
    if g >= L:
        raise ValueError(f"Gap length ({g}s) exceeds audio length ({L / sample_rate}s)")   # utils.py:175-176
    if starts is None:
        starts = gaps.draw_starts_exclusive(L, g, N * G)                                     # utils.py:179
    starts = np.asarray(starts, dtype=np.int64).reshape(N * G)
    t0, t1 = gaps.seconds_interval(starts, g, sample_rate)                                  # utils.py:186
    f0 = np.atleast_1d(gaps.time_to_frames(t0, sample_rate, hop_len))                       # dataset.py:116
    f1 = np.atleast_1d(gaps.time_to_frames(t1, sample_rate, hop_len))                       # dataset.py:117
    plan = sp.get_plan(n_fft, hop_len, win_len, "hann", True, wave.device)
    t_crop = min(gaps.cnnblstm_crop_frames(sample_rate, max_len_s, hop_len), plan.num_frames(L))   # dataset.py:89
    sam = np.stack([starts, starts + g], 1)
    frm = np.stack([f0, f1], 1)
    F = plan.n_bins
    target = sp.stft(wave, plan, t_out=t_crop)["spec"]                                      # dataset.py:102,110 (once per file)
    # measured per variant (5 s clips): 0.195 us for a full gapped transform against 0.116 us + 0.32 us / G for copy + re-transform
    # + the file's clean transform => the variant path pays off from 5 gaps per file (both are bit-identical)
    if n_fft == 512 and G >= 5:
        mag = sp.stft_gap_variants(wave, plan, sam, G, mag_kind=sp.MAG_LOG10_EPS, eps=1e-9, t_out=t_crop, gap_len_max=g)["mag"]
    else:       # few gaps per file, or the generic kernels (no variant mode): G full transforms
        mag = sp.stft(wave.repeat_interleave(G, 0), plan, gap_samples=sam, mag_kind=sp.MAG_LOG10_EPS, eps=1e-9,
                      t_out=t_crop, want_spec=False)["mag"]
    out = {"spectrogram_gaps": mag.view(N, G, F, t_crop),
           "gap_ints": np.stack([t0, t1], 1).astype(np.float32).reshape(N, G, 2),            # dataset.py:112
           "spectrogram_target_phases": target.unsqueeze(1).expand(N, G, F, t_crop),
           "gap_samples": sam.reshape(N, G, 2), "gap_frames": frm.reshape(N, G, 2)}
    if want_mask:
        mask = torch.empty((N * G, F, t_crop), dtype=torch.float32, device=wave.device)
        lib = sp._cabi.load()
        with torch.cuda.device(wave.device):
            sp.check(lib.aip_frame_mask_f32(sp._ptr(mask), N * G, F, t_crop, sp._ptr(sp._pairs(frm, N * G, wave.device)), 1,
                                            sp._stream()), "aip_frame_mask_f32")            # dataset.py:115-119
        out["gap_masks"] = mask.view(N, G, F, t_crop)
    return out


def gan_batch(wave: torch.Tensor, gap_len_s: float = 0.2, sample_rate: int = 16000, n_fft: int = 512,
              hop_length: int = 128, win_length: int = 512, window="hann", power: float = 1.0,
              spec_normalize: bool = True, gap_start_s: Optional[float] = None,
              starts: Optional[np.ndarray] = None) -> dict:
    """SpeechInpaintingDataset.__getitem__ steps 2-5 for a batch (without the channel dimension)."""
    wave = _as_batch(wave)
    B, L = wave.shape
    g = gaps.gap_len_samples(gap_len_s, sample_rate)
    if g <= 0:                                                                               # utils.py:122-124
        sam = np.zeros((B, 2), dtype=np.int64)
    elif g >= L:                                                                             # utils.py:126-129
        print(f"Warning: Gap length ({gap_len_s}s) >= audio length. Returning all zeros mask.")
        sam = np.tile(np.array([[0, L]], dtype=np.int64), (B, 1))
    else:
        if starts is None:
            if gap_start_s is None:
                starts = gaps.draw_starts_inclusive(L, g, B)                                 # utils.py:132-134
            else:
                starts = np.full(B, int(gap_start_s * sample_rate), dtype=np.int64)          # utils.py:136
        starts = np.asarray(starts, dtype=np.int64)
        sam = np.stack([starts, starts + g], 1)
    plan = sp.get_plan(n_fft, hop_length, win_length, window, True, wave.device)
    T = plan.num_frames(L)
    f0, f1 = gaps.gan_frame_range(sam[:, 0], sam[:, 1], hop_length, T)                      # GAN/dataset.py:138-147
    frm = np.stack([f0, np.maximum(f0, f1)], 1)
    kind = sp.MAG_LOG1P_POW if spec_normalize else sp.MAG_POW
    orig = sp.stft(wave, plan, mag_kind=kind, power=power, want_spec=False, want_phase=True,
                   mask_frames=frm, mask_in_gap_is_one=False, want_mask=True)               # :112-123, :150-152
    imp = sp.stft(wave, plan, gap_samples=sam, mag_kind=sp.MAG_LOG1P_POW if spec_normalize else sp.MAG_ABS,
                  power=1.0, want_spec=False)                                               # :126-135
    return {"original_magnitude": orig["mag"], "original_phase": orig["phase"], "mask": orig["mask"],
            "impaired_magnitude": imp["mag"], "gap_samples": sam, "gap_frames": frm}


def eval_cnnlstm_batch(wave: torch.Tensor, sample_rate: int = 16000, n_fft: int = 512, hop_length: int = 192,
                       win_length: int = 384, t0: float = 2.0, t1: float = 2.08) -> dict:
    """models/model_eval.py:146-154 for a batch: log10(|S * (1 - mask)| + 1e-9) with a frame-domain gap."""
    wave = _as_batch(wave)
    B = wave.shape[0]
    plan = sp.get_plan(n_fft, hop_length, win_length, "hann", True, wave.device)
    f0 = gaps.time_to_frames(t0, sample_rate, hop_length)                                    # :148
    f1 = gaps.time_to_frames(t1, sample_rate, hop_length)                                    # :149
    frm = np.tile(np.array([[f0, f1]], dtype=np.int64), (B, 1))
    full = sp.stft(wave, plan, want_spec=True, want_phase=True, mask_frames=frm, mask_in_gap_is_one=True,
                   want_mask=True)
    imp = sp.stft(wave, plan, zero_frames=frm, mag_kind=sp.MAG_LOG10_EPS, eps=1e-9, want_spec=False)
    return {"original_spectrogram": full["spec"], "original_phase": full["phase"], "mask": full["mask"],
            "log_impaired_magnitude": imp["mag"], "gap_frames": frm}


def eval_gan_batch(wave: torch.Tensor, sample_rate: int = 16000, n_fft: int = 512, hop_length: int = 128,
                   win_length: int = 512, gap_len_s: float = 0.08, gap_start_s: float = 2.0) -> dict:
    """models/model_eval.py:61-111 for a batch."""
    return gan_batch(wave, gap_len_s, sample_rate, n_fft, hop_length, win_length, "hann", 1.0, True, gap_start_s)


def backend_batch(magnitude: torch.Tensor, phase: torch.Tensor, n_fft: int = 512, hop_length: int = 192,
                  win_length: int = 384, mag_domain: int = sp.DOM_LINEAR, db_auto: bool = True) -> torch.Tensor:
    """utils.spectrogram_to_audio(mag, phase=phase, ...) for a batch (models/model_eval.py:131-140).
    ``mag_domain=DOM_POW10`` fuses the caller's ``10 ** x`` (model_eval.py:163) into the kernel prologue."""
    plan = sp.get_plan(n_fft, hop_length, win_length, "hann", True, magnitude.device)
    return sp.istft(plan, mag=magnitude, phase=phase, mag_domain=mag_domain,
                    db_auto=db_auto and mag_domain == sp.DOM_LINEAR)


def cnnblstm_backend_batch(model_out: torch.Tensor, log_spectrogram_gap: torch.Tensor, gap_mask: torch.Tensor,
                           phase: torch.Tensor, n_fft: int = 512, hop_length: int = 192, win_length: int = 384,
                           save_pcm16: bool = False) -> torch.Tensor:
    """Everything between the network's raw output and the waveform in ONE kernel: reconstruct_spectrogram's blend
    ``out * mask + in * (1 - mask)`` (models/CNNBLSTM/model.py:108), ``10 **`` (models/model_eval.py:163) and
    ``utils.spectrogram_to_audio(mag, phase=phase)`` (:179-189).  ``save_pcm16`` goes on to what ``utils.save_audio`` puts
    into the inpainted FLAC (models/model_eval.py:190-192 -> utils.py:83-87): peak-normalised, 16-bit, returned as int16."""
    plan = sp.get_plan(n_fft, hop_length, win_length, "hann", True, model_out.device)
    return sp.istft_blend(plan, model_out, log_spectrogram_gap, gap_mask, phase, mag_domain=sp.DOM_POW10,
                          normalize=save_pcm16, pcm16=save_pcm16)


class HostPipeline:
    """Host-buffer entry point of the front-end: pinned host waveforms in, pinned host spectrograms out.

    The batch is cut into chunks; chunk i's H2D copy, kernel and D2H copy run on stream i % n_streams (4 by default) with
    per-stream device buffers, so copies of neighbouring chunks overlap the kernel (PCIe is the bound).  Every stream owns
    its own forward tile counter (aip_b200.h), so any number of chunks may be in flight."""

    def __init__(self, plan: sp.StftPlan, batch: int, n_samples: int, chunk: int = 64, n_streams: int = 4,
                 t_out: Optional[int] = None):
        self.plan, self.B, self.L = plan, int(batch), int(n_samples)
        self.chunk = max(1, min(int(chunk), self.B))
        self.T = plan.num_frames(self.L)
        self.t_out = self.T if t_out is None else min(int(t_out), self.T)
        dev = plan.device
        self.streams = [torch.cuda.Stream(device=dev) for _ in range(n_streams)]
        self.d_wave = [torch.empty((self.chunk, self.L), dtype=torch.float32, device=dev) for _ in range(n_streams)]
        self.d_gaps = [torch.empty((self.chunk, 2), dtype=torch.int32, device=dev) for _ in range(n_streams)]
        self.d_mag = [torch.empty((self.chunk, plan.n_bins, self.t_out), dtype=torch.float32, device=dev)
                      for _ in range(n_streams)]
        self.h_gaps = torch.empty((self.B, 2), dtype=torch.int32, pin_memory=True)

    def logmag_gap(self, h_wave: torch.Tensor, gap_samples: np.ndarray, h_out: torch.Tensor, eps: float = 1e-9,
                   mag_kind: int = sp.MAG_LOG10_EPS, copies_only: bool = False) -> torch.Tensor:
        """h_out[b] = log10(|stft(h_wave[b] with samples [g0,g1) zeroed)| + eps); returns h_out after a sync.
        ``copies_only`` skips the kernel and moves the same bytes: the host / PCIe ceiling of this call (bench.py)."""
        if h_wave.is_cuda or h_out.is_cuda:
            raise ValueError("HostPipeline takes HOST tensors (pinned for speed)")
        if tuple(h_wave.shape) != (self.B, self.L) or tuple(h_out.shape) != (self.B, self.plan.n_bins, self.t_out):
            raise ValueError("shape mismatch with the pipeline's batch geometry")
        self.h_gaps.copy_(torch.from_numpy(np.ascontiguousarray(gap_samples, dtype=np.int32)))
        cur = torch.cuda.current_stream(self.plan.device)
        for s in self.streams:
            s.wait_stream(cur)
        for i, lo in enumerate(range(0, self.B, self.chunk)):
            hi = min(self.B, lo + self.chunk)
            n = hi - lo
            k = i % len(self.streams)
            with torch.cuda.stream(self.streams[k]):
                self.d_wave[k][:n].copy_(h_wave[lo:hi], non_blocking=True)
                self.d_gaps[k][:n].copy_(self.h_gaps[lo:hi], non_blocking=True)
                if not copies_only:
                    sp.stft(self.d_wave[k][:n], self.plan, gap_samples=self.d_gaps[k][:n], mag_kind=mag_kind, eps=eps,
                            t_out=self.t_out, want_spec=False, out={"mag": self.d_mag[k][:n]})
                h_out[lo:hi].copy_(self.d_mag[k][:n], non_blocking=True)
        for s in self.streams:
            cur.wait_stream(s)
        cur.synchronize()
        return h_out
