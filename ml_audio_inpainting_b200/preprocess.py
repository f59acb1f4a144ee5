"""Bulk preprocessing (reference pre_process_dataset.py:19-43, BASELINE.json configs[3]).

The reference walks the LibriSpeech tree and, per file: decode -> pad / truncate to 5 s -> zero a random
0.1 s range (utils.add_random_gap, exclusive upper bound) -> peak-normalise -> encode.  Here decode / encode
stay host work (codec is out of the GPU path), and the numeric part runs per batch of clips on the device:

  preprocess_batch   [N, L] waveforms -> gapped (+ peak-normalised) waveforms and, optionally, the
                     log-magnitude spectrogram of the gapped audio (the "spectrogram + gap" shape of configs[3]);
                     shards by clip over torch.distributed ranks with no collective.
  preprocess_tree    the drop-in loop over a directory tree.
"""
from __future__ import annotations

import ctypes as C
import os
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path
from typing import Optional

import numpy as np
import torch

from . import _cabi, audio_io, gaps, sharding, spectral as sp

__all__ = ["preprocess_batch", "preprocess_tree"]


def preprocess_batch(wave: torch.Tensor, gap_len: float = 0.1, sample_rate: int = 16000,
                     starts: Optional[np.ndarray] = None, normalize: bool = True, want_logmag: bool = False,
                     n_fft: int = 512, hop_length: int = 192, win_length: int = 384, want_pcm16: bool = False) -> dict:
    """``wave`` [N, L] float32 on the device (this rank's shard).  One np.random draw per clip unless ``starts``
    is given (the caller draws for the WHOLE corpus and passes its slice: see ``sharding``).  ``want_pcm16`` replaces the
    normalised float waveform by ``pcm16`` [N, L] int16: the samples ``save_audio`` writes into its 16-bit FLAC (utils.py:83-87),
    straight from the gapped waveform (half the device -> host bytes, no host-side quantisation)."""
    if not wave.is_cuda:
        raise RuntimeError("wave must be a CUDA tensor: there is no CPU path")
    N, L = wave.shape
    g = gaps.gap_len_samples(gap_len, sample_rate)
    if g >= L:
        raise ValueError(f"Gap length ({g}s) exceeds audio length ({L / sample_rate}s)")
    if starts is None:
        starts = gaps.draw_starts_exclusive(L, g, N)
    starts = np.asarray(starts, dtype=np.int64)
    sam = np.stack([starts, starts + g], 1)
    sam_d = torch.as_tensor(sam.astype(np.int32), device=wave.device)
    lib = _cabi.load()
    st = C.c_void_p(torch.cuda.current_stream(wave.device).cuda_stream)      # the stream of the tensors' device
    out = torch.empty_like(wave)
    with torch.cuda.device(wave.device):
        _cabi.check(lib.aip_gap_zero_f32(wave.data_ptr(), wave.stride(0), out.data_ptr(), out.stride(0), N, L,
                                         sam_d.data_ptr(), st), "aip_gap_zero_f32")
        res = {"audio_gap": out, "gap_samples": sam,
               "gap_int_s": np.stack(gaps.seconds_interval(starts, g, sample_rate), 1)}
        if want_pcm16:
            peaks = torch.empty(N, dtype=torch.float32, device=wave.device) if normalize else None
            res["pcm16"] = sp.wave_to_pcm16(out, normalize=normalize, peaks_out=peaks)
            if normalize:
                res["peaks"] = peaks
        elif normalize:
            peaks = torch.empty(N, dtype=torch.float32, device=wave.device)
            norm = torch.empty_like(wave)
            for lo in range(0, N, 65535):
                hi = min(N, lo + 65535)
                _cabi.check(lib.aip_peak_normalize_f32(out[lo:hi].data_ptr(), out.stride(0), norm[lo:hi].data_ptr(),
                                                       norm.stride(0), hi - lo, L, peaks[lo:hi].data_ptr(), st),
                            "aip_peak_normalize_f32")
            res["audio_gap_normalized"] = norm
            res["peaks"] = peaks
    if want_logmag:
        plan = sp.get_plan(n_fft, hop_length, win_length, "hann", True, wave.device)
        res["logmag_gap"] = sp.stft(wave, plan, gap_samples=sam_d, mag_kind=sp.MAG_LOG10_EPS, eps=1e-9,
                                    want_spec=False)["mag"]
    return res


DECODABLE = (".flac", ".wav")        # containers the built-in reader handles (audio_io); the reference lists ".mp3" too


def preprocess_tree(src_root, dst_root, gap_len: float = 0.1, sample_rate: int = 16000, max_len: float = 5,
                    supported_formats=(".flac", ".wav"), batch: int = 256, device=None, progress: bool = True,
                    io_threads: Optional[int] = None):
    """The reference's loop (pre_process_dataset.py:19-43) with the device doing the arithmetic: files are visited in
    os.walk order, one np.random draw per file in that order, results written under ``dst_root`` mirroring the tree.

    Like ``utils.save_audio`` (utils.py:54-89, ``file_format='flac'``) every output holds FLAC data whatever its suffix.
    A file whose suffix is in ``supported_formats`` but which the built-in reader cannot decode (".mp3": the reference
    decodes it through librosa / audioread) is skipped with a warning and consumes no random draw."""
    src_root, dst_root = Path(src_root), Path(dst_root)
    device = torch.device(device or f"cuda:{torch.cuda.current_device()}")
    jobs = []
    for root, subdirs, files in os.walk(src_root, topdown=True):
        rel = os.path.relpath(root, src_root)
        dest = dst_root / rel
        os.makedirs(dest, exist_ok=True)
        if len(subdirs) == 0:
            for f in files:
                if Path(f).suffix in supported_formats:
                    if Path(f).suffix.lower() not in DECODABLE:
                        print(f"Warning: {Path(root) / f}: no built-in decoder for this container, skipped")
                        continue
                    jobs.append((Path(root) / f, dest / f))
    L = int(sample_rate * max_len)
    g = gaps.gap_len_samples(gap_len, sample_rate)
    starts_all = gaps.draw_starts_exclusive(L, g, len(jobs))       # same stream as len(jobs) scalar draws
    rank, ws, _ = sharding.world()
    lo, hi = sharding.shard_bounds(len(jobs), rank, ws)
    it = range(lo, hi, batch)
    if progress:
        try:
            from tqdm import tqdm
            it = tqdm(it, desc="Pre-Processing Dataset")
        except ImportError:
            pass
    # decode and encode are host work in the native codec (csrc/aip_flac.c), which runs outside the GIL: one file per thread
    pool = ThreadPoolExecutor(max_workers=io_threads or min(32, os.cpu_count() or 1))

    def decode_into(args):
        host, i, src = args
        pcm, sr = audio_io.read_audio(src, max_samples=L)
        x = pcm.mean(axis=1, dtype=np.float32) if pcm.ndim == 2 else pcm
        if sr != sample_rate:
            raise IOError(f"{src}: sample rate {sr} != {sample_rate} (resampling is not part of the bulk path)")
        host[i, : min(L, len(x))] = x[:L]

    try:
        for b0 in it:
            b1 = min(hi, b0 + batch)
            host = np.zeros((b1 - b0, L), dtype=np.float32)
            list(pool.map(decode_into, [(host, i, src) for i, (src, _) in enumerate(jobs[b0:b1])]))
            res = preprocess_batch(torch.from_numpy(host).to(device), gap_len, sample_rate, starts=starts_all[b0:b1],
                                   want_pcm16=True)
            out = res["pcm16"].cpu().numpy()                           # the FLAC's 16-bit samples, quantised on the device
            # utils.save_audio's default format (utils.py:59)
            list(pool.map(lambda a: audio_io.write_audio(a[0][1], a[1], sample_rate, "flac"), zip(jobs[b0:b1], out)))
    finally:
        pool.shutdown()
    return len(jobs)
