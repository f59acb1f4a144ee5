#!/usr/bin/env python
"""Benchmark of the spectrogram hot path (BASELINE.json metric: audio-seconds per second).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--clips B]

A *step* is one pass of the forward hot path over one batch of synthetic 16 kHz audio:
BASELINE.json configs[1] -- "batched STFT log-magnitude + gap masking, 4096 synthetic 16 kHz 10 s
clips" at the reference's parameters (n_fft 512 / win 384 / hop 192, gap 0.2 s, log10(|S|+1e-9)).
One step = ONE launch of ``stft512_fwd_kernel`` per rank.  The JSON line also carries the inverse
leg (configs[2]: iSTFT overlap-add, batch 1024) and the round trip under ``"legs"``.

  value     whole-job audio-s/s, inputs resident in HBM, CUDA events, max over ranks
  e2e       same metric through ``frontend.HostPipeline`` (pinned HOST buffers in, HOST buffers
            out; H2D + kernel + D2H inside the timed region, 64-clip chunks on 4 streams), and the same chunked copies
            WITHOUT the kernel beside it (``copy_ceiling``: what the host / PCIe side allows on this box at this N)
  roofline  algorithmic bytes (4 L + 4 F T per clip) / kernel time vs MEASURED_PEAKS.json
  cpu_baseline  the numpy/scipy oracle port of the same step on the host cores (bounded sample; clips generated BEFORE the
            clock starts), with the one-thread figure and torch.stft on the CPU (a stronger CPU implementation) beside it
  istft / roundtrip / cnnblstm_e2e   top-level copies of the legs the metric's name also covers (configs[2], the STFT -> mask ->
            iSTFT round trip, and configs[4]: front-end -> the reference's CNN-BLSTM architecture -> back-end)

``--impl reference`` times only that CPU port (the reference's own librosa path cannot be
installed here: librosa/soundfile are absent from the image and the wheelhouse).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

SR, N_FFT, WIN, HOP = 16000, 512, 384, 192
CLIP_S = 10.0
GAP_S = 0.2
EPS = 1e-9
METRIC = "audio-seconds/sec (STFT->mask->iSTFT)"
_RESULT_OUT = sys.stdout


# ----------------------------------------------------------------------------------------------- CPU
_CPU_DISTINCT = 32      # distinct synthetic clips per worker (20 MB); the timed loop cycles over them


def _cpu_inputs(seed, L):
    """The worker's synthetic clips and gap starts -- generated BEFORE the clock starts, like the GPU arm's."""
    rng = np.random.default_rng(seed)
    g = int(GAP_S * SR)
    clips = [np.clip(0.1 * rng.standard_normal(L), -1, 1).astype(np.float32) for _ in range(_CPU_DISTINCT)]
    starts = [int(rng.integers(0, L - g)) for _ in range(_CPU_DISTINCT)]
    return clips, starts, g


def _cpu_worker(args):
    """Oracle port of one step on `n` clips: gap zeroing -> librosa.stft -> log10(|S| + 1e-9).  Timed: that work only."""
    seed, n, L = args
    from oracle import librosa_port as lr          # bench.py's cpu_baseline / reference arm only
    clips, starts, g = _cpu_inputs(seed, L)
    acc = 0.0
    t0 = time.perf_counter()
    for i in range(n):
        x = clips[i % _CPU_DISTINCT].copy()        # the step's own gapped copy (utils.add_random_gap builds a new array too)
        s = starts[i % _CPU_DISTINCT]
        x[s:s + g] = 0.0
        S = lr.stft(x, n_fft=N_FFT, hop_length=HOP, win_length=WIN)
        m = np.log10(np.abs(S) + EPS).astype(np.float32)
        acc += float(m[0, 0])
    return time.perf_counter() - t0, acc


def _torch_cpu_worker(args):
    """The same step with torch.stft on the CPU (not the reference's implementation: a stronger CPU baseline)."""
    seed, n, L, threads = args
    import torch
    torch.set_num_threads(threads)
    clips, starts, g = _cpu_inputs(seed, L)
    from oracle import librosa_port as lr
    w = torch.from_numpy(lr.fft_window("hann", WIN, N_FFT).astype(np.float32))
    xs = [torch.from_numpy(c) for c in clips]
    acc = 0.0
    t0 = time.perf_counter()
    for i in range(n):
        x = xs[i % _CPU_DISTINCT].clone()
        s = starts[i % _CPU_DISTINCT]
        x[s:s + g] = 0.0
        S = torch.stft(x, N_FFT, HOP, N_FFT, window=w, center=True, pad_mode="constant", return_complex=True)
        m = torch.log10(S.abs() + EPS)
        acc += float(m[0, 0])
    return time.perf_counter() - t0, acc


def cpu_pass(pool, workers: int, clips_per_worker: int, L: int, seed: int = 0, fn=None, extra=()):
    """Returns (audio-s/s, seconds) of the CPU step over workers*clips_per_worker clips.  The workers run concurrently and each
    times its own loop (input generation excluded); the pass takes as long as the slowest of them."""
    fn = fn or _cpu_worker
    jobs = [(seed * 1000 + w, clips_per_worker, L, *extra) for w in range(workers)]
    res = [fn(j) for j in jobs] if pool is None else pool.map(fn, jobs)
    dt = max(r[0] for r in res)
    return workers * clips_per_worker * (L / SR) / dt, dt


def make_pool(workers: int):
    if workers <= 1:
        return None
    import multiprocessing as mp
    return mp.get_context("fork").Pool(workers)     # forked BEFORE CUDA is initialised


# -------------------------------------------------------------------------------------------- clocks
class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled while the timed region runs."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                 "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            p = [q.strip() for q in ln.split(",")]
            if len(p) < 7:
                continue
            try:
                sm.append(float(p[0])); mx.append(float(p[1]))
            except ValueError:
                continue
            for name, v in zip(names, p[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------------- helpers
def measured_peak():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        try:
            return float(json.loads(p.read_text())["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def recorded_traffic(kernel: str):
    """dram bytes per launch from the committed ncu --set full capture (profiles/traffic.json), or None."""
    p = ROOT / "profiles" / "traffic.json"
    if p.exists():
        try:
            return json.loads(p.read_text()).get(kernel)
        except Exception:
            return None
    return None


def reference_arm(args, rank: int):
    """The CPU implementation of the same step, all host threads, bounded sample per step."""
    if rank != 0:
        return
    L = int(CLIP_S * SR)
    workers = os.cpu_count() or 1
    pool = make_pool(workers)
    per = max(1, args.ref_clips // workers)
    for _ in range(args.warmup):
        cpu_pass(pool, workers, max(1, per // 4), L, seed=99)
    t = []
    for k in range(args.steps):
        _, dt = cpu_pass(pool, workers, per, L, seed=k)
        t.append(dt)
    if pool is not None:
        pool.close()
    ms = 1e3 * float(np.mean(t))
    value = workers * per * CLIP_S / (ms * 1e-3)
    sample = (f"each step = a bounded sample of the configured workload: {workers * per} of its {args.clips} clips x {CLIP_S:g} s "
              f"({workers} processes x {per} clips, inputs generated before the clock starts), oracle port (numpy + scipy.fft pocketfft = "
              "what librosa runs); librosa itself is not installable here")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "audio-s/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": make_config(args.clips, max(1, args.gpus)),
        "cpu_baseline": {"value": value, "unit": "audio-s/s", "cores": workers, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "audio-s/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), file=_RESULT_OUT, flush=True)


WORKLOAD = ("configs[1]: batched STFT log-magnitude + gap masking, 4096 synthetic 16 kHz 10 s clips per GPU "
            "(n_fft 512 / win 384 / hop 192, gap 0.2 s, log10(|S|+1e-9))")


def make_config(clips_per_gpu: int, world: int) -> dict:
    """One config for both arms (ours and --impl reference): the workload both are measured on."""
    L = int(CLIP_S * SR)
    return {"workload": WORKLOAD, "clips_per_gpu": clips_per_gpu, "n_fft": N_FFT, "win_length": WIN, "hop_length": HOP,
            "sample_rate": SR, "clip_seconds": CLIP_S, "gap_seconds": GAP_S, "frames": 1 + L // HOP,
            "outputs": "log10(|S|+1e-9) f32 [B,257,T]", "sharding": f"by clip, {world} rank(s), no collective",
            "l2": "inputs (2.6 GB) and outputs (3.5 GB) per step exceed the 126 MB L2; no flush needed"}


def pin_to_gpu_cpus(index: int):
    """Bind this rank to the CPUs NVML reports as local to its GPU (pinned buffers are then first-touched from there)."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(index)
        n = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (n + 63) // 64)
        cpus = {64 * i + b for i, w in enumerate(words) for b in range(64) if (w >> b) & 1 and 64 * i + b < n}
        numa = None
        try:
            numa = int(pynvml.nvmlDeviceGetNumaNodeId(h))
        except Exception:
            pass
        if cpus:
            os.sched_setaffinity(0, cpus)
        return {"cpus": len(cpus), "first": min(cpus) if cpus else None, "last": max(cpus) if cpus else None, "numa_node": numa}
    except Exception as e:          # affinity is an optimisation, never a requirement
        return {"error": str(e)[:80]}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--clips", type=int, default=4096, help="clips per GPU (configs[1]: 4096)")
    ap.add_argument("--inv-clips", type=int, default=1024, help="clips per GPU of the iSTFT leg (configs[2]: 1024)")
    ap.add_argument("--ref-clips", type=int, default=1024, help="clips per step of the CPU arm")
    ap.add_argument("--cpu-clips", type=int, default=2048, help="clips of the cpu_baseline sample (rank 0, N=1)")
    ap.add_argument("--gl-clips", type=int, default=1024, help="clips per GPU of the Griffin-Lim leg (configs[2], 32 iterations)")
    ap.add_argument("--model-clips", type=int, default=32, help="clips per GPU of the configs[4] leg (batch 256 on 8 GPUs = 32 per GPU)")
    ap.add_argument("--no-legs", action="store_true", help="headline step only (profiling runs)")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-clock-probe", action="store_true", help="skip the 1.5 s untimed continuation (ncu launch lists)")
    ap.add_argument("--no-e2e", action="store_true")
    args = ap.parse_args()

    # The contract is ONE JSON line on stdout.  Libraries write there too (NCCL prints its version banner on stdout
    # when NCCL_DEBUG=VERSION is set in the environment): keep a private handle on the real stdout for the result line
    # and send everything else to stderr.
    global _RESULT_OUT
    sys.stdout.flush()
    _RESULT_OUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        reference_arm(args, rank)
        return

    L = int(CLIP_S * SR)
    # ---- CPU baseline first: the worker pool is forked before this process touches CUDA
    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu:
        workers = os.cpu_count() or 1
        pool = make_pool(workers)
        per = max(1, args.cpu_clips // workers)
        cpu_pass(pool, workers, 1, L, seed=7)                      # warm the workers
        v, dt = cpu_pass(pool, workers, per, L, seed=1)
        v1, dt1 = cpu_pass(None, 1, 96, L, seed=2)                  # one process, one thread
        tv1, tdt1 = cpu_pass(None, 1, 256, L, seed=3, fn=_torch_cpu_worker, extra=(1,))
        tvn, tdtn = cpu_pass(pool, workers, per, L, seed=4, fn=_torch_cpu_worker, extra=(1,))
        if pool is not None:
            pool.close()
        cpu_baseline = {"value": v, "unit": "audio-s/s", "cores": workers, "kind": "port",
                        "sample": f"{workers * per} clips x {CLIP_S:g} s of the same step ({dt:.1f} s, inputs generated before the clock "
                                  "starts), oracle port (numpy + scipy.fft pocketfft = what librosa runs), one process per core",
                        "single_thread": {"value": v1, "unit": "audio-s/s", "cores": 1, "sample": f"96 clips ({dt1:.1f} s)"},
                        "torch_stft": {"note": "torch.stft + abs + log10 on the CPU: NOT the reference's implementation, a stronger CPU "
                                               "baseline (SURVEY 8d); fp32",
                                       "single_thread": {"value": tv1, "unit": "audio-s/s", "cores": 1, "sample": f"256 clips ({tdt1:.1f} s)"},
                                       "all_cores": {"value": tvn, "unit": "audio-s/s", "cores": workers,
                                                     "sample": f"{workers * per} clips, one single-threaded process per core ({tdtn:.1f} s)"}}}

    import torch
    import torch.distributed as dist
    from ml_audio_inpainting_b200 import frontend, gaps, spectral as sp

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
    affinity = pin_to_gpu_cpus(local_rank)
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    B = args.clips
    plan = sp.get_plan(N_FFT, HOP, WIN, "hann", True, dev)
    T = plan.num_frames(L)
    F = plan.n_bins
    gen = torch.Generator(device=dev).manual_seed(1234 + rank)
    wave = (0.1 * torch.randn((B, L), generator=gen, device=dev, dtype=torch.float32)).clamp_(-1, 1)
    g = gaps.gap_len_samples(GAP_S, SR)
    np.random.seed(rank)
    starts = gaps.draw_starts_exclusive(L, g, B)                     # utils.add_random_gap's draw
    gap_dev = torch.as_tensor(np.stack([starts, starts + g], 1).astype(np.int32), device=dev)
    out = {"mag": torch.empty((B, F, T), dtype=torch.float32, device=dev)}

    def step():
        sp.stft(wave, plan, gap_samples=gap_dev, mag_kind=sp.MAG_LOG10_EPS, eps=EPS, want_spec=False, out=out)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, warmup):
        for _ in range(warmup):
            fn()
        barrier()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(steps + 1)]
        ev[0].record()
        for k in range(steps):
            fn()
            ev[k + 1].record()
        barrier()
        per = [ev[k].elapsed_time(ev[k + 1]) for k in range(steps)]
        total = ev[0].elapsed_time(ev[steps])
        if world > 1:
            t = torch.tensor([total], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            total = float(t.item())
        return total / steps, per

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ms, per = timed(step, args.steps, args.warmup)
    # the timed region lasts only K x 1.8 ms, shorter than nvidia-smi's sampling period: keep the SAME step
    # running (untimed) for ~1.5 s so that the clock / throttle record is taken under this kernel's load
    t_probe = time.perf_counter()
    while time.perf_counter() - t_probe < (0.0 if args.no_clock_probe else 1.5):
        for _ in range(50):
            step()
        torch.cuda.synchronize()
    clocks = sampler.stop() if rank == 0 else None
    if clocks is not None:
        clocks["note"] = "sampled every 100 ms over the timed region plus a 1.5 s untimed continuation of the same step"
    value = world * B * CLIP_S / (ms * 1e-3)

    # ---- roofline of the dominant kernel (one launch per step => kernel time = step time on the stream)
    peak, peak_src = measured_peak()
    alg_bytes = B * (4 * L + 4 * F * T)
    kernel_ms = float(np.mean(per))
    achieved = alg_bytes / (kernel_ms * 1e-3) / 1e9
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": recorded_traffic("stft512_fwd_kernel"), "kernel": "stft512_fwd_kernel<LOG10_EPS>",
                "algorithmic_bytes_per_launch": alg_bytes, "kernel_ms": kernel_ms, "peak_source": peak_src,
                "frac_of_8TBps_nominal": achieved / 8000.0}

    # ---- inverse leg (configs[2]) and round trip, reported beside the headline
    legs = {}
    Bi = min(args.inv_clips, B)
    if args.no_legs:
        Bi = 0
    if Bi > 0:
        legs = extra_legs(args, sp, plan, wave, gap_dev, Bi, T, F, L, world, peak, timed)

    # ---- end to end through the host-buffer API
    e2e = None
    if not args.no_e2e:
        e2e = e2e_leg(args, frontend, plan, wave, starts, g, B, L, F, T, world, dev, barrier)
        e2e["cpu_affinity"] = affinity
    # ---- configs[4]: waveform (pinned host) -> front-end -> CNN-BLSTM -> back-end -> waveform (pinned host)
    if not args.no_legs and args.model_clips > 0:
        del wave, out
        torch.cuda.empty_cache()
        legs["cnnblstm_e2e"] = cnnblstm_leg(args, frontend, sp, world, dev, barrier)

    if rank == 0:
        emit_line(args, world, value, ms, B, T, roofline, cpu_baseline, e2e, clocks, legs)
    if world > 1:
        dist.destroy_process_group()


def extra_legs(args, sp, plan, wave, gap_dev, Bi, T, F, L, world, peak, timed):
    import torch
    legs = {}
    dev = wave.device
    spec = sp.stft(wave[:Bi], plan, gap_samples=gap_dev[:Bi])["spec"]
    yout = torch.empty((Bi, plan.istft_length(T)), dtype=torch.float32, device=dev)
    ms_i, per_i = timed(lambda: sp.istft(plan, spec=spec, out=yout), args.steps, args.warmup)
    inv_bytes = Bi * (8 * F * T + 4 * yout.shape[1])
    ach = inv_bytes / (np.mean(per_i) * 1e-3) / 1e9
    legs["istft"] = {"workload": f"configs[2]: iSTFT overlap-add of masked complex spectrograms, batch {Bi} per GPU",
                     "value": world * Bi * CLIP_S / (ms_i * 1e-3), "unit": "audio-s/s", "ms_per_step": ms_i,
                     "roofline": {"bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                                  "traffic": recorded_traffic("istft512_kernel"), "kernel": "istft512_kernel<INV_SPEC>",
                                  "algorithmic_bytes_per_launch": inv_bytes}}
    sout = {"spec": spec}

    def roundtrip():
        sp.stft(wave[:Bi], plan, gap_samples=gap_dev[:Bi], out=sout)
        sp.istft(plan, spec=spec, out=yout)

    ms_r, _ = timed(roundtrip, args.steps, args.warmup)
    rt_bytes = Bi * (4 * L + 8 * F * T) + inv_bytes
    legs["roundtrip"] = {"workload": f"STFT(complex)+gap -> iSTFT, batch {Bi} per GPU (2 launches per step)",
                         "value": world * Bi * CLIP_S / (ms_r * 1e-3), "unit": "audio-s/s", "ms_per_step": ms_r,
                         "hbm_frac": rt_bytes / (ms_r * 1e-3) / 1e9 / peak}
    # P2 (GAN parameters: win 512 / hop 128) forward, same step shape
    plan2 = sp.get_plan(N_FFT, 128, 512, "hann", True, dev)
    T2 = plan2.num_frames(L)
    out2 = {"mag": torch.empty((Bi, F, T2), dtype=torch.float32, device=dev)}
    ms_2, per_2 = timed(lambda: sp.stft(wave[:Bi], plan2, gap_samples=gap_dev[:Bi], mag_kind=sp.MAG_LOG10_EPS, eps=EPS,
                                        want_spec=False, out=out2), args.steps, args.warmup)
    b2 = Bi * (4 * L + 4 * F * T2)
    legs["stft_p2"] = {"workload": f"forward log-magnitude + gap at the GAN parameters (win 512 / hop 128), batch {Bi} per GPU",
                       "value": world * Bi * CLIP_S / (ms_2 * 1e-3), "unit": "audio-s/s", "ms_per_step": ms_2,
                       "hbm_frac": b2 / (np.mean(per_2) * 1e-3) / 1e9 / peak}
    del out2, spec, sout
    # the reference's function defaults (utils.extract_spectrogram: n_fft 2048 / hop 512, utils.py:192-193): the tiled radix-16
    # kernels of csrc/aip_pow2.cu, complex spectrogram out and back in
    Bd = min(512, Bi)
    plan_d = sp.get_plan(2048, 512, 2048, "hann", True, dev)
    Td, Fd = plan_d.num_frames(L), 1025
    dout = {"spec": torch.empty((Bd, Fd, Td), dtype=torch.complex64, device=dev)}
    ms_df, _ = timed(lambda: sp.stft(wave[:Bd], plan_d, out=dout), args.steps, args.warmup)
    yd = torch.empty((Bd, plan_d.istft_length(Td)), dtype=torch.float32, device=dev)
    ms_di, _ = timed(lambda: sp.istft(plan_d, spec=dout["spec"], out=yd), args.steps, args.warmup)
    if world > 1:
        # the OPTIONAL final gather: every rank's output shard to every rank over NVLink / NVSwitch.  Not part of the
        # hot path (no collective there) and not in any other number: 512 clips' log-magnitudes per rank (the shard of a
        # 4096-clip job on 8 GPUs, 439 MB) through torch.distributed.all_gather_into_tensor (NCCL).
        import torch.distributed as dist
        Bg = min(512, Bi)
        shard = sp.stft(wave[:Bg], plan, gap_samples=gap_dev[:Bg], mag_kind=sp.MAG_LOG10_EPS, eps=EPS, want_spec=False)["mag"]
        full = torch.empty((world * Bg, F, T), dtype=torch.float32, device=dev)
        ms_g, _ = timed(lambda: dist.all_gather_into_tensor(full, shard), max(3, args.steps // 4), 2)
        nbytes = shard.numel() * 4
        legs["gather_outputs"] = {"workload": f"optional all-gather of {Bg} clips' log-magnitudes per rank to every rank (NCCL)",
                                  "ms_per_step": ms_g, "bytes_per_rank": nbytes,
                                  "received_GBps_per_gpu": (world - 1) * nbytes / (ms_g * 1e-3) / 1e9}
        del shard, full
    legs["default_params_2048"] = {
        "workload": f"utils.extract_spectrogram / spectrogram_to_audio at the reference's default n_fft 2048 / hop 512, batch {Bd} per GPU",
        "stft": {"value": world * Bd * CLIP_S / (ms_df * 1e-3), "unit": "audio-s/s", "ms_per_step": ms_df,
                 "hbm_frac": Bd * (4 * L + 8 * Fd * Td) / (ms_df * 1e-3) / 1e9 / peak},
        "istft": {"value": world * Bd * CLIP_S / (ms_di * 1e-3), "unit": "audio-s/s", "ms_per_step": ms_di,
                  "hbm_frac": Bd * (8 * Fd * Td + 4 * yd.shape[1]) / (ms_di * 1e-3) / 1e9 / peak}}
    del dout, yd
    # dataset-shaped collate (models/CNNBLSTM/dataset.py:74-121): 256 files x 25 gaps x 5 s, log-magnitudes of
    # all variants from ONE clean transform per file (aip_stft_gap_variants_f32) against 25 full gapped transforms
    Nf, G, L5 = min(256, Bi), 25, 80000
    T5 = plan.num_frames(L5)
    w5 = wave[:Nf, :L5].contiguous()
    g5 = int(0.2 * SR)
    rs = np.random.RandomState(7)
    st5 = rs.randint(0, L5 - g5, size=Nf * G)
    gaps5 = torch.as_tensor(np.stack([st5, st5 + g5], 1).astype(np.int32), device=dev)
    vout = torch.empty((Nf * G, F, T5), dtype=torch.float32, device=dev)
    ms_v, _ = timed(lambda: sp.stft_gap_variants(w5, plan, gaps5, G, mag_kind=sp.MAG_LOG10_EPS, eps=EPS, t_out=T5, out=vout, gap_len_max=g5),
                    args.steps, args.warmup)
    w5r = w5.repeat_interleave(G, 0)
    fout = {"mag": torch.empty((Nf * G, F, T5), dtype=torch.float32, device=dev)}
    ms_f, _ = timed(lambda: sp.stft(w5r, plan, gap_samples=gaps5, mag_kind=sp.MAG_LOG10_EPS, eps=EPS, t_out=T5, want_spec=False,
                                    out=fout), args.steps, args.warmup)
    same = bool(torch.equal(vout, fout["mag"]))
    vb = Nf * (4 * L5 + 4 * F * T5 * G)          # compulsory: every file read once, every variant written once
    legs["dataset_variants"] = {"workload": f"CNNBLSTM dataset item shape: {Nf} files x {G} gaps x 5 s per GPU, log10 magnitudes of every "
                                            "variant (clean transform + TMA copy pass + re-transform of the frames each gap touches: "
                                            "3 launches per step)",
                                "value": world * Nf * G * 5.0 / (ms_v * 1e-3), "unit": "gapped audio-s/s", "ms_per_step": ms_v,
                                "hbm_frac": vb / (ms_v * 1e-3) / 1e9 / peak, "compulsory_bytes_per_step": vb,
                                "ms_per_step_full_transforms": ms_f, "speedup_vs_full_transforms": ms_f / ms_v,
                                "bit_identical_to_full_transforms": same}
    del vout, fout, w5r, w5
    # configs[3]: pre_process_dataset.py-shaped 100 h corpus (72 000 clips x 5 s, gap 0.1 s = pre_process_dataset.py:38), spectrogram +
    # gap preprocessing sharded by clip: every rank transforms its 72 000 / world clips, device-resident, in sub-batches of <= 9 000
    # clips (the 8-GPU shard) that reuse one input and one output buffer (each larger than L2); no collective
    n_total, L5, sub = 72000, 80000, min(9000, 9 * Bi)
    shard = (n_total + world - 1) // world
    n_sub = (shard + sub - 1) // sub
    wsub = (0.1 * torch.randn(sub, L5, device=dev)).clamp_(-1, 1)
    g01 = int(0.1 * SR)
    st01 = np.random.RandomState(11).randint(0, L5 - g01, size=sub)
    gsub = torch.as_tensor(np.stack([st01, st01 + g01], 1).astype(np.int32), device=dev)
    T5 = plan.num_frames(L5)
    osub = {"mag": torch.empty((sub, F, T5), dtype=torch.float32, device=dev)}

    def shard_pass():
        done = 0
        for _ in range(n_sub):
            n = min(sub, shard - done)
            sp.stft(wsub[:n], plan, gap_samples=gsub[:n], mag_kind=sp.MAG_LOG10_EPS, eps=EPS, want_spec=False,
                    out={"mag": osub["mag"][:n]})
            done += n

    ms_p, _ = timed(shard_pass, max(3, args.steps // 4), 2)
    pb = shard * (4 * L5 + 4 * F * T5)
    legs["preprocess_100h"] = {"workload": f"configs[3]: 100 h corpus = {n_total} clips x 5 s, gap 0.1 s, log10 magnitudes; {shard} clips per GPU in "
                                           f"{n_sub} device-resident launches of <= {sub} clips, sharded by clip, no collective",
                               "value": n_total * 5.0 / (ms_p * 1e-3), "unit": "audio-s/s", "ms_per_step": ms_p,
                               "hbm_frac": pb / (ms_p * 1e-3) / 1e9 / peak, "algorithmic_bytes_per_gpu": pb,
                               "corpus_hours_per_second": n_total * 5.0 / 3600.0 / (ms_p * 1e-3)}
    del wsub, osub
    # Griffin-Lim, 32 iterations (configs[2] "+ Griffin-Lim 32 iters"): streaming bound 8 995 656 B / clip / iteration
    Bg = min(args.gl_clips, Bi)
    if Bg > 0:
        mag = sp.stft(wave[:Bg], plan, mag_kind=sp.MAG_ABS, want_spec=False)["mag"]
        gen = torch.Generator(device=dev).manual_seed(99)
        n_iter = 32
        ms_g, _ = timed(lambda: sp.griffinlim(plan, mag, n_iter=n_iter, generator=gen), max(2, args.steps // 5), 1)
        # bytes one iteration STREAMS with the phase update fused into the inverse kernel's load: inverse reads rebuilt[it],
        # rebuilt[it - 1] (8 F T each) and |S| (4 F T) and writes the waveform; forward reads it and writes rebuilt[it + 1]
        per_iter = 3 * 8 * F * T + 4 * F * T + 2 * 4 * HOP * (T - 1)
        streamed = Bg * (n_iter * per_iter + (8 * F * T + 4 * HOP * (T - 1)))
        gl_bytes = Bg * (n_iter * 8995656 + 2354448)             # SURVEY 8d's bound with a SEPARATE update pass, for reference
        legs["griffinlim32"] = {"workload": f"Griffin-Lim 32 iterations (momentum 0.99, random init drawn on device), batch {Bg} per GPU, "
                                            f"{2 * n_iter + 2} kernel launches per step (inverse with the phase update fused into its load and "
                                            "both complex arrays staged by TMA, forward with complex output)",
                                "value": world * Bg * CLIP_S / (ms_g * 1e-3), "unit": "audio-s/s", "ms_per_step": ms_g,
                                "streamed_bytes_per_step": streamed, "hbm_frac": streamed / (ms_g * 1e-3) / 1e9 / peak,
                                "survey_streaming_bound_bytes_per_step": gl_bytes,
                                "hbm_frac_of_survey_streaming_bound": gl_bytes / (ms_g * 1e-3) / 1e9 / peak}
    return legs


def e2e_leg(args, frontend, plan, wave, starts, g, B, L, F, T, world, dev, barrier):
    import torch
    import torch.distributed as dist
    h_wave = torch.empty((B, L), dtype=torch.float32, pin_memory=True)
    h_wave.copy_(wave)
    h_out = torch.empty((B, F, T), dtype=torch.float32, pin_memory=True)
    gaps_np = np.stack([starts, starts + g], 1).astype(np.int32)
    pipe = frontend.HostPipeline(plan, B, L, chunk=64, n_streams=4)     # tools/e2e_sweep.py: 64 x 4 is the best of 8 settings (D2H 50 GB/s)

    def e2e_step():
        pipe.logmag_gap(h_wave, gaps_np, h_out, eps=EPS)

    for _ in range(2):
        e2e_step()
    barrier()
    t0 = time.perf_counter()
    n_e2e = max(3, min(args.steps, 5))
    for _ in range(n_e2e):
        e2e_step()
    barrier()
    dt = (time.perf_counter() - t0) / n_e2e
    if world > 1:
        t = torch.tensor([dt], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dt = float(t.item())
    # the platform's ceiling for this step at this N: the SAME chunked H2D + D2H traffic on the same streams with no kernel
    # between the copies, all ranks at once (PCIe link per GPU + the host's memory / root complexes shared by the ranks)
    def copy_step():
        pipe.logmag_gap(h_wave, gaps_np, h_out, eps=EPS, copies_only=True)

    copy_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(n_e2e):
        copy_step()
    barrier()
    dtc = (time.perf_counter() - t0) / n_e2e
    if world > 1:
        t = torch.tensor([dtc], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dtc = float(t.item())
    h2d, d2h = int(B * L * 4 + B * 8), int(B * F * T * 4)
    return {"value": world * B * CLIP_S / dt, "unit": "audio-s/s", "h2d_bytes_per_step": h2d,
            "d2h_bytes_per_step": d2h, "ms_per_step": dt * 1e3, "steps": n_e2e,
            "api": "ml_audio_inpainting_b200.frontend.HostPipeline.logmag_gap (pinned host in/out, 64-clip chunks on 4 streams)",
            "copy_ceiling": {"value": world * B * CLIP_S / dtc, "unit": "audio-s/s", "ms_per_step": dtc * 1e3,
                             "h2d_GBps_per_gpu": h2d / dtc / 1e9, "d2h_GBps_per_gpu": d2h / dtc / 1e9,
                             "aggregate_GBps": world * (h2d + d2h) / dtc / 1e9,
                             "what": "the same chunked H2D + D2H copies on the same streams with no kernel, all ranks concurrently"},
            "frac_of_copy_ceiling": dtc / dt}


def cnnblstm_leg(args, frontend, sp, world, dev, barrier):
    """BASELINE configs[4]: end-to-end CNN-BLSTM inference (models/model_eval.py:48-194 for a batch), 256 clips x 5 s over 8 GPUs =
    32 per GPU: waveform in pinned host memory -> H2D -> eval_cnnlstm_batch (STFT, phase, spectrum-domain gap, log10) -> the
    reference's StackedBLSTMCNN architecture (cnn_blstm.yaml, random init, eval; stock cuDNN, out of scope) -> blend + 10** +
    phase reuse + iSTFT in one kernel -> D2H.  Only 2 x 320 KB per clip cross PCIe."""
    import torch
    import torch.distributed as dist
    from tools.cnnblstm_model import StandInBLSTMCNN
    Bm, L5 = args.model_clips, 80000
    torch.manual_seed(0)
    model = StandInBLSTMCNN().to(dev).eval()
    gen = torch.Generator().manual_seed(4321)
    h_wave = (0.1 * torch.randn((Bm, L5), generator=gen)).clamp_(-1, 1).pin_memory()
    h_out = torch.empty((Bm, 79872), dtype=torch.float32).pin_memory()
    d_wave = torch.empty((Bm, L5), dtype=torch.float32, device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)      # 256 MB > L2: written between timed iterations
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(5)]

    def step(record=False):
        if record: ev[0].record()
        d_wave.copy_(h_wave, non_blocking=True)
        if record: ev[1].record()
        fe = frontend.eval_cnnlstm_batch(d_wave)
        if record: ev[2].record()
        with torch.no_grad():
            raw = model(fe["log_impaired_magnitude"].unsqueeze(1))
        if record: ev[3].record()
        y = frontend.cnnblstm_backend_batch(raw, fe["log_impaired_magnitude"], fe["mask"], fe["original_phase"])
        h_out.copy_(y, non_blocking=True)
        if record: ev[4].record()

    for _ in range(3):
        step()
    barrier()
    n = max(3, min(args.steps, 10))
    tot, parts = [], []
    for _ in range(n):
        flush.fill_(1)
        barrier()
        t0 = time.perf_counter()
        step(record=True)
        torch.cuda.synchronize()
        tot.append(time.perf_counter() - t0)
        parts.append([ev[i].elapsed_time(ev[i + 1]) for i in range(4)])
    dt = float(np.mean(tot))
    if world > 1:
        t = torch.tensor([dt], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dt = float(t.item())
    p = np.mean(np.array(parts), 0)
    return {"workload": f"configs[4]: end-to-end CNN-BLSTM inference, {Bm} clips x 5 s per GPU ({world * Bm} over {world} GPU(s); batch 256 on "
                        "8 GPUs = 32 per GPU), pinned host waveform in -> pinned host waveform out",
            "value": world * Bm * 5.0 / dt, "unit": "audio-s/s", "ms_per_step": dt * 1e3, "steps": n,
            "ms_h2d": float(p[0]), "ms_front_end": float(p[1]), "ms_model": float(p[2]), "ms_back_end_and_d2h": float(p[3]),
            "front_plus_back_share": float((p[1] + p[3]) / max(p.sum(), 1e-9)),
            "model": "tools/cnnblstm_model.py: the reference's StackedBLSTMCNN architecture (state_dict compatible), random init, fp32 eval",
            "h2d_bytes_per_step": Bm * L5 * 4, "d2h_bytes_per_step": Bm * 79872 * 4,
            "l2": "a 256 MB buffer is written between timed iterations"}


def emit_line(args, world, value, ms, B, T, roofline, cpu_baseline, e2e, clocks, legs):
    line = {
        "metric": METRIC, "value": value, "unit": "audio-s/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": make_config(B, world),
        "metric_note": "the timed step (value, roofline, e2e, cpu_baseline) is the forward half the metric names -- STFT -> "
                       "log-magnitude + gap mask, BASELINE configs[1]; the inverse half (configs[2]), the STFT -> mask -> iSTFT round "
                       "trip and configs[4] are measured in the same run and copied to the top-level keys istft / roundtrip / cnnblstm_e2e",
        "roofline": roofline, "cpu_baseline": cpu_baseline, "e2e": e2e, "gpu_launches": args.steps,
        "clocks": clocks,
        "istft": legs.get("istft"), "roundtrip": legs.get("roundtrip"), "cnnblstm_e2e": legs.get("cnnblstm_e2e"),
        "legs": {k: v for k, v in legs.items() if k not in ("istft", "roundtrip", "cnnblstm_e2e")},
    }
    print(json.dumps(line), file=_RESULT_OUT, flush=True)


if __name__ == "__main__":
    main()
