#!/usr/bin/env python
"""Headline benchmark: channel-pair-frequencies / second for coherence + pairwise spectral
Granger on the BASELINE.json config-4 workload (256 ch x 64 trials x 60 s @ 1 kHz, 7 tapers).

    python bench.py --gpus N --steps K --warmup W              # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...    # CPU arm: the UNMODIFIED reference (baseline/_ref)
    python bench.py --impl replay ...                          # torch replay of the reference's CuPy call sequence

A "step" is one pass of the hot path (Multitaper -> CSM -> coherence_magnitude +
pairwise_spectral_granger_prediction) over ONE synthetic recording.  `value` is measured with
the recording already resident in HBM; `e2e` runs the same step through the public API from a
pinned HOST array and reads both result arrays back to the host inside the timed region.

N > 1 (one process per GPU, torchrun) is STRONG scaling by default: the one config-4 recording is split over the
ranks -- `--shard windows` (default): contiguous window ranges, no collective on the data path;
`--shard trials`: every rank holds 1/N of the trials of all windows, the partial cross-spectral sums are
reduce-scattered along the window axis (NCCL, chunked, on a side stream under the next chunk's FFT + CSM) and each
rank runs coherence + Wilson/Granger on its 1/N of the windows.  `--scaling weak` gives every rank its own
recording.  The step time is the max over ranks; `value` = pair-freqs of all ranks / that time.
"""
import argparse
import ctypes
import json
import math
import os
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: N, T, S, fs, NW, window seconds
    "cfg4": dict(N=60_000, T=64, S=256, fs=1000.0, NW=4.0, duration=1.0),
    "cfg4w4": dict(N=4_000, T=64, S=256, fs=1000.0, NW=4.0, duration=1.0),  # 4 windows of cfg4 (profiling)
    "cfg4w8": dict(N=8_000, T=64, S=256, fs=1000.0, NW=4.0, duration=1.0),  # 8 windows of cfg4 (profiling)
    "cfg3": dict(N=30_000, T=32, S=128, fs=1000.0, NW=4.0, duration=1.0),
    "cfg2": dict(N=10_000, T=16, S=64, fs=1000.0, NW=3.0, duration=1.0),
    "cfg1": dict(N=1000, T=4, S=8, fs=500.0, NW=2.0, duration=None),
    # SURVEY.md config 5 (512 ch x 128 trials @ 2 kHz, 9 tapers, 60 ms windows): full size (1000 windows; meant for
    # `--gpus 8 --shard trials`: 31 GB recording) and reduced-window variants
    "cfg5": dict(N=120_000, T=128, S=512, fs=2000.0, NW=5.0, duration=0.060, recipe="no_line"),
    "cfg5w64": dict(N=7_680, T=128, S=512, fs=2000.0, NW=5.0, duration=0.060, recipe="no_line"),
    "cfg5w8": dict(N=960, T=128, S=512, fs=2000.0, NW=5.0, duration=0.060, recipe="no_line"),
}
MEASURES = ["coherence_magnitude", "pairwise_spectral_granger_prediction"]
# the other BASELINE.json configs are parity-test cases; `--workload` can still time them
WORKLOAD_MEASURES = {"cfg1": ["coherence_magnitude"], "cfg2": ["power", "coherency"],
                     "cfg3": ["expectation_cross_spectral_matrix", "weighted_phase_lag_index"],
                     "cfg5": ["canonical_coherence", "directed_transfer_function"],
                     "cfg5w64": ["canonical_coherence", "directed_transfer_function"],
                     "cfg5w8": ["canonical_coherence", "directed_transfer_function"]}
METRIC = "channel-pair-freqs/sec (CSM+coherence+Granger)"

# ---- executed-work model of the dominant kernel (granger_herm_kernel<2, Plan1000>), DESIGN.md section 4 ----------
# per Wilson iteration: 2 inverse + 2 forward packed complex FFTs of length nfft, counted with the usual
# 5 n log2 n, times 5/6 because the last inverse / first forward radix-10 stage only produces / consumes half of its
# points (lags >= nfft/2 are zeroed by the plus operator); plus the per-bin 2x2 algebra counted from the source
# (herm_iteration: predictor 118, update G + G(P0-I) + G(P-P0) with its maximum 119 flops per bin, one flop per
# add/mul; 156 before the update was restructured).  The FFT count is the conventional 5 n log2 n: the prime-factor
# radix-10 butterflies execute fewer real operations than that, so `achieved` is not inflated by them.
FFTS_PER_ITERATION = 4
FFT_PRUNE = 5.0 / 6.0
ALGEBRA_FLOPS_PER_BIN = 118 + 119
FIXED_FP64_FLOPS_PER_BIN = 60        # load/lag-0 sums, Cholesky, transfer function, noise covariance, log ratio


def geometry(wl):
    n = wl["N"] if wl["duration"] is None else int(np.around(wl["duration"] * wl["fs"]))
    n_win = int(np.floor(wl["N"] / n - n / n + 1))
    from scipy.fft import next_fast_len
    nfft = int(next_fast_len(n))
    return n, n_win, nfft, nfft // 2 + 1


def pair_freqs(wl):
    _, n_win, _, fnn = geometry(wl)
    return n_win * fnn * wl["S"] ** 2


def measures_of(wl_name):
    return WORKLOAD_MEASURES.get(wl_name, MEASURES)


def workload_config(wl_name, wl, n_gpus, scaling="strong", shard="windows"):
    n, n_win, nfft, fnn = geometry(wl)
    digit = int(wl_name[3])
    if n_gpus == 1:
        par = "1 GPU"
    elif scaling == "weak":
        par = f"weak: one full recording per GPU x{n_gpus}, no collective"
    elif shard == "windows":
        par = (f"strong: ONE recording, {n_win} windows sharded over {n_gpus} GPUs (contiguous ranges, "
               "no collective on the data path)")
    else:
        par = (f"strong: ONE recording, trials sharded {wl['T']}/{n_gpus} per GPU; partial cross-spectral sums "
               "reduce-scattered along the window axis (chunked, on a side stream under the next chunk's FFT+CSM: "
               "NCCL, or with --reduce-impl p2p one fused pull-reduce + power + coherence kernel over NVLink peer "
               "memory), Wilson window-sharded")
    return {"workload": f"BASELINE.json configs[{digit - 1}] (SURVEY.md config {digit})"
                        f"{' (reduced windows)' if len(wl_name) > 4 else ''}: {wl['S']}-channel x {wl['T']}-trial x "
                        f"{wl['N'] / wl['fs']:g} s @ {wl['fs']:.0f} Hz, {int(2 * wl['NW'] - 1)} tapers, "
                        f"{n_win} windows of {n} samples, nfft {nfft}; " + " + ".join(measures_of(wl_name)) +
                        (" (Wilson tol 1e-8, <=60 it)" if wl_name not in WORKLOAD_MEASURES else ""),
            "recording": [wl["N"], wl["T"], wl["S"]], "pair_freqs_per_recording": pair_freqs(wl),
            "parallelism": par,
            "l2": "inputs (3.9 GB) and every intermediate exceed the 126 MB L2; no explicit flush"}


# --------------------------------------------------------------------------- #
# CPU arm: the live reference (baseline/_ref) on a bounded sample of the workload; the NumPy oracle port only when
# the reference install is missing
# --------------------------------------------------------------------------- #
def cpu_sample(wl, wl_name, seed=0, n_channels=16):
    """One bounded CPU step.  Returns dict(kind, seconds, units, sample, stage_seconds, est_full_step_seconds)."""
    from oracle import oracle as O
    from baseline import ref_arm
    n, n_win, nfft, fnn = geometry(wl)
    measures = measures_of(wl_name)
    s_full = wl["S"]
    s = min(n_channels, s_full)
    if ref_arm.available():
        r = ref_arm.bounded_sample(O.synthetic_series, wl, measures, s, seed=seed)
        kind = "reference"
        secs, units, stages = r["seconds"], r["units"], r["stage_seconds"]
    else:  # the oracle port (same algorithm restated in NumPy); only when baseline/_ref did not travel
        kind = "port"
        x = O.synthetic_series(n, wl["T"], s, wl["fs"], seed=20261017 + seed)
        t0 = time.perf_counter()
        taps = O.dpss_tapers(n, wl["NW"], O.default_n_tapers(wl["NW"]), wl["fs"])
        coef = O.multitaper_fft(x, wl["fs"], taps, n, n, nfft)
        stages = {"multitaper_fft": time.perf_counter() - t0}
        for name in measures:
            t1 = time.perf_counter()
            if name == "coherence_magnitude":
                O.coherence_magnitude(coef)
            elif name == "pairwise_spectral_granger_prediction":
                O.pairwise_granger(O.expected_csm(coef), O.power(coef))
            else:
                getattr(O, name)(coef) if name != "canonical_coherence" else O.canonical_coherence(
                    coef, np.arange(s) // max(1, s // 2))
            stages[name] = time.perf_counter() - t1
        secs = time.perf_counter() - t0
        units = fnn * s * s
    # linear extrapolation to the full step, reported SEPARATELY from the measurement: windows are a batch axis of
    # the reference, its coherence cost is ~ S^2 (un-averaged CSM) and its Granger cost ~ the pair count
    s2 = (s_full / s) ** 2
    pair_ratio = (s_full * (s_full - 1)) / max(s * (s - 1), 1)
    est = n_win * sum(v * (pair_ratio if k == "pairwise_spectral_granger_prediction" else
                           (s_full / s if k == "multitaper_fft" else s2)) for k, v in stages.items())
    sample = (f"1 of {n_win} windows, {s} of {s_full} channels (all {wl['T']} trials, all tapers, all "
              f"{s * (s - 1) // 2} pairs of those channels) through the "
              f"{'unmodified reference package (baseline/_ref)' if kind == 'reference' else 'NumPy oracle port'}: "
              + ", ".join(f"{k} {v:.2f}s" for k, v in stages.items())
              + f"; {units} pair-freqs in {secs:.2f}s measured (pair-freqs/s is intensive in windows and ~S^2 for "
                "the reference, so the sample's rate stands for the workload's)")
    return dict(kind=kind, seconds=secs, units=units, sample=sample, stage_seconds=stages,
                est_full_step_seconds=est, n_channels=s)


def cpu_threads():
    from baseline import ref_arm
    return ref_arm.cpu_threads()


def run_reference(args, wl_name, wl):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    for _ in range(args.warmup):
        cpu_sample(wl, wl_name, n_channels=4)
    rates, secs, last = [], [], None
    for i in range(args.steps):
        last = cpu_sample(wl, wl_name, seed=i)
        rates.append(last["units"] / last["seconds"])
        secs.append(last["seconds"])
    value = float(np.mean(rates))
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "pair-freqs/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": float(np.mean(secs)) * 1e3,
        "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(wl_name, wl, args.gpus, args.scaling, "windows" if args.shard == "auto" else args.shard),
        "cpu_baseline": {"value": value, "unit": "pair-freqs/s", "cores": cpu_threads(), "kind": last["kind"],
                         "sample": last["sample"], "host_cpus": os.cpu_count(),
                         "sample_pair_freqs": last["units"], "measured_seconds_per_step": float(np.mean(secs)),
                         "est_full_step_seconds": last["est_full_step_seconds"],
                         "note": "ms_per_step and value are MEASURED on the bounded sample (one step = one sample); "
                                 "est_full_step_seconds is the linear extrapolation to the whole workload and is "
                                 "never used as a measurement; BLAS/pocketfft threads as numpy configures them, the "
                                 "reference's Python pair loop is serial"},
        "e2e": {"value": value, "unit": "pair-freqs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def run_replay(args, wl_name, wl):
    """Library-call replay of the reference's CuPy backend on one GPU (baseline/replay.py), bounded sample."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch
    from oracle import oracle as O
    from baseline import replay
    assert torch.cuda.is_available()
    dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0")))
    n, n_win, nfft, fnn = geometry(wl)
    measures = [m for m in measures_of(wl_name) if m in ("coherence_magnitude", "pairwise_spectral_granger_prediction")]
    s = min(args.replay_channels, wl["S"])
    taps = O.dpss_tapers(n, wl["NW"], O.default_n_tapers(wl["NW"]), wl["fs"])
    xs = [O.synthetic_series(n, wl["T"], s, wl["fs"], seed=20261017 + i) for i in range(max(args.steps, 1))]
    for _ in range(args.warmup):
        replay.sample_step(xs[0][:, :, :4].copy(), taps, n, n, nfft, wl["fs"], measures, dev)
    secs, its = [], None
    for i in range(args.steps):
        t, out = replay.sample_step(xs[i], taps, n, n, nfft, wl["fs"], measures, dev)
        secs.append(t)
        its = out.get("_wilson_iterations", its)
    units = fnn * s * s
    value = units / float(np.mean(secs))
    line = {
        "impl": "replay", "metric": METRIC, "value": value, "unit": "pair-freqs/s", "n_gpus": 1, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": float(np.mean(secs)) * 1e3, "higher_is_better": True,
        "scaling": args.scaling, "vs_baseline": None, "dtype": "c128", "data": "synthetic",
        "config": workload_config(wl_name, wl, 1),
        "replay": {"label": "library-call replay of the reference's CuPy backend (torch ops -> cuFFT / cuBLAS / "
                            "cuSOLVER; cupy itself is not installable here) -- NOT CuPy",
                   "sample": f"1 of {n_win} windows, {s} of {wl['S']} channels, all {wl['T']} trials / tapers, all "
                             f"{s * (s - 1) // 2} pairs, complex128, host array in -> host results out; "
                             f"{units} pair-freqs per step", "mean_wilson_iterations": its,
                   "ms_each": [round(t * 1e3, 1) for t in secs]},
        "e2e": {"value": value, "unit": "pair-freqs/s", "h2d_bytes_per_step": int(xs[0].nbytes),
                "d2h_bytes_per_step": int(len(measures) * units * 8)},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# --------------------------------------------------------------------------- #
# GPU arm
# --------------------------------------------------------------------------- #
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.path = index, None, f"/tmp/sc_clocks_{os.getpid()}.csv"

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.index), "-lms", "200"], stdout=open(self.path, "w"),
                                         stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons, power = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in open(self.path):
            f = [c.strip() for c in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2])); power.append(float(f[3]))
            except ValueError:
                continue
            for nm, val in zip(names, f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(nm)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        hi = [s for s, p in zip(sm, power) if p >= 0.5 * max(power)] or sm
        return {"sm_mhz": float(np.median(hi)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons),
                "power_w_max": float(max(power)), "samples": len(sm)}


class native_stdout_to_stderr:
    """Route file descriptor 1 to stderr while native libraries initialise (their banners bypass sys.stdout);
    always restored, and a no-op if the descriptors cannot be duplicated."""

    def __enter__(self):
        self.saved = None
        try:
            sys.stdout.flush()
            self.saved = os.dup(1)
            os.dup2(2, 1)
        except OSError:
            if self.saved is not None:
                os.close(self.saved)
                self.saved = None
        return self

    def __exit__(self, *exc):
        if self.saved is not None:
            try:
                sys.stdout.flush()
            finally:
                os.dup2(self.saved, 1)
                os.close(self.saved)
        return False


def make_recording(wl, seed, device, rows=None, trials=None):
    """Same structure as oracle.synthetic_series (noise + lag-1 even->odd coupling + 40 Hz line; recipe "no_line":
    without the common sinusoid -- the config-5 recipe on which the reference's full-matrix Wilson iteration
    converges, tests/golden/make_golden.py:series_512), generated on the device: building 983 M samples with NumPy
    would take minutes.  ``rows`` = (s0, s1) / ``trials`` = (t0, t1) keep only that slab (the rest is generated and
    dropped, so shards of one seed are slices of one and the same recording)."""
    import torch
    g = torch.Generator(device=device)
    g.manual_seed(20261017 + seed)
    n_rows, T, S = wl["N"], wl["T"], wl["S"]
    block = max(1, min(n_rows, (1 << 28) // (T * S)))     # generate in time blocks of <= 1 GiB
    parts, prev_last = [], None
    r_lo, r_hi = rows if rows is not None else (0, n_rows)
    t_lo, t_hi = trials if trials is not None else (0, T)
    ph = 2 * math.pi * torch.arange(S, device=device, dtype=torch.float32) / S
    for r0 in range(0, n_rows, block):
        r1 = min(n_rows, r0 + block)
        noise = torch.randn((r1 - r0, T, S), generator=g, device=device, dtype=torch.float32)
        x = noise.clone()
        n_odd = x[:, :, 1::2].shape[-1]
        x[1:, :, 1::2] += 0.5 * noise[:-1, :, 0::2][:, :, :n_odd]
        if prev_last is not None:
            x[0, :, 1::2] += 0.5 * prev_last[:, 0::2][:, :n_odd]
        prev_last = noise[-1].clone()
        if wl.get("recipe") != "no_line":
            t = torch.arange(r0, r1, device=device, dtype=torch.float32) / wl["fs"]
            x += 0.5 * torch.sin(2 * math.pi * 40.0 * t[:, None, None] + ph[None, None, :])
        a, b = max(r0, r_lo), min(r1, r_hi)
        if b > a:
            parts.append(x[a - r0:b - r0, t_lo:t_hi].clone())
        del x, noise
    return torch.cat(parts, dim=0) if len(parts) > 1 else parts[0]


def measure_simt_peaks():
    """FP32 / FP64 FMA peaks of this device, measured live (sc_simt_peak, CUDA events, best of 5)."""
    import torch
    from spectral_connectivity_b200 import _lib
    lib = _lib.load()
    scratch = torch.empty(lib.sc_simt_peak_scratch_bytes(), dtype=torch.uint8, device="cuda")
    out = {}
    for name, dtype, n in (("fp32", 0, 1 << 14), ("fp64", 1, 1 << 13)):
        flops, best = ctypes.c_double(0.0), 0.0
        for _ in range(6):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            _lib.check(lib.sc_simt_peak(dtype, n, _lib.ptr(scratch), ctypes.byref(flops), _lib.stream_ptr()), "sc_simt_peak")
            b.record()
            torch.cuda.synchronize()
            best = max(best, flops.value / (a.elapsed_time(b) * 1e-3) / 1e12)
        out[name] = best
    return out


def init_gpu():
    """Device, CPU affinity and (N > 1) the NCCL process group of this rank -> (rank, world, local, dev, cpus)."""
    import torch
    import torch.distributed as dist

    from spectral_connectivity_b200 import _lib
    from spectral_connectivity_b200.distributed import bind_to_local_cpus
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback)"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    cpus = bind_to_local_cpus(local) if world > 1 else None    # node-local pinned buffers, before any allocation
    if world > 1:
        # keep stdout to the one JSON line: the NCCL communicator setup prints "NCCL version ..." to stdout
        with native_stdout_to_stderr():
            dist.init_process_group("nccl", device_id=dev)
            dist.barrier()
    _lib.load()
    return rank, world, local, dev, cpus


def run_gpu(args, wl_name, wl, shard, ctx):
    """One measurement (device-resident + end to end) in the given sharding mode; rank 0 returns the JSON line."""
    import torch
    import torch.distributed as dist

    import spectral_connectivity_b200 as sc
    from spectral_connectivity_b200 import _lib
    from spectral_connectivity_b200.distributed import shard_recording

    rank, world, local, dev, cpus = ctx
    n, n_win, nfft, fnn = geometry(wl)
    strong = world > 1 and args.scaling == "strong"
    by_trials = strong and shard == "trials"
    measures = measures_of(wl_name)
    has_granger = "pairwise_spectral_granger_prediction" in measures
    kw = dict(sampling_frequency=wl["fs"], time_halfbandwidth_product=wl["NW"],
              time_window_duration=wl["duration"])
    n_local_win = n_win
    if strong and not by_trials:
        w0, w1, s0, s1 = shard_recording(wl["N"], n, n, rank, world)
        n_local_win = w1 - w0
        x_dev = make_recording(wl, 0, dev, rows=(s0, s1))
    elif by_trials:
        if wl["T"] % world:
            raise SystemExit(f"--shard trials needs the trial count ({wl['T']}) to be a multiple of --gpus")
        tpr = wl["T"] // world
        x_dev = make_recording(wl, 0, dev, trials=(rank * tpr, (rank + 1) * tpr))
    else:
        x_dev = make_recording(wl, rank, dev)
    units_total = pair_freqs(wl) * (1 if strong else world)
    group = dist.group.WORLD if by_trials else None
    ckw = dict(reduce_group=group, reduce_mode="reduce_scatter", reduce_impl=args.reduce_impl) if by_trials else {}

    def build(x, output):
        m = sc.Multitaper(x, **kw)
        if strong and not by_trials:
            m._n_time_windows_override = n_local_win
        return sc.Connectivity.from_multitaper(m, output=output, **ckw)

    def run_measures(c, out=None, packed=()):
        fused = [name for name in measures if name in sc.connectivity.MEASURES]
        res = c.compute(fused, out=out, packed=packed) if fused else {}
        for name in measures:
            if name == "canonical_coherence":      # 64-channel groups
                with _lib.timed(name):
                    res[name] = c.canonical_coherence(np.arange(wl["S"]) // 64)[0]
            elif name not in fused:
                with _lib.timed(name):
                    res[name] = getattr(c, name)()
        return res

    def step_device():
        c = build(x_dev, "torch")
        return c, run_measures(c)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(v):
        if world == 1:
            return v
        t = torch.tensor([v], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- device-resident throughput ------------------------------------------------
    c = out = None
    for _ in range(args.warmup):
        c, out = step_device()
    del out
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    _lib.TIMER = _lib.StageTimer()
    launches0 = _lib.LAUNCHES
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        c, out = step_device()
    e1.record()
    barrier()
    ms_total = max_over_ranks(e0.elapsed_time(e1))
    launches = _lib.LAUNCHES - launches0
    stages = _lib.TIMER.totals()
    _lib.TIMER = None
    clocks = sampler.stop() if rank == 0 else None
    ms_step = ms_total / args.steps
    value = units_total / (ms_step * 1e-3)
    mean_iters, n_problems, sanity, executed = 0.0, 0, {}, None
    if has_granger:
        iters = c.last_granger_iterations.to(torch.float64)
        mean_iters = float(iters.mean()) if iters.numel() else 0.0
        n_problems = iters.numel()
        flags = int(c.last_granger_flags.ne(0).sum())
        executed = [int(v) for v in c.last_granger_executed.tolist()]   # of the LAST step, this rank
        gc = out["pairwise_spectral_granger_prediction"]
        coh = out["coherence_magnitude"]
        sanity = {"granger_nan_frac": float(torch.isnan(gc).float().mean()) if gc.numel() else None,
                  "granger_max": float(torch.nan_to_num(gc, nan=0.0).max()) if gc.numel() else None,
                  "coherence_mean_offdiag": float(torch.nanmean(coh)) if coh.numel() else None,
                  "wilson_flagged": flags, "wilson_mean_iters": mean_iters}
        del gc, coh
    elif hasattr(c, "last_wilson_flags") and c.last_wilson_flags is not None:
        sanity = {"wilson_flagged": int(c.last_wilson_flags.ne(0).sum()),
                  "wilson_mean_iters": float(c.last_wilson_iterations.float().mean())}
    del out

    # ---- end to end through the public API with host buffers -----------------------------
    x_host = torch.empty(x_dev.shape, dtype=torch.float32, pin_memory=True)
    x_host.copy_(x_dev)
    torch.cuda.synchronize()
    x_np = x_host.numpy()
    del c
    d2h, e2e_error, bufs = 0, None, None
    e2e_extra_warmup = 0

    def step_e2e():
        cc = build(x_np, "numpy")                      # H2D from pinned host memory; D2H of every result
        return run_measures(cc, out=bufs)

    try:
        if args.no_e2e:
            raise RuntimeError("end-to-end part skipped (--no-e2e)")
        # persistent page-locked result buffers, as a pipeline calling compute() repeatedly would hold them
        # (compute(out=...)): nothing is allocated or page-locked inside the timed region
        fused = [name for name in measures if name in sc.connectivity.MEASURES]
        probe = step_e2e()
        d2h = int(sum(v.nbytes for v in probe.values()))
        if fused:
            bufs = {name: sc.pinned_empty(probe[name].shape, probe[name].dtype) for name in fused}
        del probe
        for _ in range(max(args.e2e_warmup - 1, 0)):  # untimed: staging buffers and allocator reach steady state
            res = step_e2e()
            d2h = int(sum(v.nbytes for v in res.values()))
            del res
    except (RuntimeError, MemoryError) as exc:  # e.g. the host cannot pin N ranks x (inputs + results)
        e2e_error = f"{type(exc).__name__}: {str(exc)[:200]}"
    # every rank must take the same path through the barriers below
    e2e_ok = max_over_ranks(0.0 if e2e_error is None else 1.0) == 0.0
    barrier()
    e2e_steps = max(1, min(args.steps, 5))
    e2e_each = []
    e2e_ms = e2e_value = None
    if e2e_ok:
        # ... and, with several ranks sharing one host, until the step time has settled (page-locked buffers of N
        # ranks are first touched and migrated during the first passes: 286 / 146 / 127 ms were seen as "timed"
        # steps at N = 8 after five warm-ups): up to 10 more untimed passes, stop when two consecutive ones agree
        # to 10 % on every rank
        prev = None
        for _ in range(10 if world > 1 else 0):
            barrier()
            t1 = time.perf_counter()
            res = step_e2e()
            del res
            barrier()
            cur = max_over_ranks((time.perf_counter() - t1) * 1e3)
            settled = prev is not None and abs(cur - prev) <= 0.1 * cur
            prev = cur
            e2e_extra_warmup += 1
            if settled:
                break
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            t1 = time.perf_counter()
            res = step_e2e()
            del res
            e2e_each.append((time.perf_counter() - t1) * 1e3)
        barrier()
        e2e_ms = max_over_ranks((time.perf_counter() - t0) * 1e3) / e2e_steps
        e2e_value = units_total / (e2e_ms * 1e-3)
    # ---- the same end-to-end step with the SYMMETRIC results as packed upper triangles (compute(packed=...)): an
    # opt-in host format (not the reference's), reported beside the drop-in number because the end-to-end step is
    # bound by the device->host link once the kernels are faster than it
    e2e_packed = None
    sym = [name for name in measures if name in sc.connectivity.SYMMETRIC_MEASURES]
    if e2e_ok and sym and max_over_ranks(1.0 if args.no_packed_e2e else 0.0) == 0.0:
        bufs_p, d2h_p, err_p = None, 0, None
        try:                                            # staging may fail on one rank only: agree before any barrier
            probe = run_measures(build(x_np, "numpy"), packed=sym)
            d2h_p = int(sum(v.nbytes for v in probe.values()))
            bufs_p = {name: sc.pinned_empty(probe[name].shape, probe[name].dtype) for name in probe}
            del probe
        except (RuntimeError, MemoryError) as exc:
            err_p = f"{type(exc).__name__}: {str(exc)[:200]}"
        if max_over_ranks(0.0 if err_p is None else 1.0) == 0.0:
            for _ in range(2):
                run_measures(build(x_np, "numpy"), out=bufs_p, packed=sym)
            barrier()
            t0 = time.perf_counter()
            for _ in range(e2e_steps):
                run_measures(build(x_np, "numpy"), out=bufs_p, packed=sym)
            barrier()
            ms_p = max_over_ranks((time.perf_counter() - t0) * 1e3) / e2e_steps
            e2e_packed = {"value": units_total / (ms_p * 1e-3), "unit": "pair-freqs/s", "ms_per_step": ms_p,
                          "d2h_bytes_per_step_rank0": d2h_p, "packed_measures": sym,
                          "note": "same step, symmetric results delivered as packed upper triangles "
                                  "(compute(packed=...), unpack_upper restores the full array): opt-in format, "
                                  "NOT the reference's return shape -- the drop-in number is e2e.value"}
        else:
            e2e_packed = {"error": err_p or "another rank failed to stage its packed host buffers"}
        del bufs_p
    # ---- the host link all ranks share: pinned H2D + D2H copies issued by every rank at the same time ----------
    def host_link_gbs(nbytes=1 << 30, reps=3):
        hs, hd = torch.empty(nbytes, dtype=torch.uint8, pin_memory=True), torch.empty(nbytes, dtype=torch.uint8, pin_memory=True)
        ds, dd = torch.empty(nbytes, dtype=torch.uint8, device=dev), torch.empty(nbytes, dtype=torch.uint8, device=dev)
        s_in, s_out = _lib.side_stream(dev, "h2d"), _lib.side_stream(dev, "d2h")
        barrier()
        t0 = time.perf_counter()
        for _ in range(reps):
            with torch.cuda.stream(s_in):
                dd.copy_(hs, non_blocking=True)
            with torch.cuda.stream(s_out):
                hd.copy_(ds, non_blocking=True)
        barrier()
        dt = max_over_ranks(time.perf_counter() - t0)
        return 2.0 * nbytes * reps * world / dt / 1e9      # aggregate over ranks, both directions

    link_gbs = host_link_gbs()
    h2d_total, d2h_total = int(x_np.nbytes), d2h
    if world > 1:
        t = torch.tensor([h2d_total, d2h_total], dtype=torch.float64, device=dev)
        dist.all_reduce(t)
        h2d_total, d2h_total = int(t[0].item()), int(t[1].item())

    del x_host, x_np, bufs, x_dev
    torch.cuda.empty_cache()
    if rank != 0:
        return None

    # ---- roofline of the dominant kernel + per-stage table (rank 0's kernels) ---------------------------
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    tf32_peak = float(peaks.get("bf16_tflops", 1590.0)) / 2.0
    peak_src = "measured (MEASURED_PEAKS.json)" if peaks else "fallback (B200_PROFILING.md)"
    simt = measure_simt_peaks()
    T, S, K = (wl["T"] // world if by_trials else wl["T"]), wl["S"], int(2 * wl["NW"] - 1)
    tk = T * K                      # observations this rank contracts over (T/N under trial sharding)
    wloc = n_local_win              # windows this rank transforms
    wown = n_problems // max(S * (S - 1) // 2, 1) if has_granger else (n_win // world if by_trials else n_local_win)
    fft_flops = FFTS_PER_ITERATION * FFT_PRUNE * 5.0 * nfft * math.log2(nfft)
    alg_flops = ALGEBRA_FLOPS_PER_BIN * fnn
    alg = {
        # bytes: series read once + planar half-spectrum coefficients written once
        "mt_fft": ("hbm", 4.0 * wloc * n * T * S + 8.0 * wloc * tk * fnn * S),
        # power: a full pass over the coefficients, or -- when the CSM is computed anyway -- its real diagonal
        "power": ("hbm", (12.0 * wown * fnn * S if ("csm" in stages or has_granger) else
                          8.0 * wloc * tk * fnn * S + 4.0 * wloc * fnn * S)),
        # tensor stage (S >= 96, S % 32 == 0): TF32 flops actually issued = upper-triangular 128x128 tiles x 12 MMAs
        # (Re/Im x (hi*hi + hi*lo + lo*hi) x 2 products) x 2*128*128*8 per 8 observations; smaller S: the SIMT kernel,
        # 8 flops per observation and upper-triangle-tile pair
        "csm": (("tensor", wloc * fnn * (math.ceil(S / 128) * (math.ceil(S / 128) + 1) // 2) * math.ceil(tk / 16) * 2
                 * 12 * 2.0 * 128 * 128 * 8) if (S >= 96 and S % 32 == 0) else
                ("fp32", 8.0 * wloc * fnn * tk * S * (S + 64) / 2)),
        "epilogue": ("hbm", 12.0 * wown * fnn * S * S),
        # fp32 SIMT: per observation and UPPER-TRIANGLE pair (the kernel mirrors) Im(x_i conj x_j) = 3 flops,
        # sign / |.| / square / three sums = 6 flops (PLI family, one pass for all four sums); PLV: the complex
        # product 6, 1/|.| 5, two scaled sums 4
        "pli": ("fp32", 9.0 * wloc * fnn * tk * S * (S + 64) / 2),
        "plv": ("fp32", 15.0 * wloc * fnn * tk * S * (S + 64) / 2),
    }
    stage_rows = {}
    for name, (ms, cnt) in stages.items():
        per_step = ms / args.steps
        row = {"ms_per_step": per_step, "launches_per_step": cnt / args.steps,
               "share_of_step": per_step / ms_step}
        if name in alg:
            kind, amount = alg[name]
            if kind == "hbm":
                row.update(bound="hbm", achieved=amount / (per_step * 1e-3) / 1e9, peak=hbm_peak, unit="GB/s")
            elif kind == "tensor":
                row.update(bound="tensor", achieved=amount / (per_step * 1e-3) / 1e12, peak=tf32_peak, unit="TFLOP/s",
                           peak_source=f"{peak_src} cuBLAS bf16 peak / 2 (TF32 runs at half the bf16 rate)",
                           algorithmic_tflops=8.0 * tk * S * S * wloc * fnn / (per_step * 1e-3) / 1e12)
            else:
                row.update(bound="fp32-simt", achieved=amount / (per_step * 1e-3) / 1e12, peak=simt["fp32"],
                           unit="TFLOP/s", peak_source="measured live: sc_simt_peak FP32 FMA")
            row["frac"] = row["achieved"] / row["peak"]
        stage_rows[name] = row
    if has_granger and "granger" in stages and executed:
        # EXECUTED work of the last step on this rank, from the kernel's own counters
        it32, it64, tail_steps, probs = executed
        f32 = it32 * (fft_flops + alg_flops) + it64 * fft_flops          # defect iterations: fp32 FFTs
        f64 = it64 * alg_flops + probs * FIXED_FP64_FLOPS_PER_BIN * fnn + tail_steps * 60.0
        per_step = stages["granger"][0] / args.steps
        t_min = f32 / (simt["fp32"] * 1e12) + f64 / (simt["fp64"] * 1e12)
        row = stage_rows["granger"]
        row.update(bound="fp32+fp64-simt", achieved=(f32 + f64) / (per_step * 1e-3) / 1e12,
                   peak=(f32 + f64) / t_min / 1e12, unit="TFLOP/s", frac=t_min / (per_step * 1e-3),
                   executed={"problems": probs, "fp32_iterations_per_problem": it32 / max(probs, 1),
                             "fp64_iterations_per_problem": it64 / max(probs, 1),
                             "tail_steps_per_problem": tail_steps / max(probs, 1),
                             "fp32_flop": f32, "fp64_flop": f64,
                             "flop_model": f"per iteration {FFTS_PER_ITERATION} complex FFTs x 5 n log2 n x 5/6 "
                                           f"(= {fft_flops:.0f}) + {ALGEBRA_FLOPS_PER_BIN} flop x {fnn} bins of 2x2 "
                                           "algebra; fp64-phase iterations run their FFTs in fp32 and their algebra "
                                           f"in fp64; + {FIXED_FP64_FLOPS_PER_BIN} fp64 flop x bins per problem"},
                   peak_source=f"measured live (sc_simt_peak): FP32 FMA {simt['fp32']:.1f} TFLOP/s, FP64 FMA "
                               f"{simt['fp64']:.1f} TFLOP/s; peak = executed flop / (fp32 flop / P32 + fp64 flop / P64)",
                   reference_equivalent_tflops=(mean_iters * n_problems * 8 * 5.0 * nfft * math.log2(nfft)
                                                / (per_step * 1e-3) / 1e12))
    # DRAM traffic per launch from the committed ncu capture (scaled to this run's windows per launch)
    traffic = {}
    try:
        tj = json.load(open(os.path.join(ROOT, "profiles", "r02_traffic.json")))
        if wl_name == "cfg4" and world == 1:
            for name, row in stage_rows.items():
                if name in tj["dram_bytes_per_launch"]:
                    per_window = tj["dram_bytes_per_launch"][name] / tj["windows_in_captured_launch"]
                    row["traffic"] = per_window * n_win / max(row["launches_per_step"], 1)
                    traffic[name] = row["traffic"]
    except Exception:
        pass
    timed = {k: v for k, v in stage_rows.items() if k != "collective"}
    dom = max(timed, key=lambda k_: timed[k_]["ms_per_step"]) if timed else None
    roof = None
    if dom:
        r = stage_rows[dom]
        roof = {"kernel": dom, "bound": r.get("bound"), "achieved": r.get("achieved"), "peak": r.get("peak"),
                "unit": r.get("unit"), "frac": r.get("frac"), "traffic": traffic.get(dom),
                "peak_source": r.get("peak_source", f"{peak_src} HBM copy bandwidth"),
                "ms_per_launch": r["ms_per_step"] / max(r["launches_per_step"], 1)}
        if dom == "granger":
            roof["note"] = ("Wilson/Granger is FP32/FP64-SIMT + shared-memory bound, neither HBM nor tensor bound "
                            "(SURVEY.md 8d).  achieved = flops the kernel EXECUTED (its own per-phase iteration "
                            "counters x the flop model in stages.granger.executed) / measured time; peak = the same "
                            "flops at the measured FP32 / FP64 FMA peaks; the reference-equivalent figure "
                            "(reference iterations x 8 FFTs) is kept separately in "
                            "stages.granger.reference_equivalent_tflops")
            roof["hbm_algorithmic_gbs"] = n_problems * fnn * (8 + 4 + 4 + 8) / (r["ms_per_step"] * 1e-3) / 1e9

    scaling = args.scaling if world > 1 else "strong"
    line = {
        "metric": METRIC, "value": value, "unit": "pair-freqs/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": scaling,
        "vs_baseline": None, "dtype": "f32 spectra/CSM, f64 Wilson", "data": "synthetic",
        "config": workload_config(wl_name, wl, world, args.scaling, shard),
        "e2e": {"value": e2e_value, "unit": "pair-freqs/s", "ms_per_step": e2e_ms,
                "h2d_bytes_per_step": h2d_total, "d2h_bytes_per_step": d2h_total, "steps": e2e_steps,
                "warmup": args.e2e_warmup + e2e_extra_warmup, "ms_each_rank0": [round(v, 1) for v in e2e_each],
                "host_buffers": "pinned input; results into persistent pinned buffers (compute(out=...))",
                "cpu_affinity_rank0": cpus,
                "host_link_gbs_all_ranks": link_gbs,
                "host_link_ms_for_step_bytes": ((h2d_total + d2h_total) / (link_gbs * 1e9) * 1e3 if link_gbs else None),
                "host_link_note": "pinned H2D + D2H copies issued by every rank at once (1 GiB x 3 per direction), "
                                  "aggregate GB/s over both directions; host_link_ms_for_step_bytes = this step's "
                                  "H2D + D2H bytes at that rate: the share of the end-to-end step that is the host "
                                  "link, whatever the GPUs do",
                **({} if e2e_ok else {"error": e2e_error or "another rank failed to stage its host buffers"})},
        "e2e_packed": e2e_packed,
        # the same throughput counted in UNIQUE pairs, W Fnn S (S - 1) / 2 per recording
        "value_unique_pairs": value * (wl["S"] - 1) / (2.0 * wl["S"]),
        "gpu_launches": launches, "clocks": clocks, "roofline": roof, "stages": stage_rows, "sanity": sanity,
        "simt_peaks_tflops": simt,
    }
    if world == 1 and not args.no_cpu_baseline:
        cb = cpu_sample(wl, wl_name)
        line["cpu_baseline"] = {"value": cb["units"] / cb["seconds"], "unit": "pair-freqs/s",
                                "cores": cpu_threads(), "kind": cb["kind"], "sample": cb["sample"],
                                "host_cpus": os.cpu_count(), "measured_seconds": cb["seconds"],
                                "sample_pair_freqs": cb["units"],
                                "est_full_step_seconds": cb["est_full_step_seconds"]}
    return line


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference", "replay"])
    ap.add_argument("--workload", default="cfg4", choices=sorted(WORKLOADS))
    ap.add_argument("--scaling", default="strong", choices=["strong", "weak"])
    ap.add_argument("--shard", default="auto", choices=["auto", "windows", "trials"],
                    help="N > 1, strong scaling: 'windows' (no collective), 'trials' (reduce_scatter of the partial "
                         "cross-spectral sums); 'auto' = the headline line is window-sharded and the same run also "
                         "measures the trial-sharded mode, reported under 'trial_sharded'")
    ap.add_argument("--reduce-impl", default="p2p", choices=["nccl", "p2p"],
                    help="--shard trials: NCCL reduce_scatter, or the fused peer-memory reduce + epilogue kernel")
    ap.add_argument("--e2e-warmup", type=int, default=5)
    ap.add_argument("--replay-channels", type=int, default=32)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true", help="profiling runs: skip the end-to-end part")
    ap.add_argument("--no-packed-e2e", action="store_true", help="skip the secondary end-to-end run with packed results")
    args = ap.parse_args()
    wl = WORKLOADS[args.workload]
    if args.impl == "reference":
        run_reference(args, args.workload, wl)
    elif args.impl == "replay":
        run_replay(args, args.workload, wl)
    else:
        import torch.distributed as dist
        ctx = init_gpu()
        rank, world = ctx[0], ctx[1]
        first = "windows" if args.shard == "auto" else args.shard
        line = run_gpu(args, args.workload, wl, first, ctx)
        if args.shard == "auto" and world > 1 and args.scaling == "strong" and wl["T"] % world == 0:
            second = run_gpu(args, args.workload, wl, "trials", ctx)
            if rank == 0:
                line["trial_sharded"] = {k: second[k] for k in ("value", "unit", "ms_per_step", "e2e", "gpu_launches")}
                line["trial_sharded"]["parallelism"] = second["config"]["parallelism"]
                line["trial_sharded"]["reduce_impl"] = args.reduce_impl
                line["trial_sharded"]["stages"] = {k: {"ms_per_step": v["ms_per_step"], "frac": v.get("frac")}
                                                  for k, v in second["stages"].items()}
        if rank == 0:
            print(json.dumps(line))
        if world > 1:
            dist.destroy_process_group()


if __name__ == "__main__":
    main()
