#!/usr/bin/env python
"""Headline benchmark: channel-pair-frequencies / second for coherence + pairwise spectral
Granger on the BASELINE.json config-4 workload (256 ch x 64 trials x 60 s @ 1 kHz, 7 tapers).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # CPU arm: the NumPy oracle port

A "step" is one pass of the hot path (Multitaper -> CSM -> coherence_magnitude +
pairwise_spectral_granger_prediction) over one synthetic recording.  `value` is measured with
the recording already resident in HBM; `e2e` runs the same step through the public API from a
pinned HOST array and reads both result arrays back to the host inside the timed region.
N > 1 is weak scaling: every rank processes its own recording (window shards of an N-times
longer session), no collective on the data path; the step time is the max over ranks.
"""
import argparse
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
    "cfg3": dict(N=30_000, T=32, S=128, fs=1000.0, NW=4.0, duration=1.0),
    "cfg2": dict(N=10_000, T=16, S=64, fs=1000.0, NW=3.0, duration=1.0),
    "cfg1": dict(N=1000, T=4, S=8, fs=500.0, NW=2.0, duration=None),
    # SURVEY.md config 5 geometry (512 ch x 128 trials @ 2 kHz, 9 tapers, 60 ms windows) on 8 of its 1000 windows
    "cfg5w8": dict(N=960, T=128, S=512, fs=2000.0, NW=5.0, duration=0.060),
}
MEASURES = ["coherence_magnitude", "pairwise_spectral_granger_prediction"]
# the other BASELINE.json configs are parity-test cases; `--workload` can still time them
WORKLOAD_MEASURES = {"cfg1": ["coherence_magnitude"], "cfg2": ["power", "coherency"],
                     "cfg3": ["expectation_cross_spectral_matrix", "weighted_phase_lag_index"],
                     "cfg5w8": ["canonical_coherence", "directed_transfer_function"]}
METRIC = "channel-pair-freqs/sec (CSM+coherence+Granger)"
FP64_PEAK_NOMINAL_TFLOPS = 37.0  # B200 FP64 vector, nominal (not in MEASURED_PEAKS.json)


def geometry(wl):
    n = wl["N"] if wl["duration"] is None else int(np.around(wl["duration"] * wl["fs"]))
    n_win = int(np.floor(wl["N"] / n - n / n + 1))
    from scipy.fft import next_fast_len
    nfft = int(next_fast_len(n))
    return n, n_win, nfft, nfft // 2 + 1


def pair_freqs(wl):
    _, n_win, _, fnn = geometry(wl)
    return n_win * fnn * wl["S"] ** 2


# --------------------------------------------------------------------------- #
# CPU arm: the oracle port on a bounded sample, extrapolated (both loops of the reference are
# exactly linear: window axis is a batch dim, connectivity.py:2314 loops over pairs)
# --------------------------------------------------------------------------- #
def cpu_sample(wl, seed=0, s_sub=32, g_channels=16):
    from oracle import oracle as O
    n, n_win, nfft, fnn = geometry(wl)
    s_full = wl["S"]
    s_sub = min(s_sub, s_full)
    x = O.synthetic_series(n, wl["T"], s_full, wl["fs"], seed=20261017 + seed)  # ONE window
    k = O.default_n_tapers(wl["NW"])
    t0 = time.perf_counter()
    taps = O.dpss_tapers(n, wl["NW"], k, wl["fs"])
    coef = O.multitaper_fft(x, wl["fs"], taps, n, n, nfft)
    t_fft = time.perf_counter() - t0
    sub = coef[..., :s_sub]
    t0 = time.perf_counter()
    O.coherence_magnitude(sub, row_block=4)
    t_coh = time.perf_counter() - t0
    all_pairs = s_full * (s_full - 1) // 2
    g_ch = min(g_channels, s_full)
    n_pairs = g_ch * (g_ch - 1) // 2
    gsub = coef[..., :g_ch]
    csm = O.expected_csm(gsub, row_block=8)
    pw = O.power(gsub)
    t0 = time.perf_counter()
    _, its = O.pairwise_granger(csm, pw, return_iterations=True)  # the reference's Python pair loop
    t_gr = time.perf_counter() - t0
    est_window = t_fft + t_coh * (s_full / s_sub) ** 2 + t_gr * (all_pairs / n_pairs)
    est_total = est_window * n_win
    return dict(t_fft=t_fft, t_coh=t_coh, t_granger=t_gr, est_step_seconds=est_total,
                measured_seconds=t_fft + t_coh + t_gr, mean_wilson_iters=float(np.mean(its)),
                sample=(f"1 of {n_win} windows; multitaper FFT on all {s_full} ch ({t_fft:.2f}s), coherence on "
                        f"{s_sub} of {s_full} ch ({t_coh:.2f}s, x{(s_full / s_sub) ** 2:.0f}), Granger pair loop on {n_pairs} "
                        f"of {all_pairs} pairs ({t_gr:.2f}s, x{all_pairs / n_pairs:.0f}; the reference's second CSM "
                        f"evaluation for Granger is NOT counted); extrapolated linearly to the full step = "
                        f"{est_total:.0f}s"))


def cpu_threads():
    try:
        from threadpoolctl import threadpool_info
        return max([p.get("num_threads", 1) for p in threadpool_info()] + [1])
    except Exception:
        return 1


def run_reference(args, wl_name, wl):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    for _ in range(args.warmup):
        cpu_sample(wl, s_sub=8, g_channels=3)
    rates, secs, last = [], [], None
    for i in range(args.steps):
        last = cpu_sample(wl, seed=i)
        rates.append(pair_freqs(wl) / last["est_step_seconds"])
        secs.append(last["est_step_seconds"])
    value = float(np.mean(rates))
    n, n_win, nfft, fnn = geometry(wl)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "pair-freqs/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": float(np.mean(secs)) * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(wl_name, wl, args.gpus),
        "cpu_baseline": {"value": value, "unit": "pair-freqs/s", "cores": cpu_threads(), "kind": "port",
                         "sample": last["sample"], "host_cpus": os.cpu_count(),
                         "note": "NumPy oracle port of the reference algorithm (oracle/oracle.py); the Python "
                                 "pair loop is serial, BLAS/pocketfft threads as numpy configures them"},
        "e2e": {"value": value, "unit": "pair-freqs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def workload_config(wl_name, wl, n_gpus):
    n, n_win, nfft, fnn = geometry(wl)
    return {"workload": f"BASELINE.json configs[{int(wl_name[3]) - 1}] (SURVEY.md config {wl_name[3]}){' (reduced windows)' if len(wl_name) > 4 else ''}: {wl['S']}-channel x {wl['T']}-trial x "
                        f"{wl['N'] / wl['fs']:g} s @ {wl['fs']:.0f} Hz, {int(2 * wl['NW'] - 1)} tapers, "
                        f"{n_win} windows of {n} samples, nfft {nfft}; " + " + ".join(WORKLOAD_MEASURES.get(wl_name, MEASURES)) +
                        (" (Wilson tol 1e-8, <=60 it)" if wl_name not in WORKLOAD_MEASURES else ""),
            "per_gpu_recording": [wl["N"], wl["T"], wl["S"]], "pair_freqs_per_gpu_step": pair_freqs(wl),
            "parallelism": f"window-sharded x{n_gpus} (one recording shard per GPU, no collective)",
            "l2": "inputs (3.9 GB) and every intermediate exceed the 126 MB L2; no explicit flush"}


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


def make_recording(wl, seed, device):
    """Same structure as oracle.synthetic_series (noise + lag-1 even->odd coupling + 40 Hz line),
    generated on the device: building 983 M samples with NumPy would take minutes."""
    import torch
    g = torch.Generator(device=device)
    g.manual_seed(20261017 + seed)
    x = torch.randn((wl["N"], wl["T"], wl["S"]), generator=g, device=device, dtype=torch.float32)
    x[1:, :, 1::2] += 0.5 * x[:-1, :, 0::2][:, :, : x[:, :, 1::2].shape[-1]]
    t = torch.arange(wl["N"], device=device, dtype=torch.float32) / wl["fs"]
    ph = 2 * math.pi * torch.arange(wl["S"], device=device, dtype=torch.float32) / wl["S"]
    x += 0.5 * torch.sin(2 * math.pi * 40.0 * t[:, None, None] + ph[None, None, :])
    return x


def run_gpu(args, wl_name, wl):
    import torch
    import torch.distributed as dist

    import spectral_connectivity_b200 as sc
    from spectral_connectivity_b200 import _lib

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback)"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        # keep stdout to the one JSON line: the NCCL communicator setup prints "NCCL version ..." to stdout
        with native_stdout_to_stderr():
            dist.init_process_group("nccl", device_id=dev)
            dist.barrier()
    _lib.load()

    x_dev = make_recording(wl, rank, dev)
    n, n_win, nfft, fnn = geometry(wl)
    units = pair_freqs(wl)
    kw = dict(sampling_frequency=wl["fs"], time_halfbandwidth_product=wl["NW"],
              time_window_duration=wl["duration"])

    measures = WORKLOAD_MEASURES.get(wl_name, MEASURES)
    has_granger = "pairwise_spectral_granger_prediction" in measures

    def run_measures(c):
        fused = [name for name in measures if name in sc.connectivity.MEASURES]
        out = c.compute(fused) if fused else {}
        for name in measures:
            if name == "canonical_coherence":      # 64-channel groups
                out[name] = c.canonical_coherence(np.arange(wl["S"]) // 64)[0]
            elif name not in fused:
                with _lib.timed(name):
                    out[name] = getattr(c, name)()
        return out

    def step_device():
        m = sc.Multitaper(x_dev, **kw)
        c = sc.Connectivity.from_multitaper(m, output="torch")
        out = run_measures(c)
        return c, out

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        if world == 1:
            return ms
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
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
    value = units * world / (ms_step * 1e-3)
    mean_iters, n_problems, sanity = 0.0, 0, {}
    if has_granger:
        iters = c.last_granger_iterations.to(torch.float64)
        mean_iters = float(iters.mean())
        n_problems = iters.numel()
        flags = int(c.last_granger_flags.ne(0).sum())
        gc = out["pairwise_spectral_granger_prediction"]
        coh = out["coherence_magnitude"]
        sanity = {"granger_nan_frac": float(torch.isnan(gc).float().mean()),
                  "granger_max": float(torch.nan_to_num(gc, nan=0.0).max()),
                  "coherence_mean_offdiag": float(torch.nanmean(coh)), "wilson_flagged": flags,
                  "wilson_mean_iters": mean_iters}
        del gc, coh
    del out

    # ---- end to end through the public API with host buffers -----------------------------
    x_host = torch.empty(x_dev.shape, dtype=torch.float32, pin_memory=True)
    x_host.copy_(x_dev)
    torch.cuda.synchronize()
    x_np = x_host.numpy()

    def step_e2e():
        m = sc.Multitaper(x_np, **kw)                      # H2D from pinned host memory
        cc = sc.Connectivity.from_multitaper(m)            # output="numpy": D2H of every result
        return run_measures(cc)

    e2e_steps = max(1, min(args.steps, 3))
    del c
    d2h, e2e_error = 0, None
    try:
        for _ in range(2):  # untimed: pinned staging buffers and the caching allocator reach steady state
            res = step_e2e()
            d2h = int(sum(v.nbytes for v in res.values()))
            del res
    except (RuntimeError, MemoryError) as exc:  # e.g. the host cannot pin N ranks x (inputs + results)
        e2e_error = f"{type(exc).__name__}: {str(exc)[:200]}"
    # every rank must take the same path through the barriers below
    e2e_ok = max_over_ranks(0.0 if e2e_error is None else 1.0) == 0.0
    barrier()
    e2e_each = []
    e2e_ms = e2e_value = None
    if e2e_ok:
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            t1 = time.perf_counter()
            res = step_e2e()
            del res
            e2e_each.append((time.perf_counter() - t1) * 1e3)
        barrier()
        e2e_ms = max_over_ranks((time.perf_counter() - t0) * 1e3) / e2e_steps
        e2e_value = units * world / (e2e_ms * 1e-3)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel + per-stage table ---------------------------------
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    tf32_peak = float(peaks.get("bf16_tflops", 1590.0)) / 2.0
    peak_src = "measured" if peaks else "fallback"
    T, S, K = wl["T"], wl["S"], int(2 * wl["NW"] - 1)
    tk = T * K
    alg = {
        # bytes: series read once + planar half-spectrum coefficients written once
        "mt_fft": ("hbm", 4.0 * n_win * n * T * S + 8.0 * n_win * tk * fnn * S),
        # power: a full pass over the coefficients, or -- when the CSM is computed anyway -- its real diagonal
        "power": ("hbm", (12.0 * n_win * fnn * S if ("csm" in stages or has_granger) else
                          8.0 * n_win * tk * fnn * S + 4.0 * n_win * fnn * S)),
        # tensor stage: TF32 flops actually issued = upper-triangular 128x128 tiles x 12 MMAs (Re/Im x
        # (hi*hi + hi*lo + lo*hi) x 2 products) x 2*128*128*8 per 8 observations
        "csm": ("tensor", n_win * fnn * (math.ceil(S / 128) * (math.ceil(S / 128) + 1) // 2) * math.ceil(tk / 16) * 2
                * 12 * 2.0 * 128 * 128 * 8),
        "epilogue": ("hbm", 12.0 * n_win * fnn * S * S),
        # flops: SURVEY.md 8(d): iterations x 8 complex FFTs x 5 nfft log2(nfft) per (pair, window)
        "granger": ("fp64", mean_iters * n_problems * 8 * 5.0 * nfft * math.log2(nfft)),
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
                           algorithmic_tflops=8.0 * tk * S * S * n_win * fnn / (per_step * 1e-3) / 1e12)
            else:
                row.update(bound="fp64", achieved=amount / (per_step * 1e-3) / 1e12, peak=FP64_PEAK_NOMINAL_TFLOPS,
                           unit="TFLOP/s")
            row["frac"] = row["achieved"] / row["peak"]
        stage_rows[name] = row
    # DRAM traffic per launch from the committed ncu capture (scaled to this run's windows per launch)
    traffic = {}
    try:
        tj = json.load(open(os.path.join(ROOT, "profiles", "r01_traffic.json")))
        if wl_name == "cfg4":
            for name, row in stage_rows.items():
                if name in tj["dram_bytes_per_launch"]:
                    per_window = tj["dram_bytes_per_launch"][name] / tj["windows_in_captured_launch"]
                    row["traffic"] = per_window * n_win / max(row["launches_per_step"], 1)
                    traffic[name] = row["traffic"]
    except Exception:
        pass
    dom = max(stage_rows, key=lambda k_: stage_rows[k_]["ms_per_step"]) if stage_rows else None
    roof = None
    if dom:
        r = stage_rows[dom]
        roof = {"kernel": dom, "bound": r.get("bound"), "achieved": r.get("achieved"), "peak": r.get("peak"),
                "unit": r.get("unit"), "frac": r.get("frac"), "traffic": traffic.get(dom),
                "peak_source": ("nominal B200 FP64 vector peak (not in MEASURED_PEAKS.json)" if r.get("bound") == "fp64"
                                else f"{peak_src} HBM copy bandwidth"),
                "ms_per_launch": r["ms_per_step"] / max(r["launches_per_step"], 1),
                "note": "the dominant kernel (Wilson/Granger) is FP64/FP32-SIMT + shared-memory bound, neither HBM nor "
                        "tensor bound.  achieved = ALGORITHMIC flops of the reference iteration (SURVEY.md 8d: "
                        "reference iterations x 8 complex FFTs x 5 n log2 n per pair-window) / time; the kernel "
                        "executes far fewer (4 real-packed FFTs per iteration, closed-form tail, fp32 early "
                        "iterations), so frac measures algorithm + hardware, not pipe utilisation -- ncu pipe "
                        "numbers are in profiles/",
                "hbm_algorithmic_gbs": (n_problems * fnn * (8 + 4 + 4 + 8) / (r["ms_per_step"] * 1e-3) / 1e9
                                        if r.get("bound") == "fp64" else None)}

    line = {
        "metric": METRIC, "value": value, "unit": "pair-freqs/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32 spectra/CSM, f64 Wilson", "data": "synthetic",
        "config": workload_config(wl_name, wl, world),
        "e2e": {"value": e2e_value, "unit": "pair-freqs/s", "ms_per_step": e2e_ms,
                "h2d_bytes_per_step": int(x_np.nbytes), "d2h_bytes_per_step": d2h, "steps": e2e_steps,
                "ms_each": [round(v, 1) for v in e2e_each],
                **({} if e2e_ok else {"error": e2e_error or "another rank failed to stage its host buffers"})},
        "gpu_launches": launches, "clocks": clocks, "roofline": roof, "stages": stage_rows, "sanity": sanity,
    }
    if world == 1 and not args.no_cpu_baseline and has_granger:
        cb = cpu_sample(wl)
        line["cpu_baseline"] = {"value": units / cb["est_step_seconds"], "unit": "pair-freqs/s",
                                "cores": cpu_threads(), "kind": "port", "sample": cb["sample"],
                                "host_cpus": os.cpu_count(), "measured_seconds": cb["measured_seconds"],
                                "mean_wilson_iters": cb["mean_wilson_iters"]}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="cfg4", choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    wl = WORKLOADS[args.workload]
    if args.impl == "reference":
        run_reference(args, args.workload, wl)
    else:
        run_gpu(args, args.workload, wl)


if __name__ == "__main__":
    main()
