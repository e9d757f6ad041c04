"""Worker for the trial-sharded multi-GPU test (launched by torchrun, one rank per GPU).

Ranks hold UNEQUAL trial shards (7 trials over 2 ranks = 4 + 3) of the same windows.  Checked against the
single-GPU result on the full recording: (a) reduce_mode="all_reduce" -- every rank ends with every window;
(b) reduce_mode="reduce_scatter" -- the partial sums are reduce-scattered along the window axis (chunked, on a side
stream under the next chunk's FFT + CSM) and every rank runs the epilogues / Wilson factorisations of its own
windows only.  Covers the pairwise family, pairwise Granger, the MVAR family, canonical and global coherence."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import oracle as O  # noqa: E402
import spectral_connectivity_b200 as sc  # noqa: E402


def close(a, b, tol=2e-6):
    a, b = np.asarray(a), np.asarray(b)
    if a.shape != b.shape or not np.array_equal(np.isnan(a), np.isnan(b)):
        return False, float("inf")
    if a.size == 0:
        return True, 0.0
    err = float(np.nanmax(np.abs(a - b)) / max(np.nanmax(np.abs(b)), 1e-300))
    return err < tol, err


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
    dist.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ["LOCAL_RANK"])))
    fs = 500.0
    n_trials = 7
    x = O.synthetic_series(2500, n_trials, 6, fs, seed=17)          # 5 windows of 500 samples
    kw = dict(sampling_frequency=fs, time_halfbandwidth_product=3, time_window_duration=1.0)
    measures = ["power", "coherence_magnitude", "weighted_phase_lag_index", "phase_locking_value",
                "pairwise_spectral_granger_prediction"]
    labels = np.array([0, 0, 0, 1, 1, 1])
    trials = np.array_split(np.arange(n_trials), world)[rank]
    full_c = sc.Connectivity.from_multitaper(sc.Multitaper(x, **kw))
    full = full_c.compute(measures)
    full["canonical_coherence"] = full_c.canonical_coherence(labels)[0]
    full["global_coherence"] = full_c.global_coherence()[0]
    full["directed_transfer_function"] = full_c.directed_transfer_function()
    full["phase_slope_index"] = full_c.phase_slope_index()
    ok = True
    for mode, impl in (("all_reduce", "nccl"), ("reduce_scatter", "nccl"), ("reduce_scatter", "p2p")):
        # max_chunk_bytes=1 -> one chunk per `world` windows: several collectives, the last chunk ragged
        # impl "p2p": the fused pull-reduce + power + coherence kernel over NVLink peer memory instead of NCCL
        part = sc.Connectivity.from_multitaper(sc.Multitaper(x[:, trials], **kw), reduce_group=dist.group.WORLD,
                                               reduce_mode=mode, reduce_impl=impl, max_chunk_bytes=1)
        assert part.n_observations == n_trials * 5, part.n_observations
        got = part.compute(measures)
        got["canonical_coherence"] = part.canonical_coherence(labels)[0]
        got["global_coherence"] = part.global_coherence()[0]
        got["directed_transfer_function"] = part.directed_transfer_function()
        got["phase_slope_index"] = part.phase_slope_index()
        own = np.arange(5) if mode == "all_reduce" else part.owned_windows
        if mode == "reduce_scatter":
            counts = [torch.zeros(1, dtype=torch.int64, device="cuda") for _ in range(world)]
            dist.all_gather(counts, torch.tensor([len(own)], dtype=torch.int64, device="cuda"))
            assert sum(int(c.item()) for c in counts) == 5, counts       # every window owned exactly once
        for k, ref in full.items():
            good, err = close(got[k], ref[own], 2e-5 if k == "directed_transfer_function" else 2e-6)
            print(f"[rank {rank}] {mode:14s} {impl:4s} {k}: windows {list(own)} err {err:.2e} {'ok' if good else 'MISMATCH'}",
                  flush=True)
            ok = ok and good
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    dist.destroy_process_group()
    sys.exit(0 if int(flag.item()) == 1 else 1)


if __name__ == "__main__":
    main()
