"""Worker for the trial-sharded multi-GPU test (launched by torchrun, one rank per GPU)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import oracle as O  # noqa: E402
import spectral_connectivity_b200 as sc  # noqa: E402


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
    dist.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ["LOCAL_RANK"])))
    fs = 500.0
    x = O.synthetic_series(2000, 8, 6, fs, seed=17)
    kw = dict(sampling_frequency=fs, time_halfbandwidth_product=3, time_window_duration=1.0)
    measures = ["power", "coherence_magnitude", "weighted_phase_lag_index", "phase_locking_value",
                "pairwise_spectral_granger_prediction"]
    trials = np.array_split(np.arange(8), world)[rank]
    part = sc.Connectivity.from_multitaper(sc.Multitaper(x[:, trials], **kw), reduce_group=dist.group.WORLD)
    assert part.n_observations == 8 * 5
    got = part.compute(measures)
    ok = True
    if rank == 0:
        full = sc.Connectivity.from_multitaper(sc.Multitaper(x, **kw)).compute(measures)
        for k in measures:
            a, b = got[k], full[k]
            same_nan = np.array_equal(np.isnan(a), np.isnan(b))
            err = np.nanmax(np.abs(a - b)) / np.nanmax(np.abs(b))
            print(f"{k}: nan-mask {same_nan} err {err:.2e}")
            ok = ok and same_nan and err < 2e-6
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    dist.destroy_process_group()
    sys.exit(0 if int(flag.item()) == 1 else 1)


if __name__ == "__main__":
    main()
