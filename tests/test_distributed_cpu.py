"""world_size-2 gloo tests (CPU) of the host-side multi-GPU logic: window sharding with halo,
gathering of window-sharded results, and the all-reduce contract of trial-sharded partial sums."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import oracle as O
from spectral_connectivity_b200.distributed import (all_gather_windows, sample_range_for_windows, shard_recording,
                                                    shard_start_time, window_shard)


def test_window_shard_partition():
    for n in (0, 1, 7, 60, 1999):
        for world in (1, 2, 3, 8):
            spans = [window_shard(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


def test_shard_equals_full_windows():
    # overlapping windows: each shard's slab reproduces exactly its windows of the full recording
    n_samples, n, step = 1000, 120, 50
    x = np.random.default_rng(0).standard_normal((n_samples, 2, 3))
    full = O.sliding_windows(x, n, step)
    got = []
    for r in range(3):
        w0, w1, s0, s1 = shard_recording(n_samples, n, step, r, 3)
        assert (s0, s1) == sample_range_for_windows(w0, w1, n, step)
        part = O.sliding_windows(x[s0:s1], n, step)
        assert part.shape[0] == w1 - w0
        got.append(part)
        t_full = O.window_times(n_samples, 100.0, n, step)
        t_part = O.window_times(s1 - s0, 100.0, n, step, start_time=shard_start_time(0.0, s0, 100.0))
        assert np.allclose(t_part, t_full[w0:w1])
    assert np.array_equal(np.concatenate(got), full)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # (1) window-sharded coherence: each rank computes its windows with the oracle, gather == full
        fs, n_samples, n, step = 100.0, 700, 100, 100
        x = O.synthetic_series(n_samples, 4, 3, fs, seed=5)
        taps = O.dpss_tapers(n, 2, 3, fs)
        w0, w1, s0, s1 = shard_recording(n_samples, n, step, rank, world)
        coef = O.multitaper_fft(x[s0:s1], fs, taps, n, step, n)
        local = torch.from_numpy(O.coherence_magnitude(coef))
        gathered = all_gather_windows(local).numpy()
        ref = O.coherence_magnitude(O.multitaper_fft(x, fs, taps, n, step, n))
        ok1 = bool(np.allclose(gathered, ref, equal_nan=True) and gathered.shape == ref.shape)
        # (2) trial-sharded expectation: sum of per-rank partial sums scaled by 1/n_total == full mean
        coef_full = O.multitaper_fft(x, fs, taps, n, step, n)
        trials = np.array_split(np.arange(4), world)[rank]
        part = coef_full[:, trials]
        n_total = 4 * taps.shape[1]
        partial = torch.from_numpy((O.cross_spectral_matrix(part).sum(axis=(1, 2)) / n_total))
        dist.all_reduce(partial, op=dist.ReduceOp.SUM)
        ok2 = bool(np.allclose(partial.numpy(), O.expected_csm(coef_full)))
        q.put((rank, ok1, ok2))
    finally:
        dist.destroy_process_group()


def test_gloo_world2_window_and_trial_sharding():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert sorted(r[0] for r in results) == [0, 1]
    assert all(r[1] and r[2] for r in results), results
