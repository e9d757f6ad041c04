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


# ---- round 2: shard geometry that does not divide, reduce-group agreement, reduce_scatter ownership ----------------
from spectral_connectivity_b200.distributed import (GroupState, agree, owned_windows, plan_window_chunks,  # noqa: E402
                                                    scatter_ownership, shard_multitaper)
from spectral_connectivity_b200.transforms import Multitaper, sliding_window_count  # noqa: E402


def test_shard_recording_float_window_count():
    """the reference's float window count floor(N/step - n/step + 1) rounds below the integer for a slab
    of exactly (w-1)*step + n samples when step does not divide n (N=10000, n=200, step=150, world=4: rank 2 owns 16
    windows, the bare slab counts 15).  The slab is extended (< step samples) and shard_multitaper pins the count."""
    n_samples, n, step = 10000, 200, 150
    w0, w1, s0, s1 = shard_recording(n_samples, n, step, 2, 4)
    assert w1 - w0 == 16 and sliding_window_count(s1 - s0, n, step) == 16
    assert s1 - s0 < 15 * step + n + step            # never enough samples for a 17th window
    x = np.random.default_rng(1).standard_normal((n_samples, 1, 2))
    full = Multitaper(x, sampling_frequency=1000, time_window_duration=0.2, time_window_step=0.15, start_time=1.5)
    times, total = [], 0
    for world in (3, 4, 7):
        times, total = [], 0
        for r in range(world):
            m, (a, b) = shard_multitaper(x, r, world, sampling_frequency=1000, start_time=1.5, time_window_duration=0.2,
                                         time_window_step=0.15)
            assert m.n_time_windows == b - a == len(m.time)
            assert (m.n_time_windows - 1) * step + n <= m.time_series.shape[0]   # every window has its samples
            times.append(m.time)
            total += b - a
        assert total == full.n_time_windows
        assert np.allclose(np.concatenate(times), full.time, rtol=0, atol=1e-12)
    # sweep: every shard of every geometry yields its window count (through the pinned count where the last shard
    # cannot be extended)
    for n_samples in (1000, 9999, 12345):
        for n in (100, 120, 333):
            for step in (7, 50, 99, 120):
                if step > n:
                    continue
                for world in (2, 3, 8):
                    for r in range(world):
                        w0, w1, s0, s1 = shard_recording(n_samples, n, step, r, world)
                        assert s1 <= n_samples and (w1 == w0 or (w1 - w0 - 1) * step + n <= s1 - s0)


def test_chunk_plan_and_scatter_ownership():
    for n_win in (1, 5, 60, 61, 1000):
        for world in (1, 2, 3, 8):
            for per_window, budget in ((1, 1), (100, 450), (7, 10 ** 9)):
                chunks = plan_window_chunks(n_win, per_window, budget, world=world, multiple_of_world=True)
                assert chunks[0][0] == 0 and chunks[-1][1] == n_win
                assert all(a[1] == b[0] for a, b in zip(chunks, chunks[1:]))
                assert all((b - a) % world == 0 for a, b in chunks[:-1])     # only the last chunk may be ragged
                owned = [owned_windows(chunks, world, r) for r in range(world)]
                assert sorted(np.concatenate(owned).tolist()) == list(range(n_win))   # a partition of the windows
                assert max(len(o) for o in owned) - min(len(o) for o in owned) <= len(chunks)
                for a, b in chunks:
                    q = scatter_ownership(b - a, world, 0)[0]
                    assert q * world >= b - a and (q - 1) * world < b - a
    tail = plan_window_chunks(60, 1, 9, shrink_tail=True)
    assert tail[-1][1] == 60 and tail[-1][1] - tail[-1][0] <= 2


def test_group_state_validation():
    rows = np.array([[5, 100, 6, 20, 20, 960, 1], [5, 100, 6, 15, 15, 720, 1]])
    gs = GroupState.from_rows(rows, 1)
    assert (gs.world, gs.rank, gs.n_observations, gs.n_trials_tapers, gs.per_window_bin_bytes, gs.hermitian) == \
        (2, 1, 35, 35, 960, True)
    rows[1, 6] = 0
    assert GroupState.from_rows(rows, 0).hermitian is False
    rows[1, 0] = 6
    with pytest.raises(ValueError, match="time windows"):
        GroupState.from_rows(rows, 0)


def _worker_trials(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # UNEQUAL trial shards (7 trials over 2 ranks); the expectation scale must be the GLOBAL count and every rank
        # must plan the same chunks although its local footprint differs
        fs, n_samples, n, k = 100.0, 900, 100, 3
        n_trials = 7
        x = O.synthetic_series(n_samples, n_trials, 3, fs, seed=6)
        taps = O.dpss_tapers(n, 2, k, fs)
        coef_full = O.multitaper_fft(x, fs, taps, n, n, n)                      # 9 windows
        trials = np.array_split(np.arange(n_trials), world)[rank]
        part = coef_full[:, trials]
        row = [part.shape[0], n, 3, len(trials) * k, len(trials) * k, len(trials) * k * 3 * 8, 1]
        gs = agree(None, row, "cpu")
        ok_state = gs.n_observations == n_trials * k and gs.per_window_bin_bytes == 4 * k * 3 * 8
        chunks = plan_window_chunks(part.shape[0], gs.per_window_bin_bytes * n, 4 * gs.per_window_bin_bytes * n,
                                    world=world, multiple_of_world=True)
        ref = O.expected_csm(coef_full)
        mine, pieces = owned_windows(chunks, world, rank), []
        for w0, w1 in chunks:   # reduce_scatter along the window axis, emulated with all_reduce + ownership slice
            qn, lo, hi = scatter_ownership(w1 - w0, world, rank)
            partial = np.zeros((qn * world,) + ref.shape[1:], dtype=complex)
            partial[: w1 - w0] = O.cross_spectral_matrix(part[w0:w1]).sum(axis=(1, 2)) / gs.n_observations
            t = torch.from_numpy(partial)
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
            pieces.append(t.numpy()[rank * qn:rank * qn + (hi - lo)])
        got = np.concatenate(pieces)
        ok_rs = got.shape[0] == len(mine) and bool(np.allclose(got, ref[mine]))
        q.put((rank, bool(ok_state), ok_rs, [list(c) for c in chunks]))
    finally:
        dist.destroy_process_group()


def test_gloo_world2_unequal_trial_shards_reduce_scatter():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker_trials, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(r[1] and r[2] for r in results), results
    assert results[0][3] == results[1][3]         # identical chunk plans on both ranks
