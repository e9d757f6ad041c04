"""Multi-GPU orchestration: one process per GPU.

The reference has no distributed layer.  Two partitionings:

A. windows (default).  For the default expectation every time window is an independent unit end to end
   (FFT -> CSM -> measures -> Wilson), so ranks own disjoint, contiguous window ranges of the recording and
   NO collective is on the data path: ``shard_recording`` / ``shard_multitaper``.
B. observations (trials).  ``Connectivity(reduce_group=g, reduce_mode=...)``: every rank holds a shard of the
   trials of ALL windows and computes partial expectation sums; ``reduce_mode="reduce_scatter"`` sums them along the
   window axis so that each rank then owns the fully reduced sums of 1/world of the windows and runs the epilogues
   and the Wilson factorisations of THOSE windows only (the stage that is 84 % of the step scales with 1/world);
   ``"all_reduce"`` leaves every rank with every window.  The collective of chunk c runs on a side stream under
   the FFT + CSM of chunk c+1.  The pure host logic (what every rank must agree on, chunk plans, ownership) lives
   here so that it is testable on CPU with gloo.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

from .transforms import sliding_window_count


def window_shard(n_windows, rank, world_size):
    """Contiguous, balanced window range [w0, w1) of ``rank`` (first ranks get the remainder)."""
    if not 0 <= rank < world_size:
        raise ValueError(f"rank {rank} outside [0, {world_size})")
    base, rem = divmod(int(n_windows), int(world_size))
    w0 = rank * base + min(rank, rem)
    return w0, w0 + base + (1 if rank < rem else 0)


def sample_range_for_windows(w0, w1, n_per_window, n_per_step):
    """Samples [s0, s1) of the recording that windows [w0, w1) touch (includes the n-step halo)."""
    if w1 <= w0:
        return w0 * n_per_step, w0 * n_per_step
    return w0 * n_per_step, (w1 - 1) * n_per_step + n_per_window


def shard_recording(n_samples, n_per_window, n_per_step, rank, world_size):
    """(w0, w1, s0, s1): the windows of ``rank`` and the slab of the time axis it must hold.

    A ``Multitaper`` built on ``time_series[s0:s1]`` with the same window/step produces exactly
    windows w0..w1-1 of the full recording (same detrend, tapers and FFT per window), so the
    concatenation over ranks along the window axis equals the single-GPU result.

    The reference counts windows in FLOAT arithmetic, floor(N/step - n/step + 1) (transforms.py:1363-1365); for a
    slab of exactly (w-1)*step + n samples that expression can round just below the integer w when step does not
    divide n (e.g. N=10000, n=200, step=150: a 16-window slab counts 15).  The slab is therefore extended by up to
    step-1 samples (never enough for another window, capped at the recording) until the float count equals w1-w0;
    ``shard_multitaper`` additionally pins the count, which covers a last shard that cannot be extended."""
    n_win = sliding_window_count(n_samples, n_per_window, n_per_step)
    w0, w1 = window_shard(n_win, rank, world_size)
    s0, s1 = sample_range_for_windows(w0, w1, n_per_window, n_per_step)
    if w1 > w0:
        limit = min(n_samples, s1 + n_per_step - 1)
        while sliding_window_count(s1 - s0, n_per_window, n_per_step) < w1 - w0 and s1 < limit:
            s1 += 1
    return w0, w1, s0, s1


def shard_multitaper(time_series, rank, world_size, sampling_frequency=1000, start_time=0, **kwargs):
    """``Multitaper`` over this rank's window shard of the recording ``time_series`` (n_samples, n_trials, n_signals):
    same keywords as ``Multitaper``; ``.time`` holds the global window start times of the shard.  Returns
    (multitaper, (w0, w1))."""
    from .transforms import Multitaper
    # samples per window / step with the reference's rules (transforms.py:995-1023, 1051-1070)
    duration, step_s = kwargs.get("time_window_duration"), kwargs.get("time_window_step")
    n = kwargs.get("n_time_samples_per_window")
    if duration is not None:
        n = int(np.around(duration * sampling_frequency))
    elif n is None:
        n = time_series.shape[0]
    step = kwargs.get("n_time_samples_per_step")
    if step_s is not None:
        step = int(step_s * sampling_frequency)
    elif step is None:
        step = n
    w0, w1, s0, s1 = shard_recording(time_series.shape[0], n, step, rank, world_size)
    m = Multitaper(time_series[s0:s1], sampling_frequency=sampling_frequency,
                   start_time=shard_start_time(start_time, s0, sampling_frequency),
                   **{**kwargs, "n_time_samples_per_window": n, "n_time_samples_per_step": step})
    m._n_time_windows_override = w1 - w0
    return m, (w0, w1)


def shard_start_time(start_time, s0, sampling_frequency):
    """``start_time`` for the shard so that ``Multitaper.time`` matches the global window times."""
    return np.asarray(start_time) + s0 / sampling_frequency


def all_gather_windows(local, group=None):
    """Concatenate per-rank results along the window axis (optional; outputs may stay sharded)."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    sizes = [torch.zeros(1, dtype=torch.int64, device=local.device) for _ in range(world)]
    dist.all_gather(sizes, torch.tensor([local.shape[0]], dtype=torch.int64, device=local.device), group=group)
    sizes = [int(s.item()) for s in sizes]
    pad = max(sizes)
    buf = torch.zeros((pad,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    buf[: local.shape[0]] = local
    parts = [torch.empty_like(buf) for _ in range(world)]
    dist.all_gather(parts, buf, group=group)
    return torch.cat([p[:n] for p, n in zip(parts, sizes)], dim=0)


# ---- observation (trial) sharding: what the ranks of a reduce group must agree on --------------------------------
@dataclass(frozen=True)
class GroupState:
    """Agreed once per ``Connectivity`` from one all-gather of a row of integers per rank
    (n_windows, nfft, n_signals, local observations of the expectation, local trials x tapers,
    per-window-per-bin coefficient bytes, conjugate-symmetric flag)."""
    world: int
    rank: int
    n_observations: int          # sum over ranks of the local observation counts (shards may be unequal)
    n_trials_tapers: int         # sum over ranks of local trials x tapers (the SVD-based measures' expectation)
    per_window_bin_bytes: int    # max over ranks of trials * tapers * signals * 8
    hermitian: bool              # AND over ranks

    @classmethod
    def from_rows(cls, rows, rank, time_reduced=False):
        rows = np.asarray(rows, dtype=np.int64)
        for col, what in ((0, "time windows"), (1, "FFT samples"), (2, "signals")):
            if (rows[:, col] != rows[0, col]).any():
                raise ValueError(f"reduce_group: ranks disagree on the number of {what}: {rows[:, col].tolist()}; "
                                 "every rank must hold the same windows and signals (shard the trials)")
        return cls(world=int(rows.shape[0]), rank=int(rank), n_observations=int(rows[:, 3].sum()),
                   n_trials_tapers=int(rows[:, 4].sum()), per_window_bin_bytes=int(rows[:, 5].max()),
                   hermitian=bool(rows[:, 6].min() != 0))


def agree(group, row, device, time_reduced=False):
    """All-gather one integer row per rank (see GroupState) and return the agreed ``GroupState``."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    t = torch.tensor([int(v) for v in row], dtype=torch.int64, device=device)
    rows = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(rows, t, group=group)
    return GroupState.from_rows(torch.stack(rows).cpu().numpy(), dist.get_rank(group), time_reduced)


def plan_window_chunks(n_windows, per_window_bytes, max_chunk_bytes, world=1, multiple_of_world=False,
                       shrink_tail=False, grow_head=False):
    """Window ranges [(w0, w1), ...] of the streaming pass.  ``per_window_bytes`` must be a value every rank of a
    reduce group agrees on (GroupState), so that all ranks issue the same sequence of collectives.  With
    ``multiple_of_world`` (reduce_scatter along the window axis) full chunks hold a multiple of ``world`` windows.
    ``shrink_tail``: the last chunks shrink geometrically so that the final device->host copy, which nothing can
    overlap, is small (host output, single GPU).  ``grow_head``: the first chunks grow geometrically (1, 2, 4, ...
    windows) so that the first transform only waits for the first slab of a host->device copy that is still
    streaming in (host input)."""
    n_windows = int(n_windows)
    wc = max(1, min(n_windows, int(max_chunk_bytes) // max(int(per_window_bytes), 1)))
    if multiple_of_world and world > 1:
        wc = max(world, wc - wc % world)
    bounds, w0, head = [], 0, 1
    while w0 < n_windows:
        left = n_windows - w0
        size = min(wc, left)
        if grow_head and head < size:
            size, head = head, head * 2
        elif shrink_tail and left <= wc and left > 1:
            size = (left + 1) // 2
        bounds.append((w0, w0 + size))
        w0 += size
    return bounds


def scatter_ownership(n_items, world, rank):
    """reduce_scatter of a chunk of ``n_items`` windows padded to ``q * world``: (q, lo, hi) -- every rank receives
    q rows, rows [lo, hi) of the chunk are this rank's valid ones (hi - lo <= q; trailing ranks may own none)."""
    q = -(-int(n_items) // int(world))
    lo = min(rank * q, n_items)
    hi = min((rank + 1) * q, n_items)
    return q, lo, hi


def owned_windows(chunks, world, rank):
    """Global window indices a rank owns after reduce_scatter over every chunk, in output order."""
    out = []
    for w0, w1 in chunks:
        _, lo, hi = scatter_ownership(w1 - w0, world, rank)
        out.extend(range(w0 + lo, w0 + hi))
    return np.asarray(out, dtype=np.int64)


def bind_to_local_cpus(device_index):
    """Pin the calling process to the CPU cores (NUMA node) closest to GPU ``device_index`` (NVML's ideal CPU
    affinity), so that pinned host buffers allocated afterwards are node-local and N ranks do not all stage
    through NUMA node 0.  Returns the CPU list, or None when NVML / the affinity call is unavailable."""
    import os
    try:
        import pynvml
        pynvml.nvmlInit()
        handle = pynvml.nvmlDeviceGetHandleByIndex(int(device_index))
        n_cpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(handle, (n_cpu + 63) // 64)
        cpus = [w * 64 + b for w, word in enumerate(words) for b in range(64) if (word >> b) & 1 and w * 64 + b < n_cpu]
        allowed = sorted(set(cpus) & os.sched_getaffinity(0))
        if allowed:
            os.sched_setaffinity(0, allowed)
            return allowed
    except Exception:
        return None
    return None
