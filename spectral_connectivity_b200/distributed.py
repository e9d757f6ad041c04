"""Multi-GPU orchestration: one process per GPU, windows sharded across ranks.

The reference has no distributed layer.  For the default expectation every
time window is an independent unit end to end (FFT -> CSM -> measures -> Wilson), so ranks own
disjoint, contiguous window ranges and no collective is needed on the data path (SURVEY.md
section 8e, partitioning A).  Trial sharding (partitioning B) is ``Connectivity(reduce_group=...)``:
partial sums are all-reduced before the epilogues.
"""
from __future__ import annotations

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
    concatenation over ranks along the window axis equals the single-GPU result."""
    n_win = sliding_window_count(n_samples, n_per_window, n_per_step)
    w0, w1 = window_shard(n_win, rank, world_size)
    s0, s1 = sample_range_for_windows(w0, w1, n_per_window, n_per_step)
    return w0, w1, s0, s1


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
