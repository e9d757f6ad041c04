"""The live reference (Eden-Kramer-Lab/spectral_connectivity, unmodified) as the timed CPU arm.

``baseline/_ref/spectral_connectivity`` is an install of the reference package (``pip install
--target`` needs the ``hatchling`` build backend, which the offline wheelhouse does not hold, so
``__graft_entry__.build()`` installs the pure-Python package by copying its directory -- exactly the files pip
would have laid down; the directory is git-ignored, so no reference source enters the repository).  The package
``__init__`` imports xarray (absent in this image); the three hot-path modules only need numpy + scipy and are
imported through the shim of SURVEY.md's appendix.
"""
import os
import sys
import time
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, "_ref")

_mods = None


def available():
    return os.path.isfile(os.path.join(REF_DIR, "spectral_connectivity", "connectivity.py"))


def load():
    """(transforms, connectivity, minimum_phase_decomposition) modules of the installed reference."""
    global _mods
    if _mods is None:
        if not available():
            raise ImportError(f"the reference is not installed under {REF_DIR}; run __graft_entry__.build() in the "
                              "build container (where /root/reference exists)")
        pkg = types.ModuleType("spectral_connectivity")
        pkg.__path__ = [os.path.join(REF_DIR, "spectral_connectivity")]
        sys.modules["spectral_connectivity"] = pkg
        import spectral_connectivity.connectivity as C
        import spectral_connectivity.minimum_phase_decomposition as M
        import spectral_connectivity.transforms as T
        _mods = (T, C, M)
    return _mods


def cpu_threads():
    try:
        from threadpoolctl import threadpool_info
        return max([p.get("num_threads", 1) for p in threadpool_info()] + [1])
    except Exception:
        return 1


def sample_step(x, fs, nw, duration, measures, group_labels=None):
    """One pass of the reference's public API over the recording ``x`` (n_samples, n_trials, n_signals), float64:
    Multitaper -> Connectivity.from_multitaper -> each measure.  Returns (seconds, {measure: seconds}, outputs)."""
    T, C, _ = load()
    times, out = {}, {}
    t_all = time.perf_counter()
    m = T.Multitaper(x, sampling_frequency=fs, time_halfbandwidth_product=nw, time_window_duration=duration)
    c = C.Connectivity.from_multitaper(m)           # runs m.fft() (connectivity.py:366-400)
    times["multitaper_fft"] = time.perf_counter() - t_all
    for name in measures:
        t0 = time.perf_counter()
        if name == "canonical_coherence":
            out[name] = c.canonical_coherence(group_labels)[0]
        elif name == "expectation_cross_spectral_matrix":
            out[name] = c._expectation_cross_spectral_matrix()
        else:
            out[name] = getattr(c, name)()
        times[name] = time.perf_counter() - t0
    return time.perf_counter() - t_all, times, out


def bounded_sample(synth, wl, measures, n_channels, seed=0):
    """The bench workload ``wl`` cut down to a size the reference finishes in seconds: ONE window (all trials, all
    tapers) of ``n_channels`` channels; every pairwise measure is evaluated on all pairs of those channels.
    pair-freqs/s is an intensive quantity of the reference (its coherence cost is proportional to S^2 -- the
    un-averaged cross-spectral tensor -- and its Granger cost to the pair count), so the rate measured on the sample
    is the rate of the full workload up to the (1 - 1/S) pair-count factor, which favours the sample."""
    n = wl["N"] if wl["duration"] is None else int(np.around(wl["duration"] * wl["fs"]))
    s = min(n_channels, wl["S"])
    x = synth(n, wl["T"], s, wl["fs"], seed=20261017 + seed)
    labels = np.arange(s) // max(1, s // 2) if "canonical_coherence" in measures else None
    secs, times, out = sample_step(x, wl["fs"], wl["NW"], wl["duration"], measures, labels)
    first = next(iter(out.values()))
    fnn = first.shape[-3] if first.ndim >= 3 else 0
    units = fnn * s * s                              # W = 1 window
    return dict(seconds=secs, stage_seconds=times, units=units, n_channels=s, n_windows=1, fnn=fnn)
