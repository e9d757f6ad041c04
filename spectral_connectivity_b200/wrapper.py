"""``multitaper_connectivity``: the reference's high-level entry point (wrapper.py:137-287) on the B200 backend.

Same arguments.  The reference builds a fresh ``Connectivity`` -- and therefore re-runs ``Multitaper.fft()`` and the
cross-spectral matrix -- for EVERY requested method (wrapper.py:265-287 -> connectivity_to_xarray ->
Connectivity.from_multitaper); here all methods share ONE transform and ONE expectation pass: the pairwise family
and pairwise Granger go through a single ``Connectivity.compute`` call, and the measures built on the full
cross-spectral matrix (MVAR family, canonical / global coherence, phase slope index) reuse the cached matrix
(``share_csm``).  Results are labelled ``xarray`` objects when xarray is importable (it is not part of this image) and
a ``LabeledResults`` dict of NumPy arrays with the same coordinates otherwise.
"""
from __future__ import annotations

from logging import getLogger

import numpy as np

from .connectivity import MEASURES, Connectivity
from .transforms import Multitaper

logger = getLogger(__name__)

# what the reference's `method=None` expands to (every public measure its xarray interface supports,
# wrapper.py:228-261); the two Granger variants it lists raise NotImplementedError there and here
DEFAULT_METHODS = (
    "blockwise_spectral_granger_prediction", "coherence_magnitude", "coherence_phase", "coherency",
    "conditional_spectral_granger_prediction", "debiased_squared_phase_lag_index",
    "debiased_squared_weighted_phase_lag_index", "imaginary_coherence", "pairwise_phase_consistency",
    "pairwise_spectral_granger_prediction", "phase_lag_index", "phase_locking_value", "power",
    "weighted_phase_lag_index")
MVAR_METHODS = ("directed_transfer_function", "directed_coherence", "partial_directed_coherence",
                "generalized_partial_directed_coherence", "direct_directed_transfer_function")


class LabeledResults(dict):
    """{method: ndarray} plus ``coords`` (time, frequency, source, target) and ``dims`` per method."""

    def __init__(self, coords):
        super().__init__()
        self.coords = coords
        self.dims = {}


def _dims_of(name, arr):
    if name == "power":
        return ("time", "frequency", "source")
    if name == "phase_slope_index":
        return ("time", "source", "target")
    if arr.ndim == 2:
        return ("time", "frequency")
    return ("time", "frequency", "source", "target")


def multitaper_connectivity(time_series, sampling_frequency, time_window_duration=None, method=None,
                            signal_names=None, squeeze=False, connectivity_kwargs=None, **kwargs):
    """Multitaper transform + the requested connectivity measure(s) in one shared pass (wrapper.py:137-287).

    ``time_series``: (n_times, n_trials, n_signals) or (n_times, n_signals).  ``method``: a name, a list of names,
    or None for every measure the reference's xarray interface supports.  ``squeeze`` with two signals returns the
    [first, last] entry of pairwise measures only.  ``connectivity_kwargs`` go to the measure methods
    (e.g. ``group_labels`` is not supported by the reference's interface either); other keyword arguments go to
    ``Multitaper``.  A single method name returns one array (DataArray), otherwise a mapping (Dataset)."""
    connectivity_kwargs = dict(connectivity_kwargs or {})
    single = isinstance(method, str)
    methods = list(DEFAULT_METHODS) if method is None else ([method] if single else list(method))
    ts = np.asarray(time_series) if not hasattr(time_series, "shape") else time_series
    if len(ts.shape) == 2:  # (n_times, n_signals) -> one trial (prepare_time_series, transforms.py:1174-1297)
        ts = ts[:, None, :]
    m = Multitaper(ts, sampling_frequency=sampling_frequency, time_window_duration=time_window_duration, **kwargs)
    c = Connectivity.from_multitaper(m, share_csm=True)
    n_sig = m.n_signals
    names = list(signal_names) if signal_names is not None else list(range(n_sig))
    coords = {"time": np.atleast_1d(m.time), "frequency": c.frequencies, "source": names, "target": names}
    out = LabeledResults(coords)
    fused = [name for name in methods if name in MEASURES]
    if fused:
        fused_kwargs = {k: v for k, v in connectivity_kwargs.items()
                        if k in ("tolerance", "max_iterations", "tail_extrapolation", "mixed_precision", "pairs")}
        out.update(c.compute(fused, **fused_kwargs))
    for name in methods:
        if name in out:
            continue
        try:
            fn = getattr(c, name)
            res = fn(**connectivity_kwargs) if name not in MVAR_METHODS else fn()
            out[name] = res[0] if isinstance(res, tuple) else res
        except NotImplementedError as exc:
            if len(methods) == 1:
                raise exc
            logger.warning(f"{name} is not implemented")
    for name in list(out):
        arr = out[name]
        if squeeze and n_sig == 2 and arr.ndim >= 2 and arr.shape[-1] == 2 and arr.shape[-2] == 2:
            arr = arr[..., 0, -1]
        out[name] = arr
        out.dims[name] = _dims_of(name, arr)
    try:
        import xarray as xr
    except ImportError:
        return out[methods[0]] if single and methods[0] in out else out
    ds = xr.Dataset()
    for name, arr in out.items():
        dims = out.dims[name]
        ds[name] = xr.DataArray(arr, dims=dims, coords={d: coords[d] for d in dims}, name=name)
    return ds[methods[0]] if single and methods[0] in ds else ds
