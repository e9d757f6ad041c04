"""``Multitaper``: drop-in for the reference's spectral transform class, B200 backend.

Mirrors ``spectral_connectivity.transforms.Multitaper`` (transforms.py:442-1171): same
constructor keywords, properties, validation and warnings.  The index arithmetic
(samples per window/step, window count, FFT length, frequencies, window start times)
is computed on the host with the reference's formulas because it must be bit exact.
The transform itself -- window gather, detrend, taper product, FFT (:1147-1171, :1300-1405)
-- is one fused CUDA kernel behind ``sc_mt_fft`` (csrc/mtfft.cu).
"""
from __future__ import annotations

import ctypes
import warnings
from logging import getLogger

import numpy as np
import torch
from scipy.fft import fftfreq, next_fast_len

from . import _lib
from ._dpss import TAPER_MULTIPLIER, dpss_windows, make_tapers  # noqa: F401

logger = getLogger(__name__)

# (window, trial, taper) axes reduced by each expectation type (connectivity.py:67-75)
EXPECTATION_AXES = {
    "time": (0,),
    "trials": (1,),
    "tapers": (2,),
    "time_trials": (0, 1),
    "time_tapers": (0, 2),
    "trials_tapers": (1, 2),
    "time_trials_tapers": (0, 1, 2),
}


def _device():
    if not torch.cuda.is_available():
        raise RuntimeError(
            "spectral_connectivity_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback.")
    return torch.device("cuda", torch.cuda.current_device())


def expectation_map(shape_wtk, expectation_type):
    """Linear maps (b_w,b_t,b_k,r_w,r_t,r_k), kept dims, B, R for the planar layout."""
    reduced = EXPECTATION_AXES[expectation_type]
    bm, rm = [0, 0, 0], [0, 0, 0]
    nb = nr = 1
    for ax in (2, 1, 0):  # fastest axis last
        if ax in reduced:
            rm[ax] = nr
            nr *= shape_wtk[ax]
        else:
            bm[ax] = nb
            nb *= shape_wtk[ax]
    kept = tuple(shape_wtk[ax] for ax in range(3) if ax not in reduced)
    return bm + rm, kept, nb, nr


def twiddles(nfft, dtype, device):
    """exp(-2 pi i q / nfft) computed in float64 on the host."""
    q = np.arange(nfft, dtype=np.float64)
    w = np.exp(-2j * np.pi * q / nfft)
    return torch.from_numpy(w.astype(np.complex128 if dtype == torch.complex128 else np.complex64)).to(device)


def sliding_window_count(n_samples, window, step):
    """Number of windows, float arithmetic then floor (transforms.py:1363-1365)."""
    return int(np.floor((n_samples / step) - (window / step) + 1))


class Multitaper:
    """Multitaper spectral transform of (n_time_samples, n_trials, n_signals) data.

    Parameters are those of the reference (transforms.py:574-589).  ``time_series`` may be
    a NumPy array or a torch tensor (host or device); it is held on the GPU as float32.
    """

    def __init__(self, time_series, sampling_frequency=1000, time_halfbandwidth_product=3,
                 detrend_type="constant", time_window_duration=None, time_window_step=None,
                 n_tapers=None, tapers=None, start_time=0, n_fft_samples=None,
                 n_time_samples_per_window=None, n_time_samples_per_step=None, is_low_bias=True):
        shape = tuple(time_series.shape)
        ndim = len(shape)
        if ndim != 3:
            msg = ("Expected 3D array with shape (n_time_samples, n_trials, n_signals), "
                   f"but got {ndim}D array with shape {shape}.\n")
            if ndim == 1:
                msg += "For a single time series use time_series[:, np.newaxis, np.newaxis]."
            elif ndim == 2:
                msg += ("For 2D data say what the second axis is: time_series[:, np.newaxis, :] for "
                        "(n_time, n_signals) or time_series[:, :, np.newaxis] for (n_time, n_trials).")
            else:
                msg += f"Arrays with {ndim} dimensions are not supported."
            raise ValueError(msg)
        if sampling_frequency <= 0:
            raise ValueError(f"sampling_frequency must be positive, got {sampling_frequency}.")
        if time_halfbandwidth_product < 1:
            raise ValueError(
                f"time_halfbandwidth_product must be at least 1, got {time_halfbandwidth_product}.")
        if time_halfbandwidth_product > 10:
            warnings.warn(f"time_halfbandwidth_product = {time_halfbandwidth_product} is unusually large; "
                          "values above 10 apply very heavy spectral smoothing.", UserWarning, stacklevel=2)
        if time_window_duration is not None and time_window_duration <= 0:
            raise ValueError(f"time_window_duration must be positive, got {time_window_duration}.")
        if time_window_step is not None and time_window_step <= 0:
            raise ValueError(f"time_window_step must be positive, got {time_window_step}.")
        if (time_window_step is not None and time_window_duration is not None
                and time_window_step > time_window_duration):
            warnings.warn(f"time_window_step ({time_window_step}s) is larger than time_window_duration "
                          f"({time_window_duration}s): this creates gaps between analysis windows.",
                          UserWarning, stacklevel=2)
        if detrend_type not in _lib.DETREND:
            raise ValueError(f"Invalid trend type '{detrend_type}' is not supported. "
                             "Valid options are 'linear'/'l', 'constant'/'c' or None.")
        n_time, _, n_signals = shape
        if n_time < n_signals:
            warnings.warn(f"Your time series has only {n_time} time points but {n_signals} signals. "
                          "This seems unusual and your data may be transposed. Expected shape: "
                          "(n_time_samples, n_trials, n_signals).", UserWarning, stacklevel=2)

        # Without a CUDA device the object can still be built (host-side properties, validation);
        # any transform then fails loudly in _device() -- there is no CPU compute path.
        dev = _device() if torch.cuda.is_available() else torch.device("cpu")
        if isinstance(time_series, torch.Tensor):
            ts = time_series
        else:
            ts = torch.from_numpy(np.ascontiguousarray(time_series))
        if not ts.is_floating_point():
            ts = ts.to(torch.float64)
        # host -> device copy (transforms.py:590 `xp.asarray`), then float32 on the device
        self._h2d_events = None
        self._finite_flag = None
        self._host_src = None
        if dev.type == "cuda" and not ts.is_cuda and ts.numel() * ts.element_size() >= self._ASYNC_H2D_BYTES:
            # Large host arrays stream to the device in slabs of the time axis on a side stream; the
            # transform of a window chunk only waits for the slabs it reads, so the copy overlaps the
            # compute.  The NaN/Inf scan (transforms.py:754-774) runs on the device behind each slab and its
            # warning is raised at the first synchronisation point (fft() / Connectivity.compute()).
            self._stream_to_device(ts, dev)
        else:
            self.time_series = ts.to(dev, non_blocking=True).to(torch.float32).contiguous()
            if not self._all_finite(self.time_series):
                warnings.warn(self._NONFINITE_MSG, UserWarning, stacklevel=2)

        self.sampling_frequency = sampling_frequency
        self.time_halfbandwidth_product = time_halfbandwidth_product
        self.detrend_type = detrend_type
        self._time_window_duration = time_window_duration
        self._time_window_step = time_window_step
        self.is_low_bias = is_low_bias
        self.start_time = np.asarray(start_time)
        self._n_fft_samples = n_fft_samples
        self._tapers = None if tapers is None else np.asarray(tapers, dtype=np.float64)
        self._n_tapers = n_tapers
        self._n_time_samples_per_window = n_time_samples_per_window
        self._n_samples_per_time_step = n_time_samples_per_step
        self._tapers_dev = None
        self._tw = {}

    @staticmethod
    def _all_finite(x):
        """transforms.py:754: one device pass (sc_nonfinite_flag) instead of isfinite().all()'s five."""
        if not x.is_cuda:
            return bool(torch.isfinite(x).all())
        flag = torch.zeros(1, dtype=torch.int32, device=x.device)
        _lib.check(_lib.load().sc_nonfinite_flag(_lib.ptr(x), x.numel(), _lib.ptr(flag), _lib.stream_ptr()),
                   "sc_nonfinite_flag")
        return int(flag.item()) == 0

    _ASYNC_H2D_BYTES = 256 << 20
    _NONFINITE_MSG = ("Input time_series contains NaN or infinite values. This will produce "
                      "invalid spectral estimates.")

    def _stream_to_device(self, ts, dev, n_slabs=16):
        ts = ts.contiguous()
        self._host_src = ts  # keep the host buffer alive until the copies have run
        n_rows = ts.shape[0]
        self.time_series = torch.empty(tuple(ts.shape), dtype=torch.float32, device=dev)
        self._finite_flag = torch.zeros(1, dtype=torch.int32, device=dev)  # sc_nonfinite_flag ORs 1 into it
        lib = _lib.load()
        self._h2d_events = []
        copy_stream = _lib.side_stream(dev, "h2d")
        copy_stream.wait_stream(torch.cuda.current_stream(dev))
        rows = max(1, -(-n_rows // n_slabs))
        # the first slabs grow geometrically (1/16 of a regular slab, doubling): the first window chunk of a streamed
        # pass then waits ~1 ms for its samples instead of a whole regular slab (4.9 ms at config 4)
        bounds, r0, size = [], 0, max(1, rows // 16)
        while r0 < n_rows:
            r1 = min(n_rows, r0 + size)
            bounds.append((r0, r1))
            r0, size = r1, min(rows, 2 * size)
        with torch.cuda.stream(copy_stream):
            for r0, r1 in bounds:
                dst = self.time_series[r0:r1]
                if ts.dtype == torch.float32:
                    dst.copy_(ts[r0:r1], non_blocking=True)
                else:
                    dst.copy_(ts[r0:r1].to(dev, non_blocking=True))
                _lib.check(lib.sc_nonfinite_flag(_lib.ptr(dst), dst.numel(), _lib.ptr(self._finite_flag),
                                                 ctypes.c_void_p(copy_stream.cuda_stream)), "sc_nonfinite_flag")
                ev = torch.cuda.Event()
                ev.record(copy_stream)
                self._h2d_events.append((r1, ev))
        self.time_series.record_stream(copy_stream)

    def _wait_rows(self, row_end):
        """Make the current stream wait for the host->device slabs covering rows [0, row_end)."""
        if not self._h2d_events:
            return
        cur = torch.cuda.current_stream(self.time_series.device)
        for r1, ev in self._h2d_events:
            cur.wait_event(ev)
            if r1 >= row_end:
                break

    def _check_finite_deferred(self):
        """Raise the deferred non-finite warning of the streamed copy (synchronises once)."""
        if self._finite_flag is not None:
            flag, self._finite_flag = self._finite_flag, None
            self._wait_rows(self.time_series.shape[0])
            if int(flag.item()) != 0:
                warnings.warn(self._NONFINITE_MSG, UserWarning, stacklevel=3)
            self._h2d_events = None
            self._host_src = None

    def __repr__(self):
        return ("Multitaper("
                f"sampling_frequency={self.sampling_frequency!r}, "
                f"time_halfbandwidth_product={self.time_halfbandwidth_product!r}, "
                f"time_window_duration={self.time_window_duration!r}, "
                f"time_window_step={self.time_window_step!r}, "
                f"detrend_type={self.detrend_type!r}, "
                f"start_time={self.start_time}, "
                f"n_tapers={self.n_tapers}"
                ")")

    # ---- host-side properties (reference formulas, bit exact) ---------------
    @property
    def tapers(self):
        """(n_time_samples_per_window, n_tapers) float64 tapers (transforms.py:925-945)."""
        if self._tapers is None:
            self._tapers = make_tapers(self.n_time_samples_per_window, self.sampling_frequency,
                                       self.time_halfbandwidth_product, self.n_tapers,
                                       is_low_bias=self.is_low_bias)
        return self._tapers

    @property
    def time_window_duration(self):
        if self._time_window_duration is None:
            self._time_window_duration = self.n_time_samples_per_window / self.sampling_frequency
        return self._time_window_duration

    @property
    def time_window_step(self):
        if self._time_window_step is None:
            self._time_window_step = self.n_time_samples_per_step / self.sampling_frequency
        return self._time_window_step

    @property
    def n_tapers(self):
        """``floor(2 NW - 1)`` unless given (transforms.py:979-993)."""
        if self._n_tapers is None:
            return int(np.floor(TAPER_MULTIPLIER * self.time_halfbandwidth_product - 1))
        return self._n_tapers

    @property
    def n_time_samples_per_window(self):
        """``int(around(duration * fs))`` (transforms.py:995-1023)."""
        if self._n_time_samples_per_window is None and self._time_window_duration is None:
            self._n_time_samples_per_window = self.time_series.shape[0]
        elif self._time_window_duration is not None:
            self._n_time_samples_per_window = int(
                np.around(self.time_window_duration * self.sampling_frequency))
        return self._n_time_samples_per_window

    @property
    def n_fft_samples(self):
        if self._n_fft_samples is None:
            self._n_fft_samples = next_fast_len(self.n_time_samples_per_window)
        return self._n_fft_samples

    @property
    def frequencies(self):
        return fftfreq(self.n_fft_samples, 1.0 / self.sampling_frequency)

    @property
    def n_time_samples_per_step(self):
        """``int(step * fs)`` -- truncation (transforms.py:1051-1070)."""
        if self._n_samples_per_time_step is None and self._time_window_step is None:
            self._n_samples_per_time_step = self.n_time_samples_per_window
        elif self._time_window_step is not None:
            self._n_samples_per_time_step = int(self.time_window_step * self.sampling_frequency)
        return self._n_samples_per_time_step

    _n_time_windows_override = None  # distributed.shard_multitaper pins the shard's window count

    @property
    def n_time_windows(self):
        if self._n_time_windows_override is not None:
            return self._n_time_windows_override
        return sliding_window_count(self.time_series.shape[0], self.n_time_samples_per_window,
                                    self.n_time_samples_per_step)

    @property
    def time(self):
        """Start time of each window (transforms.py:1072-1091)."""
        starts = np.arange(self.n_time_windows) * self.n_time_samples_per_step
        original_time = np.arange(0, self.time_series.shape[0]) / self.sampling_frequency
        return self.start_time + original_time[starts]

    @property
    def n_signals(self):
        return self.time_series.shape[-1]

    @property
    def n_trials(self):
        return self.time_series.shape[1]

    @property
    def frequency_resolution(self):
        return TAPER_MULTIPLIER * self.time_halfbandwidth_product / self.time_window_duration

    @property
    def nyquist_frequency(self):
        return self.sampling_frequency / 2

    # ---- device path ------------------------------------------------------------
    def _device_tapers(self):
        if self._tapers_dev is None:
            tp = np.ascontiguousarray(np.asarray(self.tapers, dtype=np.float64).T)  # (K, n)
            if tp.shape[1] != self.n_time_samples_per_window:
                raise ValueError(f"tapers must have shape (n_time_samples_per_window, n_tapers) = "
                                 f"({self.n_time_samples_per_window}, K), got {self.tapers.shape}")
            self._tapers_dev = torch.from_numpy(tp.astype(np.float32)).to(self.time_series.device)
        return self._tapers_dev

    def _twiddle(self, dtype=torch.complex64):
        key = (self.n_fft_samples, dtype)
        if key not in self._tw:
            self._tw[key] = twiddles(self.n_fft_samples, dtype, self.time_series.device)
        return self._tw[key]

    @property
    def n_tapers_effective(self):
        """Taper count after the low-bias filter (what the coefficient array really holds)."""
        return int(np.asarray(self.tapers).shape[1])

    def _transform(self, out, layout, n_freq_out, w0, n_win, w_out0=0, mapping=None, n_reduce=0):
        """Enqueue sc_mt_fft for windows [w0, w0+n_win) on the current stream."""
        lib = _lib.load()
        if not self.time_series.is_cuda:
            self.time_series = self.time_series.to(_device())
        n_samples, n_trials, n_signals = self.time_series.shape
        taps = self._device_tapers()
        n, step, nfft = self.n_time_samples_per_window, self.n_time_samples_per_step, self.n_fft_samples
        ws_bytes = lib.sc_mt_fft_workspace_bytes(n, nfft)
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=out.device) if ws_bytes else None
        tw = self._twiddle()
        self._wait_rows((w0 + n_win - 1) * step + n)
        with _lib.timed("mt_fft"):
            rc = lib.sc_mt_fft(_lib.ptr(self.time_series), n_samples, n_trials, n_signals, _lib.ptr(taps), n,
                               taps.shape[0], step, w0, n_win, w_out0, nfft, _lib.DETREND[self.detrend_type],
                               1.0 / float(self.sampling_frequency), _lib.ptr(tw), layout,
                               n_freq_out, _lib.map6(mapping) if mapping is not None else None, n_reduce,
                               _lib.ptr(out), _lib.ptr(ws), ws_bytes, _lib.stream_ptr())
            _lib.check(rc, "sc_mt_fft")
        return out

    def fft(self):
        """Fourier coefficients, shape (n_time_windows, n_trials, n_tapers, n_fft_samples,
        n_signals), two-sided, as a complex64 CUDA tensor (the analogue of the reference
        returning an ``xp`` array, transforms.py:1147-1171)."""
        logger.info(self)
        n_win = self.n_time_windows
        shape = (n_win, self.n_trials, self.n_tapers_effective, self.n_fft_samples, self.n_signals)
        out = torch.empty(shape, dtype=torch.complex64, device=_device())
        if out.numel():
            self._transform(out, _lib.LAYOUT_REFERENCE, self.n_fft_samples, 0, n_win)
        self._check_finite_deferred()
        # Tag the result so that the reference's two-step idiom ``Connectivity(fourier_coefficients=m.fft(), ...)``
        # can take the fused real-series path (half spectrum, no re-layout) as long as the tensor is unmodified.
        out._sc_source = (self, out._version)
        return out
