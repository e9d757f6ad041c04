"""``Connectivity``: drop-in for the reference's connectivity-measure class, B200 backend.

Mirrors ``spectral_connectivity.connectivity.Connectivity`` (connectivity.py:163-1650) for
the hot path: cross-spectral matrix + expectation (:441-526), power / coherency / coherence
magnitude & phase / imaginary coherence (:612-743), phase-locking value, pairwise phase
consistency, phase-lag index and its weighted / debiased variants (:897-1159), and pairwise
spectral Granger prediction through Wilson factorisation (:1161-1213, :2282-2340).

Differences from the reference that a user can see, all deliberate:
  * the un-averaged (W,T,K,F,S,S) tensor is never materialised, so ``blocks`` is accepted
    and ignored; the device pipeline computes the spectra / CSM in fp32 (complex64) and Wilson / Granger in
    mixed fp32 / fp64; ``dtype=np.complex128`` keeps every Wilson iteration in fp64 (default ``None`` =
    the device default; the reference's default is complex128);
  * one streaming pass over window chunks can produce several measures (``compute``);
  * results are float32 (complex64) NumPy arrays, or CUDA tensors with ``output="torch"``.
"""
from __future__ import annotations

import warnings
from itertools import combinations
from logging import getLogger

import numpy as np
import torch

from . import _lib
from .transforms import EXPECTATION_AXES, expectation_map, twiddles

logger = getLogger(__name__)

EXPECTATION = tuple(EXPECTATION_AXES)
TIKHONOV_REGULARIZATION_FACTOR = 1e-12  # connectivity.py:79

# measure -> (needs, epilogue code or None, complex output?)
_PAIRWISE = {
    "coherency": ("csm", _lib.M_COHERENCY, True),
    "coherence_magnitude": ("csm", _lib.M_COHERENCE_MAG, False),
    "coherence_phase": ("csm", _lib.M_COHERENCE_PHASE, False),
    "imaginary_coherence": ("csm", _lib.M_IMAG_COHERENCE, False),
    "phase_locking_value": ("plv", _lib.M_PLV, False),
    "pairwise_phase_consistency": ("plv", _lib.M_PPC, False),
    "phase_lag_index": ("pli", _lib.M_PLI, False),
    "weighted_phase_lag_index": ("pli", _lib.M_WPLI, False),
    "debiased_squared_phase_lag_index": ("pli", _lib.M_DPLI2, False),
    "debiased_squared_weighted_phase_lag_index": ("pli", _lib.M_DWPLI2, False),
}
# Every expectation type is factorised; the reference couples the Wilson freeze / convergence test across the leading
# axis of the CSM (minimum_phase_decomposition.py:290, 310-315: several kept tapers of one window, or -- for
# 'time_trials_tapers' -- single frequency bins), the device treats every kept index as its own problem: results differ
# by O(tolerance) = 1e-8 absolute (tests/test_gpu_round2.py::test_granger_and_dtf_every_expectation_type, 1e-5 vs
# the live reference).
_GRANGER_OK = tuple(EXPECTATION_AXES)
_SYMM_SLOTS = {}  # (group name, device) -> peer-mapped partial-sum buffers of reduce_impl="p2p"
# real measures that are symmetric in (i, j): may be returned as packed upper triangles (compute(packed=...))
SYMMETRIC_MEASURES = ("coherence_magnitude", "phase_locking_value", "pairwise_phase_consistency",
                      "debiased_squared_phase_lag_index", "debiased_squared_weighted_phase_lag_index")
MEASURES = ("power", "expectation_cross_spectral_matrix", "_phase_locking_value",
            "pairwise_spectral_granger_prediction") + tuple(_PAIRWISE)


def _expectation_error(expectation_type):
    words = set(str(expectation_type).split("_"))
    msg = (f"Invalid expectation_type '{expectation_type}' is not supported.\n"
           "This parameter controls which dimensions to average over when computing the "
           "cross-spectral matrix.\n")
    if words.issubset({"time", "trials", "tapers"}):
        for key in EXPECTATION:
            if set(key.split("_")) == words:
                msg += f"\nDid you mean '{key}'? (The words must be in a specific order)\n"
                break
    msg += "\nValid options are:\n" + "".join(f"  - '{k}'\n" for k in sorted(EXPECTATION))
    msg += "\nMost common: 'trials_tapers' (average over both trials and tapers)"
    return msg


class Connectivity:
    """Connectivity measures from multitaper Fourier coefficients.

    Parameters follow the reference (connectivity.py:277-285).  ``fourier_coefficients`` is
    a 5-D complex array (n_time_windows, n_trials, n_tapers, n_fft_samples, n_signals),
    NumPy or torch.  Extra keyword-only options: ``output`` ("numpy" | "torch"),
    ``max_chunk_bytes`` (planar-coefficient budget per streamed window chunk),
    ``reduce_group`` (a torch.distributed group whose ranks hold disjoint observation
    shards -- e.g. trials -- of the same windows) with ``reduce_mode`` ("all_reduce": every rank ends with every
    window; "reduce_scatter": partial sums are summed along the window axis and every rank computes the measures
    of its own windows, ``owned_windows``), and ``share_csm`` (cache the expected cross-spectral matrix across
    measures).
    """

    def __init__(self, fourier_coefficients, expectation_type="trials_tapers", frequencies=None,
                 time=None, blocks=None, dtype=None, *, output="numpy",
                 max_chunk_bytes=None, reduce_group=None, reduce_mode="all_reduce", reduce_impl="nccl",
                 share_csm=False, _multitaper=None):
        src = getattr(fourier_coefficients, "_sc_source", None)
        if (_multitaper is None and src is not None and isinstance(fourier_coefficients, torch.Tensor)
                and fourier_coefficients._version == src[1]):
            # unmodified output of Multitaper.fft(): same as from_multitaper (connectivity.py:366-400)
            _multitaper = src[0]
            self._coef_given = fourier_coefficients
        self._mt = _multitaper
        if _multitaper is None:
            if fourier_coefficients.ndim != 5:
                raise ValueError(
                    f"fourier_coefficients must be 5-dimensional, got {fourier_coefficients.ndim}D array.\n"
                    "Expected shape: (n_time_windows, n_trials, n_tapers, n_fft_samples, n_signals)\n"
                    f"Got shape: {tuple(fourier_coefficients.shape)}\n\n"
                    "If you have time series data, use the Multitaper class to transform it:\n"
                    "  m = Multitaper(time_series, sampling_frequency=your_fs, ...)\n"
                    "  fourier_coefficients = m.fft()")
        if expectation_type not in EXPECTATION_AXES:
            raise ValueError(_expectation_error(expectation_type))
        if output not in ("numpy", "torch"):
            raise ValueError("output must be 'numpy' or 'torch'")
        if reduce_mode not in ("all_reduce", "reduce_scatter"):
            raise ValueError("reduce_mode must be 'all_reduce' or 'reduce_scatter'")
        if reduce_impl not in ("nccl", "p2p"):
            raise ValueError("reduce_impl must be 'nccl' or 'p2p'")
        if not torch.cuda.is_available():
            raise RuntimeError("spectral_connectivity_b200 needs a CUDA device; there is no CPU fallback.")
        self._device = torch.device("cuda", torch.cuda.current_device())
        if _multitaper is None:
            coef = fourier_coefficients
            if not isinstance(coef, torch.Tensor):
                coef = torch.from_numpy(np.ascontiguousarray(coef))
            if not coef.is_complex():
                coef = coef.to(torch.float32).to(torch.complex64)
            self._coef = coef.to(self._device, non_blocking=True).to(torch.complex64).contiguous()
            if not bool(torch.isfinite(torch.view_as_real(self._coef)).all()):
                warnings.warn("fourier_coefficients contains NaN or Inf values. Check the input time "
                              "series, the windowing parameters and any preprocessing.", UserWarning,
                              stacklevel=2)
            self._shape = tuple(self._coef.shape)
            self._hermitian = self._is_conjugate_symmetric(self._coef)
        else:
            m = _multitaper
            self._coef = getattr(self, "_coef_given", None)
            self._shape = (m.n_time_windows, m.n_trials, m.n_tapers_effective, m.n_fft_samples, m.n_signals)
            self._hermitian = True  # real time series: X(-f) = conj X(f)
        self.expectation_type = expectation_type
        self._frequencies = frequencies
        self._blocks = blocks
        # ``dtype`` (connectivity.py:277-285 lets the caller pick the cross-spectral precision): None (default) and
        # complex64 = the device default, float32 spectra / CSM with the mixed-precision Wilson factorisation;
        # complex128 = keep every Wilson / Granger iteration in float64 (mixed_precision=False becomes the default of
        # the Granger methods).  Spectra and the CSM stay float32 either way (the tensor-core contraction).
        if dtype is not None and np.dtype(dtype) not in (np.dtype(np.complex64), np.dtype(np.complex128)):
            raise ValueError(f"dtype must be complex64 or complex128, got {dtype}")
        self._dtype = dtype
        self._fp64_wilson = dtype is not None and np.dtype(dtype) == np.dtype(np.complex128)
        self._output = output
        # streamed window chunk: 4 GiB of planar coefficients; under a reduce group 8 GiB of partial sums per
        # collective (with reduce_scatter a rank keeps 1/world of a chunk, and the full-matrix Wilson factorisation
        # downstream wants several windows per batch -- BASELINE config 5: 252 MB per window)
        if max_chunk_bytes is None:
            max_chunk_bytes = (8 << 30) if reduce_group is not None else (4 << 30)
        self._max_chunk_bytes = int(max_chunk_bytes)
        self._reduce_group = reduce_group
        self._reduce_mode = reduce_mode
        # reduce_impl="p2p" (reduce_scatter mode, cross-spectral sums): instead of NCCL, one kernel per rank pulls the
        # rows it owns from every peer's partial sums over NVLink (torch symmetric memory), adds them in rank order and
        # fuses the power extraction and one coherence-family epilogue (csrc/peer_reduce.cu)
        self._reduce_impl = reduce_impl
        self._symm = None
        # share_csm: keep the expected cross-spectral matrix of the first measure that needs it (if it fits in
        # _CSM_CACHE_BYTES) so that later measures -- the MVAR family, canonical / global coherence, the phase slope
        # index, further compute() calls -- skip the FFT + CSM pass (the reference recomputes both per measure,
        # wrapper.py:265-287)
        self._share_csm = bool(share_csm)
        self._csm_cache = {}
        self.time = time if not isinstance(time, torch.Tensor) else time.cpu().numpy()
        self.owned_windows = None  # reduce_scatter mode: global indices of the windows this rank's results hold
        self._group_state = None
        if reduce_group is not None:
            self._agree_with_group()
        self.last_granger_iterations = None
        self.last_granger_flags = None
        self.last_granger_executed = None

    def _agree_with_group(self):
        """Ranks of ``reduce_group`` hold disjoint OBSERVATION shards (e.g. trials) of the same windows.  Everything
        the collectives depend on must be identical on every rank, so it is agreed once here: the global observation
        counts (shards may be unequal: array_split of 8 trials over 3 ranks), the window-chunk size (from the
        largest per-window footprint in the group) and the conjugate-symmetry flag (AND over ranks)."""
        import torch.distributed as dist

        from .distributed import agree
        n_win, n_trials, n_tapers, nfft, n_sig = self._shape
        local_obs = int(np.prod([self._shape[a] for a in EXPECTATION_AXES[self.expectation_type]]))
        row = [n_win, nfft, n_sig, local_obs, n_trials * n_tapers, n_trials * n_tapers * n_sig * 8,
               1 if self._hermitian else 0]
        world = dist.get_world_size(self._reduce_group)
        self._group_state = agree(self._reduce_group, row, self._device,
                                  0 in EXPECTATION_AXES[self.expectation_type])
        self._hermitian = self._group_state.hermitian
        if self._reduce_mode == "reduce_scatter" and world > 1:
            from .distributed import owned_windows
            self.owned_windows = owned_windows(self._chunk_bounds(nfft, "trials_tapers"), world, self._group_state.rank)

    @staticmethod
    def _is_conjugate_symmetric(coef, rtol=1e-6):
        """True when X(-f) = conj X(f) along the frequency axis (coefficients of real time series): then the
        non-negative bins carry everything and the half-spectrum Wilson/Granger kernels apply."""
        nfft = coef.shape[3]
        if nfft < 2 or coef.numel() == 0:
            return False
        scale = float(coef.abs().max())
        if not np.isfinite(scale) or scale == 0.0:
            return False
        if float(coef[:, :, :, 0].imag.abs().max()) > rtol * scale:
            return False
        mirrored = torch.flip(coef[:, :, :, 1:], dims=(3,)).conj()
        return float((coef[:, :, :, 1:] - mirrored).abs().max()) <= rtol * scale

    @classmethod
    def from_multitaper(cls, multitaper_instance, expectation_type="trials_tapers", blocks=None,
                        dtype=None, **kwargs):
        """Fused path (connectivity.py:366-400): the coefficients are produced window chunk by
        window chunk straight into the layout the CSM kernels read; ``m.fft()`` is not
        materialised."""
        return cls(None, expectation_type=expectation_type, time=multitaper_instance.time,
                   frequencies=multitaper_instance.frequencies, blocks=blocks, dtype=dtype,
                   _multitaper=multitaper_instance, **kwargs)

    # ---- bookkeeping ---------------------------------------------------------------
    @property
    def fourier_coefficients(self):
        if self._coef is None:
            self._coef = self._mt.fft()
        return self._coef

    @property
    def frequencies(self):
        """Non-negative frequencies with the Nyquist sign fix (connectivity.py:402-424)."""
        if self._frequencies is None:
            return None
        freqs = np.asarray(self._frequencies)
        freqs = freqs[: len(freqs) // 2 + 1]
        if len(freqs) > 0 and freqs[-1] < 0:
            freqs = freqs.copy()
            freqs[-1] = abs(freqs[-1])
        return freqs

    @property
    def all_frequencies(self):
        return None if self._frequencies is None else np.asarray(self._frequencies)

    @property
    def n_observations(self):
        """connectivity.py:594-610."""
        if self._group_state is not None:
            return self._group_state.n_observations  # sum over the ranks' (possibly unequal) shards
        return int(np.prod([self._shape[a] for a in EXPECTATION_AXES[self.expectation_type]]))

    def _finish(self, t):
        if self._output == "torch":
            return t
        # device -> host through a pinned staging tensor (torch caches pinned allocations)
        host = torch.empty(t.shape, dtype=t.dtype, pin_memory=True)
        host.copy_(t, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        return host.numpy()

    # ---- streaming engine ----------------------------------------------------------
    def _scatter(self, expectation_type=None):
        """True when results are window-sharded over the reduce group (reduce_scatter along the window axis)."""
        gs = self._group_state
        et = expectation_type or self.expectation_type
        return (gs is not None and gs.world > 1 and self._reduce_mode == "reduce_scatter"
                and 0 not in EXPECTATION_AXES[et])

    def _chunk_bounds(self, n_freq, expectation_type=None):
        from .distributed import plan_window_chunks
        et = expectation_type or self.expectation_type
        n_win, n_trials, n_tapers, nfft, n_sig = self._shape
        if 0 in EXPECTATION_AXES[et]:
            return [(0, n_win)]
        gs = self._group_state
        if gs is None or gs.world == 1:
            per_window = n_trials * n_tapers * n_freq * n_sig * 8
            streaming_in = self._mt is not None and bool(getattr(self._mt, "_h2d_events", None))
            return plan_window_chunks(n_win, per_window, self._max_chunk_bytes, shrink_tail=self._output == "numpy",
                                      grow_head=streaming_in)
        # reduce group: every rank must derive the same bounds -> only group-agreed numbers; the chunk is also the
        # granularity of the collective, so the reduced sums (CSM: S^2 * 8 bytes per window and bin) count too.
        # Sized for all nfft bins whatever ``n_freq`` is streamed, so that the window ownership under
        # reduce_scatter is one fixed property of the object (``owned_windows``), not of the measure.
        per_window = max(gs.per_window_bin_bytes * nfft, nfft * n_sig * n_sig * 8)
        return plan_window_chunks(n_win, per_window, self._max_chunk_bytes, world=gs.world,
                                  multiple_of_world=self._scatter(et))

    def _owned(self, n_freq, expectation_type=None):
        """reduce_scatter mode: how many windows this rank's results hold (else None = all)."""
        if not self._scatter(expectation_type):
            return None
        return len(self.owned_windows)

    def _kept_dims(self, expectation_type=None, n_local_windows=None):
        et = expectation_type or self.expectation_type
        dims = []
        for a in range(3):
            if a in EXPECTATION_AXES[et]:
                continue
            dims.append(self._shape[a] if a != 0 or n_local_windows is None else n_local_windows)
        return tuple(dims)

    def _coefficients_chunk(self, w0, w1, n_freq, expectation_type):
        """Planar coefficients [nb][n_freq][2][R][S] of windows [w0, w1) for the expectation type -> (xp, nb, R)."""
        lib = _lib.load()
        n_win, n_trials, n_tapers, nfft, n_sig = self._shape
        mapping, kept, nb, nr = expectation_map((w1 - w0, n_trials, n_tapers), expectation_type)
        xp = torch.empty((nb, n_freq, 2, nr, n_sig), dtype=torch.float32, device=self._device)
        if self._mt is not None:
            self._mt._transform(xp, _lib.LAYOUT_PLANAR, n_freq, w0, w1 - w0, 0, mapping, nr)
        else:
            rc = lib.sc_repack_coefficients(_lib.ptr(self._coef[w0:w1]), w1 - w0, n_trials, n_tapers,
                                            nfft, n_sig, n_freq, _lib.map6(mapping), nr, _lib.ptr(xp),
                                            _lib.stream_ptr())
            _lib.check(rc, "sc_repack_coefficients")
        return xp, nb, nr

    def _symmetric_slots(self, nbytes):
        """Three peer-mapped buffers (torch symmetric memory over the reduce group) of at least ``nbytes`` each: the
        partial cross-spectral sums of three consecutive chunks.  Triple buffering makes ONE device-side barrier per
        chunk sufficient: a slot is rewritten by the CSM kernel of chunk c+3, which this rank only enqueues after it
        has waited for its own reduce of chunk c+1, i.e. after barrier c+1, which every peer enters only after its
        reduce of chunk c -- the last reader of the slot -- has finished."""
        import torch.distributed._symmetric_memory as symm_mem
        # allocation + rendezvous is a collective handshake (IPC handle exchange, ~100 ms): cached per process and
        # group, grown when a larger chunk comes along (every rank sees the same sizes, so all ranks grow together)
        key = (self._reduce_group.group_name, self._device.index)
        slots = _SYMM_SLOTS.get(key)
        if slots is None or slots["nbytes"] < nbytes:
            bufs = [symm_mem.empty(int(nbytes), dtype=torch.uint8, device=self._device) for _ in range(3)]
            handles = [symm_mem.rendezvous(b, self._reduce_group) for b in bufs]
            slots = _SYMM_SLOTS[key] = dict(nbytes=int(nbytes), bufs=bufs, handles=handles, used=0)
        self._symm = slots
        return slots

    def _reduced(self, n_freq, kinds, expectation_type=None, scale=None, fuse=None):
        """Stream window chunks and yield the expectation sums of each: dict(o0, o1, w0, w1, csm, power, plv, pli)
        -- tensors hold rows [o0, o1) of THIS rank's output batch axis (``kinds`` selects which are produced;
        "power" comes from the CSM diagonal when the CSM is produced anyway).

        With a reduce group the partial sums of chunk c are summed over the ranks on a side stream
        (reduce_scatter along the window axis, or all_reduce) while the main stream already runs the FFT + CSM of
        chunk c+1; the consumer's epilogues / Wilson factorisations of chunk c then wait for the collective only."""
        import torch.distributed as dist

        from .distributed import scatter_ownership
        lib = _lib.load()
        et = expectation_type or self.expectation_type
        n_win, n_trials, n_tapers, nfft, n_sig = self._shape
        dev = self._device
        gs = self._group_state
        reduce = gs is not None and gs.world > 1
        scatter = self._scatter(et)
        time_kept = 0 not in EXPECTATION_AXES[et]
        if scale is None:
            scale = 1.0 / self.n_observations
        comm = _lib.side_stream(dev, "comm") if reduce else None
        st = _lib.stream_ptr()
        kinds = set(kinds)
        # fused peer-memory reduce: reduce_scatter of the cross-spectral sums only (the headline path), even S
        p2p = (scatter and self._reduce_impl == "p2p" and kinds <= {"csm", "power"} and "csm" in kinds
               and n_sig % 2 == 0)
        symm = None
        if p2p:
            bounds_ = self._chunk_bounds(n_freq, et)
            max_rows = max(scatter_ownership(w1 - w0, gs.world, gs.rank)[0] * gs.world for w0, w1 in bounds_)
            bw_ = 1
            for a in (1, 2):
                if a not in EXPECTATION_AXES[et]:
                    bw_ *= self._shape[a]
            symm = self._symmetric_slots(max_rows * bw_ * n_freq * n_sig * n_sig * 8)
        cached = self._csm_cache.get((n_freq, et)) if kinds <= {"csm", "power"} and "csm" in kinds else None
        if cached is not None:      # share_csm: an earlier measure already produced these sums
            rows_per_item = max(1, (1 << 30) // max(n_freq * n_sig * n_sig * 8, 1))
            for o0 in range(0, cached.shape[0], rows_per_item):
                o1 = min(cached.shape[0], o0 + rows_per_item)
                item = dict(o0=o0, o1=o1, w0=o0, w1=o1, csm=cached[o0:o1])
                if "power" in kinds:
                    power = torch.empty((o1 - o0, n_freq, n_sig), dtype=torch.float32, device=dev)
                    with _lib.timed("power"):
                        _lib.check(lib.sc_power_from_csm(_lib.ptr(item["csm"]), (o1 - o0) * n_freq, n_sig,
                                                         _lib.ptr(power), st), "sc_power_from_csm")
                    item["power"] = power
                yield item
            return
        want_power = "power" in kinds
        sums = [k for k in ("csm", "plv", "pli") if k in kinds]
        if want_power and "csm" not in kinds:
            sums.append("power")

        def partial_sums(w0, w1):
            xp, nb, nr = self._coefficients_chunk(w0, w1, n_freq, et)
            bw = nb // (w1 - w0) if time_kept else nb        # batch rows per window
            rows = nb
            if scatter:
                q, lo, hi = scatter_ownership(w1 - w0, gs.world, gs.rank)
                rows = q * gs.world * bw                     # padded so that every rank receives q windows
            item = dict(w0=w0, w1=w1, nb=nb, bw=bw, rows=rows)

            def alloc(shape_tail, dtype, lead=None):
                t = torch.empty(((rows,) if lead is None else (lead, rows)) + shape_tail, dtype=dtype, device=dev)
                if rows > nb:
                    (t[nb:] if lead is None else t[:, nb:]).zero_()
                return t
            for kind in sums:
                if kind == "csm" and p2p:
                    slot = symm["used"] % 3
                    symm["used"] += 1
                    nfl = rows * n_freq * n_sig * n_sig * 2
                    t = torch.view_as_complex(symm["bufs"][slot].view(torch.float32)[:nfl].view(rows, n_freq, n_sig, n_sig, 2))
                    if rows > nb:
                        t[nb:].zero_()
                    item["slot"] = slot
                    with _lib.timed("csm"):
                        _lib.check(lib.sc_csm(_lib.ptr(xp), nb, n_freq, nr, n_sig, scale, _lib.CSM_CROSS, _lib.ptr(t), st),
                                   "sc_csm")
                elif kind == "csm":
                    t = alloc((n_freq, n_sig, n_sig), torch.complex64)
                    with _lib.timed("csm"):
                        _lib.check(lib.sc_csm(_lib.ptr(xp), nb, n_freq, nr, n_sig, scale, _lib.CSM_CROSS, _lib.ptr(t), st),
                                   "sc_csm")
                elif kind == "plv":
                    t = alloc((n_freq, n_sig, n_sig), torch.complex64)
                    with _lib.timed("plv"):
                        _lib.check(lib.sc_csm(_lib.ptr(xp), nb, n_freq, nr, n_sig, scale, _lib.CSM_PLV, _lib.ptr(t), st),
                                   "sc_csm[plv]")
                elif kind == "pli":
                    if rows > nb:   # the kernel writes 4 planes of nb rows: compute compact, then pad
                        c = torch.empty((4, nb, n_freq, n_sig, n_sig), dtype=torch.float32, device=dev)
                    else:
                        c = torch.empty((4, rows, n_freq, n_sig, n_sig), dtype=torch.float32, device=dev)
                    with _lib.timed("pli"):
                        _lib.check(lib.sc_csm(_lib.ptr(xp), nb, n_freq, nr, n_sig, scale, _lib.CSM_PLI, _lib.ptr(c), st),
                                   "sc_csm[pli]")
                    if rows > nb:
                        t = alloc((n_freq, n_sig, n_sig), torch.float32, lead=4)
                        t[:, :nb] = c
                    else:
                        t = c
                else:
                    t = alloc((n_freq, n_sig), torch.float32)
                    with _lib.timed("power"):
                        _lib.check(lib.sc_power(_lib.ptr(xp), nb, n_freq, nr, n_sig, scale, _lib.ptr(t), st), "sc_power")
                item[kind] = t
            return item

        enq = dict(o=0)

        def enqueue_peer_reduce(item):
            """barrier (all partial sums of the chunk are written) + the fused pull-reduce kernel, on the side stream"""
            import ctypes
            q, lo, hi = scatter_ownership(item["w1"] - item["w0"], gs.world, gs.rank)
            bw = item["bw"]
            valid = (hi - lo) * bw
            o0 = enq["o"]
            enq["o"] = o0 + valid
            csm_r = torch.empty((valid, n_freq, n_sig, n_sig), dtype=torch.complex64, device=dev)
            power_r = torch.empty((valid, n_freq, n_sig), dtype=torch.float32, device=dev) if want_power else None
            code, dest = (-1, None) if fuse is None else (fuse[0], fuse[1](o0, o0 + valid))
            hdl = symm["handles"][item["slot"]]
            ptrs = (ctypes.c_void_p * gs.world)(*[int(p_) for p_ in hdl.buffer_ptrs])
            ev = torch.cuda.Event()
            ev.record()
            comm.wait_event(ev)
            with torch.cuda.stream(comm), _lib.timed("collective"):
                hdl.barrier(channel=0)
                _lib.check(lib.sc_peer_reduce_csm(ptrs, gs.world, lo * bw * n_freq, valid * n_freq, n_sig,
                                                  _lib.ptr(csm_r), _lib.ptr(power_r), code, _lib.ptr(dest),
                                                  _lib.stream_ptr()), "sc_peer_reduce_csm")
                done = torch.cuda.Event()
                done.record(comm)
            item.update(done=done, csm=csm_r, power=power_r, p2p=True, fused=dest is not None, valid=valid, o0=o0)

        def enqueue_collective(item):
            if p2p:
                return enqueue_peer_reduce(item)
            ev = torch.cuda.Event()
            ev.record()
            comm.wait_event(ev)
            with torch.cuda.stream(comm), _lib.timed("collective"):
                for kind in sums:
                    t = item[kind]
                    if not scatter:
                        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self._reduce_group)
                        continue
                    own = item["rows"] // gs.world
                    if kind == "pli":
                        r = torch.empty((4, own) + tuple(t.shape[2:]), dtype=t.dtype, device=dev)
                        for a in range(4):
                            dist.reduce_scatter_tensor(r[a], t[a], op=dist.ReduceOp.SUM, group=self._reduce_group)
                    else:
                        r = torch.empty((own,) + tuple(t.shape[1:]), dtype=t.dtype, device=dev)
                        if t.is_complex():   # NCCL has no complex types: sum the interleaved (re, im) floats
                            dist.reduce_scatter_tensor(torch.view_as_real(r), torch.view_as_real(t),
                                                       op=dist.ReduceOp.SUM, group=self._reduce_group)
                        else:
                            dist.reduce_scatter_tensor(r, t, op=dist.ReduceOp.SUM, group=self._reduce_group)
                    item[kind + "_partial"] = t      # keep the send buffer alive until the consumer has run
                    item[kind] = r
                done = torch.cuda.Event()
                done.record(comm)
            item["done"] = done

        state = dict(o=0)

        def finish(item):
            if "done" in item:
                torch.cuda.current_stream().wait_event(item["done"])
            if item.get("p2p"):     # reduced matrix, power (and the fused measure) came out of the peer-reduce kernel
                q, lo, hi = scatter_ownership(item["w1"] - item["w0"], gs.world, gs.rank)
                item["w0"], item["w1"] = item["w0"] + lo, item["w0"] + hi
                item["o1"] = item["o0"] + item["valid"]
                state["o"] = item["o1"]
                return item
            nb, bw = item["nb"], item["bw"]
            if scatter:
                q, lo, hi = scatter_ownership(item["w1"] - item["w0"], gs.world, gs.rank)
                valid = (hi - lo) * bw
                item["w0"], item["w1"] = item["w0"] + lo, item["w0"] + hi
                o0 = state["o"]
            else:
                valid = nb
                o0 = item["w0"] * bw if time_kept else 0
            for kind in sums:
                t = item[kind]
                item[kind] = t[:, :valid] if kind == "pli" else t[:valid]
            item["o0"], item["o1"] = o0, o0 + valid
            state["o"] = o0 + valid
            if want_power and "csm" in kinds:
                power = torch.empty((valid, n_freq, n_sig), dtype=torch.float32, device=dev)
                if valid:
                    with _lib.timed("power"):  # the diagonal of the CSM: no second pass over the coefficients
                        _lib.check(lib.sc_power_from_csm(_lib.ptr(item["csm"]), valid * n_freq, n_sig, _lib.ptr(power), st),
                                   "sc_power_from_csm")
                item["power"] = power
            return item

        pending = None
        for w0, w1 in self._chunk_bounds(n_freq, et):
            item = partial_sums(w0, w1)
            if reduce:
                enqueue_collective(item)
                if pending is not None:
                    yield finish(pending)
                pending = item
            else:
                yield finish(item)
        if pending is not None:
            yield finish(pending)

    _CSM_CACHE_BYTES = 24 << 30

    def _local_csm(self, n_freq, expectation_type=None, scale=None, private=False):
        """The expected cross-spectral matrix of every window this rank holds, c64 [rows][n_freq][S][S]; cached on the
        object with ``share_csm`` (``private=True`` returns a buffer the caller may overwrite)."""
        et = expectation_type or self.expectation_type
        n_win, n_trials, n_tapers, nfft, n_sig = self._shape
        key = (n_freq, et)
        hit = self._csm_cache.get(key)
        if hit is None and self._hermitian and n_freq == nfft and (nfft // 2 + 1, et) in self._csm_cache:
            half = self._csm_cache[(nfft // 2 + 1, et)]      # real series: C(-f) = conj C(f)
            mirror = half[:, 1:nfft - nfft // 2].flip(1).conj()
            return torch.cat([half, mirror], dim=1)
        if hit is not None:
            return hit.clone() if private else hit
        n_local = self._owned(n_freq, et)
        kept = self._kept_dims(et, n_local_windows=n_local)
        rows = int(np.prod(kept)) if kept else 1
        csm = torch.empty((rows, n_freq, n_sig, n_sig), dtype=torch.complex64, device=self._device)
        for item in self._reduced(n_freq, ["csm"], et, scale=scale):
            csm[item["o0"]:item["o1"]] = item["csm"]
        if self._share_csm and csm.numel() * 8 <= self._CSM_CACHE_BYTES:
            self._csm_cache[key] = csm
            return csm.clone() if private else csm
        return csm

    def compute(self, measures, pairs=None, tolerance=1e-8, max_iterations=60, tail_extrapolation=True,
                mixed_precision=None, out=None, packed=()):
        """Compute several measures in ONE streaming pass over window chunks.

        ``measures``: iterable of names from ``MEASURES``.  Returns {name: array}.  Per chunk
        the Fourier coefficients, power and cross-spectral matrix are produced once and
        shared by all requested measures (the reference recomputes them per measure).

        ``tail_extrapolation`` (Granger, real-series path only): sum the geometric tail of the
        reference's Wilson iteration in closed form instead of iterating through it -- same
        stopping iterate, same iteration count, results equal to ~1e-7 relative.
        ``False`` runs every iteration like the reference.  ``mixed_precision`` (same path): the
        first Wilson iterations (non-constant part of the update still > 4e-4 of |G|) run in fp32, the rest in
        fp64; moves results by ~5e-7 relative.  ``False`` keeps the factorisation in fp64 throughout
        (``dtype=np.complex128`` at construction does not change this default: see ``Connectivity``).

        ``out`` (output="numpy" only): {name: pinned host array from ``pinned_empty``} to receive results, so that
        a pipeline calling ``compute`` repeatedly does not allocate (and page-lock) gigabytes per call.

        ``packed``: names from ``SYMMETRIC_MEASURES`` to return as packed upper triangles, shape
        (..., n_frequencies, S (S + 1) / 2) with element (i, j >= i) at ``i S - i (i - 1) / 2 + j - i`` (diagonal
        included; ``unpack_upper`` restores the full array).  An opt-in format without a reference counterpart: it
        halves the device -> host bytes of those results, which is what bounds an end-to-end pass once the kernels are
        faster than the host link.

        With ``reduce_mode="reduce_scatter"`` the results hold this rank's windows only (``owned_windows``)."""
        lib = _lib.load()
        measures = list(measures)
        packed = tuple(packed or ())
        for name in packed:
            if name not in SYMMETRIC_MEASURES or name not in measures:
                raise ValueError(f"packed: '{name}' is not a requested symmetric measure "
                                 f"(symmetric measures: {', '.join(SYMMETRIC_MEASURES)})")
        if mixed_precision is None:
            mixed_precision = not self._fp64_wilson
        for name in measures:
            if name not in MEASURES:
                raise ValueError(f"unknown measure '{name}'")
        n_win, n_trials, n_tapers, nfft, n_sig = self._shape
        fnn = nfft // 2 + 1
        want_granger = "pairwise_spectral_granger_prediction" in measures
        if want_granger and self.expectation_type not in _GRANGER_OK:
            raise NotImplementedError(
                f"pairwise_spectral_granger_prediction with expectation_type='{self.expectation_type}' couples "
                "the Wilson convergence test across a kept axis (minimum_phase_decomposition.py:290, "
                ":313-315); only " + ", ".join(_GRANGER_OK) + " are supported.")
        two_sided = want_granger and not self._hermitian
        n_freq = nfft if two_sided else fnn
        kept = self._kept_dims(n_local_windows=self._owned(n_freq))
        n_batch = int(np.prod(kept)) if kept else 1
        dev = self._device
        needs = set()
        for name in measures:
            if name in _PAIRWISE:
                needs.add(_PAIRWISE[name][0])
        if "_phase_locking_value" in measures:
            needs.add("plv")
        if want_granger or "expectation_cross_spectral_matrix" in measures:
            needs.add("csm")
        if want_granger or "power" in measures or any(
                m in measures for m in ("coherency", "coherence_magnitude", "coherence_phase", "imaginary_coherence")):
            needs.add("power")

        res = {}
        for name in measures:
            if name == "power":
                res[name] = torch.empty((n_batch, n_freq, n_sig), dtype=torch.float32, device=dev)
            elif name == "pairwise_spectral_granger_prediction":
                res[name] = torch.full((n_batch, fnn, n_sig, n_sig), float("nan"), dtype=torch.float32, device=dev)
            elif name in ("expectation_cross_spectral_matrix", "_phase_locking_value", "coherency"):
                res[name] = torch.empty((n_batch, n_freq, n_sig, n_sig), dtype=torch.complex64, device=dev)
            else:
                res[name] = torch.empty((n_batch, n_freq, n_sig, n_sig), dtype=torch.float32, device=dev)

        tri = n_sig * (n_sig + 1) // 2
        res_packed = {name: torch.empty((n_batch, n_freq, tri), dtype=torch.float32, device=dev) for name in packed}

        pair_t = None
        n_pairs = n_sig * (n_sig - 1) // 2
        if want_granger:
            if pairs is not None:
                pair_np = np.ascontiguousarray(np.asarray(pairs, dtype=np.int32).reshape(-1, 2))
                if pair_np.size and (pair_np.min() < 0 or pair_np.max() >= n_sig):
                    raise ValueError("pairs contain signal indices outside [0, n_signals)")
                if (pair_np[:, 0] == pair_np[:, 1]).any():
                    raise ValueError("pairs must reference two different signals")
                pair_t = torch.from_numpy(pair_np).to(dev)
                n_pairs = pair_np.shape[0]
            ws_bytes = lib.sc_wilson_workspace_bytes(nfft)
            gr_ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
            tw128 = twiddles(nfft, torch.complex128, dev)
            tw64 = twiddles(nfft, torch.complex64, dev)
            it_all = torch.zeros((n_pairs, n_batch), dtype=torch.int32, device=dev)
            fl_all = torch.zeros((n_pairs, n_batch), dtype=torch.int32, device=dev)
            # executed work (fp32 iterations, fp64 iterations, closed-form tail steps, problems): roofline accounting
            exec_cnt = torch.zeros(4, dtype=torch.int64, device=dev)

        # output="numpy": stream each finished chunk to pinned host memory on a side stream while the
        # next chunk computes
        host, copy_stream = {}, None
        if self._output == "numpy":
            copy_stream = _lib.side_stream(dev, "d2h")
            for name, t in res.items():
                shp = (t.shape[0], fnn) + (tuple(t.shape[2:]) if name not in res_packed else (tri,))
                given = None if out is None else out.get(name)
                if given is not None:
                    g = given if isinstance(given, torch.Tensor) else torch.from_numpy(given)
                    if g.numel() != int(np.prod(shp)) or g.dtype != t.dtype or not g.is_contiguous():
                        raise ValueError(f"out['{name}'] must be a contiguous {t.dtype} array with {int(np.prod(shp))} "
                                         f"elements (shape {kept + shp[1:]})")
                    host[name] = g.reshape(shp)
                else:
                    host[name] = torch.empty(shp, dtype=t.dtype, pin_memory=True)
        elif out is not None:
            raise ValueError("out= is for output='numpy'; with output='torch' results stay on the device")

        def offload(name, b0, b1):
            if copy_stream is None or b1 <= b0:
                return
            ev = torch.cuda.Event()
            ev.record()
            copy_stream.wait_event(ev)
            with torch.cuda.stream(copy_stream):
                host[name][b0:b1].copy_(res_packed.get(name, res[name])[b0:b1, :fnn], non_blocking=True)

        st = _lib.stream_ptr()
        # reduce_impl="p2p": the first coherence-family measure is produced by the peer-reduce kernel itself
        fuse, fused_name = None, None
        if self._reduce_impl == "p2p" and self._scatter() and n_freq == fnn:
            for name in measures:
                if name in _PAIRWISE and _PAIRWISE[name][0] == "csm":
                    fused_name = name
                    fuse = (_PAIRWISE[name][1], lambda o0, o1, _n=name: res[_n][o0:o1])
                    break
        for item in self._reduced(n_freq, needs, fuse=fuse):
            b0, b1 = item["o0"], item["o1"]
            nb = b1 - b0
            if nb == 0:
                continue
            csm, power, plv, pli = item.get("csm"), item.get("power"), item.get("plv"), item.get("pli")
            if "expectation_cross_spectral_matrix" in res:
                res["expectation_cross_spectral_matrix"][b0:b1] = csm
                offload("expectation_cross_spectral_matrix", b0, b1)
            if "power" in res:
                res["power"][b0:b1] = power
                offload("power", b0, b1)
            if "_phase_locking_value" in res:
                res["_phase_locking_value"][b0:b1] = plv
                offload("_phase_locking_value", b0, b1)
            for name in measures:
                if name not in _PAIRWISE:
                    continue
                src_kind, code, _ = _PAIRWISE[name]
                if name == fused_name and item.get("fused"):
                    if name in res_packed:
                        _lib.check(lib.sc_pack_upper(_lib.ptr(res[name][b0:b1]), nb * n_freq, n_sig,
                                                     _lib.ptr(res_packed[name][b0:b1]), st), "sc_pack_upper")
                    offload(name, b0, b1)
                    continue
                src = {"csm": csm, "plv": plv, "pli": pli}[src_kind]
                if src_kind == "pli" and not src.is_contiguous():
                    src = src.contiguous()
                dst = res[name][b0:b1]
                with _lib.timed("epilogue"):
                    _lib.check(lib.sc_pairwise_epilogue(code, _lib.ptr(src),
                                                        _lib.ptr(power) if src_kind == "csm" else None, nb, n_freq,
                                                        n_sig, float(self.n_observations), _lib.ptr(dst), st),
                               f"sc_pairwise_epilogue[{name}]")
                    if name in res_packed:
                        _lib.check(lib.sc_pack_upper(_lib.ptr(dst), nb * n_freq, n_sig, _lib.ptr(res_packed[name][b0:b1]),
                                                     st), "sc_pack_upper")
                offload(name, b0, b1)
            if want_granger:
                # With host output the chunk's windows go through the Wilson kernel in sub-launches of a few windows,
                # each followed by its device->host copy: the Granger results are half of the bytes that leave the
                # device and would otherwise all become available at the very end of the chunk (the copy stream idles,
                # then lags: 18 ms of copies were still queued after the last kernel of a config-4 pass).
                sub = nb if copy_stream is None else max(1, min(nb, self._GRANGER_SUB_WINDOWS))
                for k0 in range(0, nb, sub):
                    k1 = min(nb, k0 + sub)
                    it_c = torch.zeros((n_pairs, k1 - k0), dtype=torch.int32, device=dev)
                    fl_c = torch.zeros((n_pairs, k1 - k0), dtype=torch.int32, device=dev)
                    dst = res["pairwise_spectral_granger_prediction"][b0 + k0:b0 + k1]
                    with _lib.timed("granger"):
                        rc = lib.sc_granger_pairwise(_lib.ptr(csm[k0:k1]), _lib.ptr(power[k0:k1]), k1 - k0, n_freq, nfft,
                                                     1 if self._hermitian else 0, n_sig, _lib.ptr(pair_t), n_pairs,
                                                     float(tolerance), int(max_iterations),
                                                     1 if tail_extrapolation else 0, 1 if mixed_precision else 0,
                                                     _lib.ptr(tw128), _lib.ptr(tw64), _lib.ptr(dst),
                                                     _lib.ptr(it_c), _lib.ptr(fl_c), _lib.ptr(exec_cnt), _lib.ptr(gr_ws),
                                                     ws_bytes, st)
                        _lib.check(rc, "sc_granger_pairwise")
                    it_all[:, b0 + k0:b0 + k1] = it_c
                    fl_all[:, b0 + k0:b0 + k1] = fl_c
                    offload("pairwise_spectral_granger_prediction", b0 + k0, b0 + k1)
            del item, csm, power, plv, pli

        if want_granger:
            self.last_granger_iterations = it_all
            self.last_granger_flags = fl_all
            self.last_granger_executed = exec_cnt
            n_bad = int((fl_all & _lib.FLAG_NOT_CONVERGED).ne(0).sum())
            n_spd = int((fl_all & _lib.FLAG_NOT_SPD).ne(0).sum())
            if n_bad:  # minimum_phase_decomposition.py:318-322
                logger.warning(f"Maximum iterations reached. {fl_all.numel() - n_bad} of {fl_all.numel()} converged")
            if n_spd:  # minimum_phase_decomposition.py:78-82 (the reference falls back to a random start)
                logger.warning(f"Computing the initial conditions using the Cholesky failed for {n_spd} "
                               "(pair, window) problems; their Granger values are NaN.")

        if self._mt is not None:
            self._mt._check_finite_deferred()
        result = {}
        if copy_stream is not None:
            copy_stream.synchronize()
            for name, t in host.items():
                result[name] = t.reshape(kept + tuple(t.shape[1:])).numpy()
            return result
        for name, t in res.items():
            t = res_packed.get(name, t)
            if t.shape[1] != fnn:
                t = t[:, :fnn]
            tail = tuple(t.shape[1:])
            result[name] = t.reshape(kept + tail)
        return result

    def _one(self, name, **kw):
        return self.compute([name], **kw)[name]

    # ---- reference-named measures -------------------------------------------------------
    def _two_sided(self, name):
        """Expectation over all Nfft bins (the reference's private two-sided quantities)."""
        n_win, n_trials, n_tapers, nfft, n_sig = self._shape
        kept = self._kept_dims(n_local_windows=self._owned(nfft))
        n_batch = int(np.prod(kept)) if kept else 1
        shape = (n_batch, nfft, n_sig) if name == "power" else (n_batch, nfft, n_sig, n_sig)
        dtype = torch.float32 if name == "power" else torch.complex64
        out = torch.empty(shape, dtype=dtype, device=self._device)
        for item in self._reduced(nfft, [name]):
            out[item["o0"]:item["o1"]] = item[name]
        return self._finish(out.reshape(kept + tuple(out.shape[1:])))

    @property
    def _power(self):
        """Two-sided E[|X|^2] (connectivity.py:441-445)."""
        return self._two_sided("power")

    def _expectation_cross_spectral_matrix(self, fcn=None, dtype=None):
        """Two-sided expected cross-spectral matrix (connectivity.py:463-526).  Arbitrary Python
        ``fcn`` callbacks cannot run inside the fused kernels; the measures that need one
        (PLV, PLI family) have dedicated device modes."""
        if fcn is not None:
            raise NotImplementedError("per-observation callbacks are fused on the device; use the measure methods")
        return self._two_sided("csm")

    @property
    def _cross_spectral_matrix(self):
        """Un-averaged X_i conj(X_j), shape (W,T,K,Nfft,S,S) (connectivity.py:447-461).  Provided for
        API parity; the measures never materialise it."""
        lib = _lib.load()
        n_win, n_trials, n_tapers, nfft, n_sig = self._shape
        coef = self.fourier_coefficients
        nb = n_win * n_trials * n_tapers
        xp = torch.empty((nb, nfft, 2, 1, n_sig), dtype=torch.float32, device=self._device)
        mapping = (n_trials * n_tapers, n_tapers, 1, 0, 0, 0)
        _lib.check(lib.sc_repack_coefficients(_lib.ptr(coef), n_win, n_trials, n_tapers, nfft, n_sig, nfft,
                                              _lib.map6(mapping), 1, _lib.ptr(xp), _lib.stream_ptr()), "sc_repack")
        out = torch.empty((nb, nfft, n_sig, n_sig), dtype=torch.complex64, device=self._device)
        _lib.check(lib.sc_csm(_lib.ptr(xp), nb, nfft, 1, n_sig, 1.0, _lib.CSM_CROSS, _lib.ptr(out),
                              _lib.stream_ptr()), "sc_csm")
        return self._finish(out.reshape(n_win, n_trials, n_tapers, nfft, n_sig, n_sig))

    def power(self):
        """Power spectral density, (..., n_frequencies, n_signals) (connectivity.py:612-630)."""
        return self._one("power")

    def coherency(self):
        """Complex coherency, NaN diagonal (connectivity.py:632-657)."""
        return self._one("coherency")

    def coherence_phase(self):
        """connectivity.py:659-673."""
        return self._one("coherence_phase")

    def coherence_magnitude(self):
        """Squared coherence magnitude clipped to [0, 1] (connectivity.py:675-702)."""
        return self._one("coherence_magnitude")

    def imaginary_coherence(self):
        """connectivity.py:704-743."""
        return self._one("imaginary_coherence")

    def _phase_locking_value(self):
        """Complex E[x/|x|] (connectivity.py:897-903)."""
        return self._one("_phase_locking_value")

    def phase_locking_value(self):
        """connectivity.py:905-931."""
        return self._one("phase_locking_value")

    def phase_lag_index(self):
        """connectivity.py:933-982."""
        return self._one("phase_lag_index")

    def weighted_phase_lag_index(self):
        """connectivity.py:984-1028."""
        return self._one("weighted_phase_lag_index")

    def debiased_squared_phase_lag_index(self):
        """connectivity.py:1030-1058."""
        return self._one("debiased_squared_phase_lag_index")

    def debiased_squared_weighted_phase_lag_index(self):
        """connectivity.py:1060-1127."""
        return self._one("debiased_squared_weighted_phase_lag_index")

    def pairwise_phase_consistency(self):
        """connectivity.py:1129-1159."""
        return self._one("pairwise_phase_consistency")

    def pairwise_spectral_granger_prediction(self, tolerance=1e-8, max_iterations=60, tail_extrapolation=True,
                                             mixed_precision=None):
        """Spectral Granger prediction for every signal pair; [..., i, j] is the influence
        j -> i (connectivity.py:1161-1191)."""
        return self._one("pairwise_spectral_granger_prediction", tolerance=tolerance,
                         max_iterations=max_iterations, tail_extrapolation=tail_extrapolation,
                         mixed_precision=mixed_precision)

    def subset_pairwise_spectral_granger_prediction(self, pairs, tolerance=1e-8, max_iterations=60,
                                                    tail_extrapolation=True, mixed_precision=None):
        """connectivity.py:1193-1213."""
        return self._one("pairwise_spectral_granger_prediction", pairs=pairs, tolerance=tolerance,
                         max_iterations=max_iterations, tail_extrapolation=tail_extrapolation,
                         mixed_precision=mixed_precision)

    def conditional_spectral_granger_prediction(self):
        raise NotImplementedError  # connectivity.py:1215-1219

    def blockwise_spectral_granger_prediction(self):
        raise NotImplementedError  # connectivity.py:1221-1224

    # ---- MVAR family from the full-matrix Wilson factor ----------
    _GRANGER_SUB_WINDOWS = 3   # host output: windows per Wilson sub-launch of a chunk (see compute)
    _MVAR_MAX_SIGNALS = 1024
    _MVAR_CACHE_BYTES = 48 << 30  # above this the factor is not cached: measures stream over window chunks

    def _mvar_check(self):
        if self.expectation_type not in _GRANGER_OK:
            raise NotImplementedError(
                f"MVAR measures with expectation_type='{self.expectation_type}' couple the Wilson convergence test "
                "across a kept axis; only " + ", ".join(_GRANGER_OK) + " are supported.")
        n_sig = self._shape[4]
        if n_sig > self._MVAR_MAX_SIGNALS:
            raise NotImplementedError(
                f"full-matrix Wilson factorisation on the device handles up to {self._MVAR_MAX_SIGNALS} signals "
                f"(got {n_sig}); use pairwise_spectral_granger_prediction or a signal subset")

    def _mvar_bytes(self):
        """Device bytes of the cached path: CSM, G, H, A and the Wilson workspace for every kept index."""
        n_win, n_trials, n_tapers, nfft, n_sig = self._shape
        n_freq = nfft // 2 + 1 if self._hermitian else nfft
        kept = self._kept_dims(n_local_windows=self._owned(n_freq))
        n_batch = int(np.prod(kept)) if kept else 1
        return 8 * n_batch * n_freq * n_sig * n_sig * 16

    def _mvar_from_csm(self, csm, tolerance, max_iterations):
        """Expected CSM c64 [b][n_freq][S][S] -> dict(g, h, sigma, a, iters, flags) (connectivity.py:567-588,
        1679-1748).  The Tikhonov terms use the mean over the batch passed in."""
        lib = _lib.load()
        nfft, n_sig = self._shape[3], self._shape[4]
        fnn = nfft // 2 + 1
        herm = 1 if self._hermitian else 0
        n_batch, n_freq = csm.shape[0], csm.shape[1]
        dev = self._device
        st = _lib.stream_ptr()
        csm = csm.to(torch.complex128)
        g = torch.empty_like(csm)
        iters = torch.zeros(n_batch, dtype=torch.int32, device=dev)
        flags = torch.zeros(n_batch, dtype=torch.int32, device=dev)
        tw = twiddles(nfft, torch.complex128, dev)
        ws_bytes = lib.sc_wilson_general_workspace_bytes(n_batch, n_freq, n_sig)
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
        _lib.check(lib.sc_wilson(_lib.ptr(csm), n_batch, n_freq, nfft, herm, n_sig, float(tolerance), int(max_iterations),
                                 _lib.ptr(tw), _lib.ptr(g), _lib.ptr(iters), _lib.ptr(flags), _lib.ptr(ws), ws_bytes, st),
                   "sc_wilson")
        del csm
        h0 = torch.empty((n_batch, n_sig, n_sig), dtype=torch.float64, device=dev)
        _lib.check(lib.sc_mvar_lag0(_lib.ptr(g), n_batch, n_freq, nfft, herm, n_sig, _lib.ptr(h0), st), "sc_mvar_lag0")
        # Tikhonov terms (mean over all windows of the batch, connectivity.py:1742-1746, :585) stay on the device: a
        # scalar tensor passed by pointer, no host synchronisation between Wilson and the measures
        lam = TIKHONOV_REGULARIZATION_FACTOR * (h0 * h0).mean()
        h = torch.empty((n_batch, fnn, n_sig, n_sig), dtype=torch.complex128, device=dev)
        sigma = torch.empty((n_batch, n_sig, n_sig), dtype=torch.float64, device=dev)
        need = max(lib.sc_mvar_workspace_bytes(n_batch, 2, n_sig), lib.sc_mvar_workspace_bytes(n_batch * fnn, 1, n_sig))
        if need > ws_bytes:
            del ws
            ws = torch.empty(need, dtype=torch.uint8, device=dev)
            ws_bytes = need
        _lib.check(lib.sc_mvar_transfer(_lib.ptr(g), _lib.ptr(h0), 0.0, _lib.ptr(lam), n_batch, n_freq, fnn, n_sig,
                                        _lib.ptr(h), _lib.ptr(sigma), _lib.ptr(ws), ws_bytes, st), "sc_mvar_transfer")
        hr = torch.view_as_real(h)
        lam_a = TIKHONOV_REGULARIZATION_FACTOR * (hr * hr).sum() / h.numel()
        del hr
        a = torch.empty_like(h)
        _lib.check(lib.sc_mvar_inverse(_lib.ptr(h), 0.0, _lib.ptr(lam_a), n_batch * fnn, n_sig, _lib.ptr(a), _lib.ptr(ws),
                                       ws_bytes, st), "sc_mvar_inverse")
        return dict(g=g, h=h, sigma=sigma, a=a, iters=iters, flags=flags, n_batch=n_batch, fnn=fnn, n_sig=n_sig)

    def _mvar_warn(self, flags):
        n_batch = flags.numel()
        n_bad = int((flags & _lib.FLAG_NOT_CONVERGED).ne(0).sum())
        if n_bad:
            logger.warning(f"Maximum iterations reached. {n_batch - n_bad} of {n_batch} converged")
        if int((flags & _lib.FLAG_NOT_SPD).ne(0).sum()):
            logger.warning("Computing the initial conditions using the Cholesky failed; those factors are NaN.")

    def _mvar(self, tolerance=1e-8, max_iterations=60):
        """Factor the expected CSM once and cache H (non-negative bins), the noise covariance and the
        MVAR Fourier coefficients (the reference re-runs Wilson on every property access,
        connectivity.py:567-588)."""
        if getattr(self, "_mvar_cache", None) is not None:
            return self._mvar_cache
        self._mvar_check()
        lib = _lib.load()
        n_win, n_trials, n_tapers, nfft, n_sig = self._shape
        n_freq = nfft // 2 + 1 if self._hermitian else nfft
        kept = self._kept_dims(n_local_windows=self._owned(n_freq))
        n_batch = int(np.prod(kept)) if kept else 1
        csm = self._local_csm(n_freq)
        m = self._mvar_from_csm(csm, tolerance, max_iterations)
        self._mvar_warn(m["flags"])
        self.last_wilson_iterations, self.last_wilson_flags = m["iters"], m["flags"]
        m["kept"] = kept
        self._mvar_cache = m
        return m

    def _mvar_measure_of(self, m, code):
        lib = _lib.load()
        nb, fnn, n_sig = m["n_batch"], m["fnn"], m["n_sig"]
        out = torch.empty((nb, fnn, n_sig, n_sig), dtype=torch.float32, device=self._device)
        n_scratch = nb * n_sig + (2 * nb * fnn * n_sig if n_sig > 32 else 0)
        scratch = torch.empty(n_scratch, dtype=torch.float64, device=self._device)
        _lib.check(lib.sc_mvar_measure(code, _lib.ptr(m["h"]), _lib.ptr(m["a"]), _lib.ptr(m["sigma"]), nb, fnn, n_sig,
                                       _lib.ptr(scratch), _lib.ptr(out), _lib.stream_ptr()), "sc_mvar_measure")
        return out

    def _mvar_measure(self, code, tolerance=1e-8, max_iterations=60):
        if getattr(self, "_mvar_cache", None) is not None or self._mvar_bytes() <= self._MVAR_CACHE_BYTES:
            m = self._mvar(tolerance, max_iterations)
            out = self._mvar_measure_of(m, code)
            return self._finish(out.reshape(m["kept"] + tuple(out.shape[1:])))
        # Too large to keep the factor of every window (BASELINE config 5: 256 GB): stream window chunks,
        # factor + measure per chunk, keep only the result.  The 1e-12-relative Tikhonov terms then use the
        # chunk mean instead of the mean over all windows (a 1e-12-relative change).
        self._mvar_check()
        n_win, n_trials, n_tapers, nfft, n_sig = self._shape
        n_freq = nfft // 2 + 1 if self._hermitian else nfft
        fnn = nfft // 2 + 1
        kept = self._kept_dims(n_local_windows=self._owned(n_freq))
        n_batch = int(np.prod(kept)) if kept else 1
        to_host = self._output != "torch"
        result = (torch.empty((n_batch, fnn, n_sig, n_sig), dtype=torch.float32, pin_memory=True) if to_host else
                  torch.empty((n_batch, fnn, n_sig, n_sig), dtype=torch.float32, device=self._device))
        all_iters = torch.zeros(n_batch, dtype=torch.int32, device=self._device)
        all_flags = torch.zeros(n_batch, dtype=torch.int32, device=self._device)
        per_item = 8 * n_freq * n_sig * n_sig * 16
        sub = max(1, self._MVAR_CACHE_BYTES // per_item)
        copy_stream = _lib.side_stream(self._device, "d2h") if to_host else None
        for item in self._reduced(n_freq, ["csm"]):
            b0, b1, csm = item["o0"], item["o1"], item["csm"]
            for c0 in range(0, b1 - b0, sub):
                c1 = min(b1 - b0, c0 + sub)
                m = self._mvar_from_csm(csm[c0:c1], tolerance, max_iterations)
                out = self._mvar_measure_of(m, code)
                all_iters[b0 + c0:b0 + c1] = m["iters"]
                all_flags[b0 + c0:b0 + c1] = m["flags"]
                if to_host:  # device -> host on the copy stream, under the next sub-chunk's factorisation
                    ev = torch.cuda.Event()
                    ev.record()
                    copy_stream.wait_event(ev)
                    out.record_stream(copy_stream)
                    with torch.cuda.stream(copy_stream):
                        result[b0 + c0:b0 + c1].copy_(out, non_blocking=True)
                else:
                    result[b0 + c0:b0 + c1].copy_(out, non_blocking=True)
                del m, out
            del item, csm
        if copy_stream is not None:
            copy_stream.synchronize()
        self._mvar_warn(all_flags)
        self.last_wilson_iterations, self.last_wilson_flags = all_iters, all_flags
        result = result.reshape(kept + (fnn, n_sig, n_sig))
        return result.numpy() if to_host else result

    @property
    def _minimum_phase_factor(self):
        """connectivity.py:567-569 (half spectrum on the real-series path)."""
        m = self._mvar()
        return self._finish(m["g"].reshape(m["kept"] + tuple(m["g"].shape[1:])))

    @property
    def _transfer_function(self):
        """connectivity.py:571-574, non-negative frequencies."""
        m = self._mvar()
        return self._finish(m["h"].reshape(m["kept"] + tuple(m["h"].shape[1:])))

    @property
    def _noise_covariance(self):
        """connectivity.py:576-578."""
        m = self._mvar()
        return self._finish(m["sigma"].reshape(m["kept"] + tuple(m["sigma"].shape[1:])))

    @property
    def _MVAR_Fourier_coefficients(self):
        """connectivity.py:580-588."""
        m = self._mvar()
        return self._finish(m["a"].reshape(m["kept"] + tuple(m["a"].shape[1:])))

    def directed_transfer_function(self):
        """connectivity.py:1237-1266."""
        return self._mvar_measure(0)

    def directed_coherence(self):
        """connectivity.py:1268-1296."""
        return self._mvar_measure(1)

    def partial_directed_coherence(self, keep_cupy=False):
        """connectivity.py:1298-1343."""
        return self._mvar_measure(2)

    def generalized_partial_directed_coherence(self):
        """connectivity.py:1345-1380."""
        return self._mvar_measure(3)

    def direct_directed_transfer_function(self):
        """connectivity.py:1382-1426."""
        return self._mvar_measure(4)

    def _trials_tapers_csm(self, n_freq, private=False):
        """Expected CSM over trials x tapers per window (what the SVD-based measures are built on; they
        merge trials and tapers whatever ``expectation_type`` says, connectivity.py:1953-1976).  Rows = this rank's
        windows (all of them unless reduce_scatter)."""
        n_win, n_trials, n_tapers, nfft, n_sig = self._shape
        gs = self._group_state
        n_obs = gs.n_trials_tapers if gs is not None else n_trials * n_tapers   # global count under a reduce group
        return self._local_csm(n_freq, "trials_tapers", scale=1.0 / n_obs, private=private)

    def canonical_coherence(self, group_labels):
        """Squared canonical coherence between signal groups, shape (n_windows, n_frequencies, n_groups,
        n_groups) with NaN diagonal, and the sorted group labels (connectivity.py:745-820).  Computed from
        the expected CSM (block whitening C_aa^-1/2 C_ab C_bb^-1/2) instead of per-group SVDs of the coefficients.

        * groups of up to 64 signals: one fused kernel (two Cholesky factorisations, two triangular solves and the
          top eigenvalue per (window, frequency, group pair) in shared memory) -- BASELINE config 5's grouping;
        * a group with at least as many signals as observations (trials x tapers) spans the whole observation space:
          the reference's SVD whitening U V^H then has V unitary, every singular value of the whitened cross product
          is exactly 1, and so is the result for every pair that involves such a group (connectivity.py:1997-2000,
          2027-2031) -- returned directly;
        * larger full-rank groups: same block whitening through batched cuSOLVER calls (torch.linalg, complex128),
          a library path kept for completeness."""
        lib = _lib.load()
        group_labels = np.asarray(group_labels)
        n_win, n_trials, n_tapers, nfft, n_sig = self._shape
        if group_labels.shape != (n_sig,):
            raise ValueError(f"group_labels must have one entry per signal ({n_sig}), got shape {group_labels.shape}")
        labels = np.unique(group_labels)
        n_groups = len(labels)
        fnn = nfft // 2 + 1
        n_loc = self._owned(fnn, "trials_tapers")
        n_win = n_win if n_loc is None else n_loc       # reduce_scatter: this rank's windows only
        out = torch.full((n_win, fnn, n_groups, n_groups), float("nan"), dtype=torch.float32, device=self._device)
        if n_groups >= 2 and n_win > 0:
            members = [np.flatnonzero(group_labels == lab) for lab in labels]
            sizes = np.array([len(m) for m in members])
            gs = self._group_state
            n_obs = gs.n_trials_tapers if gs is not None else n_trials * n_tapers
            saturated = sizes >= n_obs                  # rank-deficient groups: canonical coherence is exactly 1
            regular = np.flatnonzero(~saturated)
            if len(regular) >= 2:
                csm = self._trials_tapers_csm(fnn)
                sub = torch.full((n_win, fnn, len(regular), len(regular)), float("nan"), dtype=torch.float32,
                                 device=self._device)
                if sizes[regular].max() <= 64:
                    order = np.concatenate([members[g] for g in regular]).astype(np.int32)
                    offsets = np.concatenate([[0], np.cumsum(sizes[regular])]).astype(np.int32)
                    gidx = torch.from_numpy(order).to(self._device)
                    goff = torch.from_numpy(offsets).to(self._device)
                    flags = torch.zeros(n_win * fnn, dtype=torch.int32, device=self._device)
                    _lib.check(lib.sc_canonical_coherence(_lib.ptr(csm), n_win, fnn, n_sig, _lib.ptr(gidx), _lib.ptr(goff),
                                                          len(regular), int(sizes[regular].max()), _lib.ptr(sub),
                                                          _lib.ptr(flags), _lib.stream_ptr()), "sc_canonical_coherence")
                    bad = int(flags.ne(0).sum())
                else:
                    bad = self._canonical_large_groups(csm, [members[g] for g in regular], sub)
                if bad:
                    logger.warning("canonical_coherence: some group blocks of the cross-spectral matrix are not positive "
                                   "definite (linearly dependent signals in a group); those entries are NaN.")
                ridx = torch.from_numpy(regular).to(self._device)
                out[:, :, ridx[:, None], ridx[None, :]] = sub
            for g in np.flatnonzero(saturated):
                out[:, :, g, :] = 1.0
                out[:, :, :, g] = 1.0
            diag = torch.arange(n_groups, device=self._device)
            out[:, :, diag, diag] = float("nan")
        return self._finish(out), labels

    def _canonical_large_groups(self, csm, members, out):
        """Block whitening for groups of more than 64 signals (batched cuSOLVER through torch.linalg, complex128):
        sigma_max^2(L_a^-1 C_ab L_b^-H) with C_gg = L_g L_g^H.  Returns the number of non-positive-definite blocks."""
        dev = self._device
        idx = [torch.from_numpy(np.asarray(m)).to(dev) for m in members]
        n_bad = 0
        per_window = csm.shape[1] * max(len(m) for m in members) ** 2 * 16 * 6
        step = max(1, (2 << 30) // max(per_window, 1))
        for w0 in range(0, csm.shape[0], step):
            c = csm[w0:w0 + step].to(torch.complex128)
            chol = []
            for ia in idx:
                fac, info = torch.linalg.cholesky_ex(c[:, :, ia[:, None], ia[None, :]])
                n_bad += int(info.ne(0).sum())
                chol.append((fac, info.ne(0)))
            for a in range(len(idx)):
                for b in range(a + 1, len(idx)):
                    cab = c[:, :, idx[a][:, None], idx[b][None, :]]
                    m = torch.linalg.solve_triangular(chol[a][0], cab, upper=False)
                    m = torch.linalg.solve_triangular(chol[b][0], m.mH, upper=False)
                    val = torch.linalg.matrix_norm(m, ord=2) ** 2
                    val = torch.where(chol[a][1] | chol[b][1], torch.full_like(val, float("nan")), val)
                    out[w0:w0 + step, :, a, b] = out[w0:w0 + step, :, b, a] = val.to(torch.float32)
        return n_bad

    def global_coherence(self, max_rank=1):
        """The ``max_rank`` largest eigenvalues of the cross-spectral matrix per (window, frequency) over ALL n_fft
        bins and their eigenvectors: shapes (n_windows, n_fft, max_rank) and (n_windows, n_fft, n_signals, max_rank)
        (connectivity.py:822-895, 2245-2279).  Eigenvectors are defined up to a phase, as in the reference's SVD.
        Eigenpairs are found one at a time (repeated squaring, then deflation C <- C - lambda v v^H); the ORDER
        along the last axis is the reference's: descending from its dense SVD when max_rank >= n_signals - 1,
        ascending from scipy's ``svds`` otherwise (:2258-2276)."""
        lib = _lib.load()
        n_win, n_trials, n_tapers, nfft, n_sig = self._shape
        max_rank = int(max_rank)
        if max_rank < 1:
            raise ValueError("max_rank must be at least 1")
        if n_sig > 1024:
            raise NotImplementedError("global_coherence on the device handles up to 1024 signals")
        gs = self._group_state
        n_obs = gs.n_trials_tapers if gs is not None else n_trials * n_tapers
        n_keep = min(max_rank, n_sig, n_obs)            # the SVD of an (S x TK) matrix has min(S, TK) singular values
        csm = self._trials_tapers_csm(nfft, private=max_rank > 1)   # deflation overwrites its buffer
        n_win = csm.shape[0]                            # reduce_scatter: this rank's windows only
        val = torch.empty((n_keep, n_win, nfft), dtype=torch.float32, device=self._device)
        vec = torch.empty((n_keep, n_win, nfft, n_sig), dtype=torch.complex64, device=self._device)
        n_mat = n_win * nfft
        per = max(1, lib.sc_global_coherence_workspace_bytes(1, n_sig))
        chunk = n_mat if n_sig <= 64 else max(1, min(n_mat, (8 << 30) // per))  # <= 8 GiB of squaring workspace
        ws_bytes = lib.sc_global_coherence_workspace_bytes(chunk, n_sig)
        ws = torch.empty(max(ws_bytes, 1), dtype=torch.uint8, device=self._device)
        csm_f = csm.reshape(n_mat, n_sig, n_sig)
        st = _lib.stream_ptr()
        for r in range(n_keep):
            val_f, vec_f = val[r].reshape(n_mat), vec[r].reshape(n_mat, n_sig)
            for m0 in range(0, n_mat, max(chunk, 1)):
                m1 = min(n_mat, m0 + chunk)
                _lib.check(lib.sc_global_coherence(_lib.ptr(csm_f[m0:m1]), m1 - m0, n_sig, _lib.ptr(val_f[m0:m1]),
                                                   _lib.ptr(vec_f[m0:m1]), _lib.ptr(ws) if ws_bytes else None, ws_bytes,
                                                   st), "sc_global_coherence")
            if r + 1 < n_keep and n_mat:
                _lib.check(lib.sc_hermitian_deflate(_lib.ptr(csm_f), n_mat, n_sig, _lib.ptr(val_f), _lib.ptr(vec_f), st),
                           "sc_hermitian_deflate")
        val = val.permute(1, 2, 0)                      # (W, nfft, rank), descending
        vec = vec.permute(1, 2, 3, 0)                   # (W, nfft, S, rank)
        if max_rank < n_sig - 1:                        # the reference's svds branch returns ascending order
            val, vec = val.flip(-1), vec.flip(-1)
        return self._finish(val.contiguous()), self._finish(vec.contiguous())

    def _significant_band_phase(self, frequencies_of_interest, frequency_resolution, significance_threshold):
        """Shared front half of ``delay`` / ``group_delay`` (connectivity.py:1466-1495, 1556-1578): band-passed
        coherency of every signal pair (i < j), its unwrapped phase masked where the coherence is not significant.
        The coherency comes from the device; the significance bookkeeping is host index work (_statistics.py)."""
        from . import _statistics as stats
        freqs = np.asarray(self.frequencies)
        step = 1 if frequency_resolution is None else int(np.ceil(frequency_resolution / (freqs[1] - freqs[0])))
        coh = self.coherency()
        coh = coh.cpu().numpy() if isinstance(coh, torch.Tensor) else np.asarray(coh)
        if frequencies_of_interest is not None:
            band = (frequencies_of_interest[0] < freqs) & (freqs < frequencies_of_interest[1])
            coh, freqs = np.take(coh, band.nonzero()[0], axis=-3), freqs[band]
        n_sig = coh.shape[-1]
        pairs = np.asarray(list(combinations(np.arange(n_sig), 2)))
        coh = coh[..., pairs[:, 0], pairs[:, 1]].astype(np.complex128)
        keep = stats.significant_frequencies(coh, self.n_observations, step,
                                             significance_threshold=significance_threshold)
        phase = np.ma.masked_array(np.unwrap(np.angle(coh), axis=-2), mask=~keep)
        return phase, freqs, pairs, n_sig

    def group_delay(self, frequencies_of_interest=None, frequency_resolution=None, significance_threshold=0.05):
        """Average time delay of a broadband signal: slope of the significant coherence phase against frequency
        (connectivity.py:1428-1528).  Returns (delay, slope, r_value), each (..., n_signals, n_signals)."""
        from scipy.stats.mstats import linregress
        phase, freqs, pairs, n_sig = self._significant_band_phase(frequencies_of_interest, frequency_resolution,
                                                                  significance_threshold)
        fit = np.ma.apply_along_axis(lambda y: linregress(freqs, y=y), -2, phase)
        shape = (*phase.shape[:-2], n_sig, n_sig)
        i, j = pairs[:, 0], pairs[:, 1]
        slope = np.full(shape, np.nan)
        slope[..., i, j] = np.asarray(fit[..., 0, :], dtype=float)
        slope[..., j, i] = -1 * np.asarray(fit[..., 0, :], dtype=float)
        r_value = np.ones(shape)
        r_value[..., i, j] = np.asarray(fit[..., 2, :], dtype=float)
        r_value[..., j, i] = np.asarray(fit[..., 2, :], dtype=float)
        return slope / (2 * np.pi), slope, r_value

    def delay(self, frequencies_of_interest=None, frequency_resolution=None, significance_threshold=0.05, n_range=3):
        """Candidate delays (coherence phase + 2 pi k) / (2 pi), k = -n_range..n_range, of the significant
        frequencies (connectivity.py:1530-1585).  Shape (..., n_frequencies, 2 n_range + 1, n_signals, n_signals)."""
        phase, _, pairs, n_sig = self._significant_band_phase(frequencies_of_interest, frequency_resolution,
                                                              significance_threshold)
        turns = 2 * np.pi * np.arange(-n_range, n_range + 1)
        cand = np.rollaxis((turns + phase[..., np.newaxis]) / (2 * np.pi), -1, -2)
        out = np.full((*phase.shape[:-1], len(turns), n_sig, n_sig), np.nan)
        out[..., pairs[:, 0], pairs[:, 1]] = cand
        out[..., pairs[:, 1], pairs[:, 0]] = -cand
        return out

    def phase_slope_index(self, frequencies_of_interest=None, frequency_resolution=None):
        """Weighted average of the coherency phase slope projected on the imaginary axis, shape
        (..., n_signals, n_signals) (connectivity.py:1587-1650).  The coherency of each window chunk stays on
        the device; the band-pass (lo < f < hi) and the independent-frequency subsampling are index work done
        on the host with the reference's formulas."""
        lib = _lib.load()
        n_win, n_trials, n_tapers, nfft, n_sig = self._shape
        fnn = nfft // 2 + 1
        freqs = np.asarray(self.frequencies)
        keep = np.arange(fnn)
        if frequencies_of_interest is not None:
            keep = np.flatnonzero((frequencies_of_interest[0] < freqs) & (freqs < frequencies_of_interest[1]))
        step = 1
        if frequency_resolution is not None:  # connectivity.py:2076-2100
            step = int(np.ceil(frequency_resolution / (freqs[1] - freqs[0])))
        keep = keep[np.arange(0, keep.size, step)]
        if keep.size < 2:
            raise IndexError("phase_slope_index needs at least two frequencies in the band of interest")
        fidx = torch.from_numpy(keep.astype(np.int32)).to(self._device)
        kept = self._kept_dims(n_local_windows=self._owned(fnn))
        n_batch = int(np.prod(kept)) if kept else 1
        out = torch.empty((n_batch, n_sig, n_sig), dtype=torch.float32, device=self._device)
        st = _lib.stream_ptr()
        for item in self._reduced(fnn, ["csm", "power"]):
            b0, b1, csm, power = item["o0"], item["o1"], item["csm"], item["power"]
            nb = b1 - b0
            if nb == 0:
                continue
            coh = torch.empty_like(csm)      # not in place: the CSM may be the shared cache
            _lib.check(lib.sc_pairwise_epilogue(_lib.M_COHERENCY, _lib.ptr(csm), _lib.ptr(power), nb, fnn, n_sig,
                                                float(self.n_observations), _lib.ptr(coh), st), "sc_pairwise_epilogue")
            _lib.check(lib.sc_phase_slope_index(_lib.ptr(coh), nb, fnn, n_sig, _lib.ptr(fidx), int(keep.size),
                                                _lib.ptr(out[b0:b1]), st), "sc_phase_slope_index")
        return self._finish(out.reshape(kept + (n_sig, n_sig)))


def pinned_empty(shape, dtype=np.float32):
    """Page-locked host array (NumPy view of a pinned torch tensor) for ``Connectivity.compute(out=...)``: reuse it
    across calls so that a pipeline does not allocate and page-lock its result buffers per call."""
    t = torch.empty(tuple(shape), dtype=getattr(torch, np.dtype(dtype).name), pin_memory=True)
    return t.numpy()


def unpack_upper(packed, n_signals):
    """Full symmetric array (..., S, S) from the packed upper triangles (..., S (S + 1) / 2) that
    ``Connectivity.compute(packed=...)`` returns (NumPy array or torch tensor)."""
    i, j = np.triu_indices(n_signals)
    if isinstance(packed, torch.Tensor):
        full = torch.empty(tuple(packed.shape[:-1]) + (n_signals, n_signals), dtype=packed.dtype, device=packed.device)
        ti, tj = torch.from_numpy(i).to(packed.device), torch.from_numpy(j).to(packed.device)
        full[..., ti, tj] = packed
        full[..., tj, ti] = packed
        return full
    full = np.empty(tuple(packed.shape[:-1]) + (n_signals, n_signals), dtype=packed.dtype)
    full[..., i, j] = packed
    full[..., j, i] = packed
    return full


def all_pairs(n_signals):
    return np.array(list(combinations(range(n_signals), 2)), dtype=np.int32)
