"""ctypes binding of libsc_b200.so (the C ABI declared in include/sc_b200.h).

There is no CPU fallback: if the CUDA library is missing the import of any compute
entry point fails loudly with build instructions.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import c_char_p, c_double, c_float, c_int, c_int64, c_void_p

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("SC_B200_LIB") or os.path.join(_HERE, "libsc_b200.so")  # override: kernel experiments

# symbol -> (restype, argtypes); must list every function of include/sc_b200.h
_P = c_void_p
SIGNATURES = {
    "sc_version": (c_int, []),
    "sc_last_error": (c_char_p, []),
    "sc_mt_fft": (c_int, [_P, c_int64, c_int64, c_int64, _P, c_int, c_int, c_int, c_int64, c_int64, c_int64,
                          c_int, c_int, c_float, _P, c_int, c_int, ctypes.POINTER(c_int64), c_int64, _P, _P,
                          c_int64, _P]),
    "sc_mt_fft_workspace_bytes": (c_int64, [c_int, c_int]),
    "sc_repack_coefficients": (c_int, [_P, c_int64, c_int64, c_int64, c_int64, c_int64, c_int,
                                       ctypes.POINTER(c_int64), c_int64, _P, _P]),
    "sc_power": (c_int, [_P, c_int64, c_int64, c_int64, c_int64, c_float, _P, _P]),
    "sc_nonfinite_flag": (c_int, [_P, c_int64, _P, _P]),
    "sc_power_from_csm": (c_int, [_P, c_int64, c_int64, _P, _P]),
    "sc_pack_upper": (c_int, [_P, c_int64, c_int64, _P, _P]),
    "sc_csm": (c_int, [_P, c_int64, c_int64, c_int64, c_int64, c_float, c_int, _P, _P]),
    "sc_csm_simt": (c_int, [_P, c_int64, c_int64, c_int64, c_int64, c_float, c_int, _P, _P]),
    "sc_pairwise_epilogue": (c_int, [c_int, _P, _P, c_int64, c_int64, c_int64, c_double, _P, _P]),
    "sc_phase_slope_index": (c_int, [_P, c_int64, c_int64, c_int64, _P, c_int, _P, _P]),
    "sc_wilson2": (c_int, [_P, c_int64, c_int, c_double, c_int, _P, _P, _P, _P, _P, c_int64, _P]),
    "sc_granger_pairwise": (c_int, [_P, _P, c_int64, c_int, c_int, c_int, c_int64, _P, c_int64, c_double, c_int,
                                    c_int, c_int, _P, _P, _P, _P, _P, _P, _P, c_int64, _P]),
    "sc_simt_peak": (c_int, [c_int, c_int, _P, ctypes.POINTER(c_double), _P]),
    "sc_simt_peak_scratch_bytes": (c_int64, []),
    "sc_wilson_workspace_bytes": (c_int64, [c_int]),
    "sc_wilson": (c_int, [_P, c_int64, c_int, c_int, c_int, c_int, c_double, c_int, _P, _P, _P, _P, _P, c_int64, _P]),
    "sc_wilson_general_workspace_bytes": (c_int64, [c_int64, c_int, c_int]),
    "sc_mvar_lag0": (c_int, [_P, c_int64, c_int, c_int, c_int, c_int, _P, _P]),
    "sc_mvar_transfer": (c_int, [_P, _P, c_double, _P, c_int64, c_int, c_int, c_int, _P, _P, _P, c_int64, _P]),
    "sc_mvar_inverse": (c_int, [_P, c_double, _P, c_int64, c_int, _P, _P, c_int64, _P]),
    "sc_mvar_workspace_bytes": (c_int64, [c_int64, c_int, c_int]),
    "sc_canonical_coherence": (c_int, [_P, c_int64, c_int, c_int, _P, _P, c_int, c_int, _P, _P, _P]),
    "sc_global_coherence": (c_int, [_P, c_int64, c_int, _P, _P, _P, c_int64, _P]),
    "sc_global_coherence_workspace_bytes": (c_int64, [c_int64, c_int]),
    "sc_hermitian_deflate": (c_int, [_P, c_int64, c_int, _P, _P, _P]),
    "sc_peer_reduce_csm": (c_int, [_P, c_int, c_int64, c_int64, c_int, _P, _P, c_int, _P, _P]),
    "sc_peer_reduce": (c_int, [_P, c_int, c_int64, c_int64, _P, _P]),
    "sc_mvar_measure": (c_int, [c_int, _P, _P, _P, c_int64, c_int, c_int, _P, _P, _P]),
}

# constants of include/sc_b200.h
DETREND = {None: 0, "constant": 1, "c": 1, "linear": 2, "l": 2}
LAYOUT_PLANAR, LAYOUT_REFERENCE = 0, 1
CSM_CROSS, CSM_PLV, CSM_PLI = 0, 1, 2
M_COHERENCY, M_COHERENCE_MAG, M_COHERENCE_PHASE, M_IMAG_COHERENCE = 0, 1, 2, 3
M_PLV, M_PPC, M_PLI, M_WPLI, M_DPLI2, M_DWPLI2 = 4, 5, 6, 7, 8, 9
FLAG_NOT_CONVERGED, FLAG_NOT_SPD = 1, 2

_lib = None
LAUNCHES = 0  # kernels launched through the ABI by this process (every checked call launches one)


class StageTimer:
    """Optional per-stage CUDA-event timing (bench.py): events are recorded on the current
    stream around each ABI call; ``totals()`` synchronises and sums milliseconds per stage."""

    def __init__(self):
        self.events = []

    def span(self, name):
        import torch
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        self.events.append((name, a, b))
        return a, b

    def totals(self):
        import torch
        torch.cuda.synchronize()
        out = {}
        for name, a, b in self.events:
            ms, n = out.get(name, (0.0, 0))
            out[name] = (ms + a.elapsed_time(b), n + 1)
        return out


TIMER = None  # set to a StageTimer to collect per-kernel times


class timed:
    """with timed("stage"): <one ABI call>"""

    def __init__(self, name):
        self.name = name
        self.ev = None

    def __enter__(self):
        if TIMER is not None:
            self.ev = TIMER.span(self.name)
            self.ev[0].record()
        return self

    def __exit__(self, *exc):
        if self.ev is not None:
            self.ev[1].record()
        return False


class NativeLibraryError(ImportError):
    pass


def load():
    """Load libsc_b200.so once; raise NativeLibraryError if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise NativeLibraryError(
            f"{LIB_PATH} not found. The B200 CUDA library is required (no CPU fallback). "
            "Build it with `python -c 'import __graft_entry__ as g; g.build()'` or "
            "`make -C spectral_connectivity_b200/csrc`.")
    import torch  # noqa: F401  (loads libcudart.so.12 that the library links against)
    lib = ctypes.CDLL(LIB_PATH, mode=ctypes.RTLD_GLOBAL)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


_SIDE_STREAMS = {}


def side_stream(device, kind):
    """Persistent per-device copy stream ("h2d" / "d2h").  A fresh torch.cuda.Stream() per call would walk
    through torch's stream pool, and the caching allocator keeps free blocks per stream: temporaries of one
    call could not be reused by the next, which then pays cudaMalloc (sporadic 20-300 ms) instead."""
    import torch
    key = (torch.device(device).index, kind)
    st = _SIDE_STREAMS.get(key)
    if st is None:
        # the collective stream gets the higher priority: its few CTAs must be placed before the persistent
        # Wilson / Granger grid of the previous chunk fills every SM
        st = _SIDE_STREAMS[key] = torch.cuda.Stream(device=device, priority=-1 if kind == "comm" else 0)
    return st


def check(rc, what=""):
    global LAUNCHES
    LAUNCHES += 1
    if rc != 0:
        msg = load().sc_last_error()
        raise RuntimeError(f"{what} failed (code {rc}): {msg.decode() if msg else ''}")


def ptr(t):
    """Device pointer of a torch tensor (or None)."""
    return None if t is None else c_void_p(t.data_ptr())


def stream_ptr():
    import torch
    return c_void_p(torch.cuda.current_stream().cuda_stream)


def map6(values):
    return (c_int64 * 6)(*[int(v) for v in values])
