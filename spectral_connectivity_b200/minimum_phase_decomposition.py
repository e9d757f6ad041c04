"""Wilson spectral factorisation, drop-in for
``spectral_connectivity.minimum_phase_decomposition.minimum_phase_decomposition``
(minimum_phase_decomposition.py:227-322), backed by ``sc_wilson2`` (csrc/wilson.cu)."""
from __future__ import annotations

from logging import getLogger

import numpy as np
import torch

from . import _lib
from .transforms import twiddles

logger = getLogger(__name__)


def minimum_phase_decomposition(cross_spectral_matrix, tolerance=1e-8, max_iterations=60,
                                return_info=False):
    """Minimum-phase factor G with S = G G^H.

    ``cross_spectral_matrix``: (n_time_points, n_fft_samples, 2, 2) or (n_fft_samples, 2, 2),
    two-sided in frequency, NumPy or torch.  Each leading index converges independently
    (minimum_phase_decomposition.py:310-315).  Returns complex128 (NumPy for NumPy input, CUDA
    tensor for torch input); with ``return_info`` also (iterations, flags) int32 arrays.

    2x2 matrices (the pairwise-Granger case) use the fused single-kernel path (``sc_wilson2``);
    S x S matrices up to S = 32 use the batched warp-per-matrix path and larger ones (up to 1024) the blocked
    Gauss-Jordan / tiled-GEMM path (both ``sc_wilson``).
    """
    lib = _lib.load()
    is_torch = isinstance(cross_spectral_matrix, torch.Tensor)
    csm = cross_spectral_matrix if is_torch else torch.from_numpy(np.ascontiguousarray(cross_spectral_matrix))
    if csm.ndim not in (3, 4):
        raise NotImplementedError("only (n_time_points, n_fft_samples, S, S) or (n_fft_samples, S, S) inputs: "
                                  "extra kept axes couple the convergence test across problems")
    n_sig = csm.shape[-1]
    if csm.shape[-2] != n_sig:
        raise ValueError("cross_spectral_matrix must be square in its last two dimensions")
    if n_sig > 1024:
        raise NotImplementedError("device Wilson factorisation handles up to 1024 x 1024 matrices")
    if not torch.cuda.is_available():
        raise RuntimeError("spectral_connectivity_b200 needs a CUDA device; there is no CPU fallback.")
    dev = torch.device("cuda", torch.cuda.current_device())
    shape = tuple(csm.shape)
    c = csm.to(dev).to(torch.complex128).reshape((-1,) + shape[-3:]).contiguous()
    nb, nfft = c.shape[0], c.shape[1]
    out = torch.empty_like(c)
    iters = torch.zeros(nb, dtype=torch.int32, device=dev)
    flags = torch.zeros(nb, dtype=torch.int32, device=dev)
    tw = twiddles(nfft, torch.complex128, dev)
    if n_sig == 2:
        ws_bytes = lib.sc_wilson_workspace_bytes(nfft)
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
        rc = lib.sc_wilson2(_lib.ptr(c), nb, nfft, float(tolerance), int(max_iterations), _lib.ptr(tw), _lib.ptr(out),
                            _lib.ptr(iters), _lib.ptr(flags), _lib.ptr(ws), ws_bytes, _lib.stream_ptr())
        _lib.check(rc, "sc_wilson2")
    else:
        ws_bytes = lib.sc_wilson_general_workspace_bytes(nb, nfft, n_sig)
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
        rc = lib.sc_wilson(_lib.ptr(c), nb, nfft, nfft, 0, n_sig, float(tolerance), int(max_iterations), _lib.ptr(tw),
                           _lib.ptr(out), _lib.ptr(iters), _lib.ptr(flags), _lib.ptr(ws), ws_bytes, _lib.stream_ptr())
        _lib.check(rc, "sc_wilson")
    n_bad = int((flags & _lib.FLAG_NOT_CONVERGED).ne(0).sum())
    if n_bad:
        logger.warning(f"Maximum iterations reached. {nb - n_bad} of {nb} converged")
    if int((flags & _lib.FLAG_NOT_SPD).ne(0).sum()):
        logger.warning("Computing the initial conditions using the Cholesky failed; those factors are NaN.")
    out = out.reshape(shape)
    if not is_torch:
        out = out.cpu().numpy()
        iters, flags = iters.cpu().numpy(), flags.cpu().numpy()
    return (out, iters, flags) if return_info else out
