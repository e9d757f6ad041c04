"""Library-call REPLAY of the reference's CuPy backend (NOT CuPy, NOT the product).

``cupy`` cannot be installed in this image (no network), so BASELINE.json's ">= 10x the reference's own CuPy-backend
throughput" has no directly measurable denominator.  This module issues, with torch ops on the GPU, the same
sequence of library calls the reference makes when ``xp = cupy`` -- same shapes, same complex128 precision, the same
materialised intermediates and the same host synchronisation per Wilson iteration -- so the work lands in the same
cuFFT / cuBLAS / cuSOLVER / elementwise kernels CuPy would dispatch to (call sites C1-C13):

  C1  strided window gather + copy            transforms.py:1372-1374
  C2  per-window mean removal                 transforms.py:1860
  C3  broadcast taper product, materialised   transforms.py:1402-1404
  C4  fft(n=nfft, axis=-2) / fs               transforms.py:1405
  C5  k=1 batched matmul -> UN-AVERAGED CSM   connectivity.py:1799-1822
  C6  mean over trials x tapers               connectivity.py:67-75
  C7  coherency normalisation, |.|^2, clip    connectivity.py:649-657, 700-702
  C8-C12  Wilson loop per pair                minimum_phase_decomposition.py:227-322 (host sync per iteration)
  C13 transfer function / noise covariance    connectivity.py:1705, 1739-1748

Every number produced with it is labelled "library-call replay".  Used only by ``bench.py --impl replay``.
"""
import time
from itertools import combinations

import numpy as np
import torch

EPS = float(np.finfo(float).eps)
TIKHONOV = 1e-12


def multitaper_fft(x, tapers, n, step, nfft, fs):
    """x (N,T,S) float64 cuda, tapers (n,K) float64 cuda -> (W,T,K,nfft,S) complex128 (non-contiguous view, like
    the reference's swapaxes)."""
    n_win = int(np.floor(x.shape[0] / step - n / step + 1))
    xs = x.permute(1, 2, 0)                                          # _add_axes/moveaxis: time last
    win = xs.unfold(-1, n, step)[..., :n_win, :].permute(2, 0, 1, 3).contiguous()   # C1 (W,T,S,n) copy
    win = win - win.mean(dim=-1, keepdim=True)                                       # C2
    projected = win[..., None] * tapers[None, None, None, :, :]                      # C3 (W,T,S,n,K) materialised
    coef = torch.fft.fft(projected, n=nfft, dim=-2) / fs                             # C4
    return coef.transpose(2, -1)                                                     # (W,T,K,nfft,S)


def _cross_spectral_matrix(coef):
    a = coef[..., None]                                                              # (...,S,1)
    return torch.matmul(a, a.conj().transpose(-1, -2))                               # C5 (W,T,K,F,S,S)


def expectation_csm(coef):
    return _cross_spectral_matrix(coef).mean(dim=(1, 2))                             # C6


def power(coef):
    return (coef * coef.conj()).real.mean(dim=(1, 2))


def coherence_magnitude(coef):
    """connectivity.py:632-702: _power twice, expectation CSM, normalise, non-negative bins, |.|^2, clip."""
    nfft = coef.shape[-2]
    p1, p2 = power(coef), power(coef)
    norm = torch.sqrt(p1[..., :, None] * p2[..., None, :]).clamp_min(EPS)
    c = expectation_csm(coef) / norm
    s = c.shape[-1]
    idx = torch.arange(s, device=c.device)
    c[..., idx, idx] = float("nan")
    c = c[:, : nfft // 2 + 1]
    return (c.real ** 2 + c.imag ** 2).clamp(0, 1)


def _plus(b):
    nf, s = b.shape[-3], b.shape[-1]
    c = torch.fft.ifft(b, dim=-3)                                                    # C10
    c[..., 0, :, :] *= 0.5
    r, q = torch.tril_indices(s, s, offset=-1)
    c[..., 0, r, q] = 0
    c[..., (nf + 1) // 2:, :, :] = 0
    return torch.fft.fft(c, dim=-3)


def wilson(csm, tolerance=1e-8, max_iterations=60):
    """minimum_phase_decomposition.py:227-322 with the per-iteration host synchronisation of `xp.all(...)`."""
    lead = csm.shape[0]
    eye = torch.eye(csm.shape[-1], dtype=csm.dtype, device=csm.device)
    lag0 = torch.fft.ifft(csm, dim=-3)[..., 0:1, :, :].real                          # C8
    g = torch.linalg.cholesky(lag0).transpose(-1, -2).to(csm.dtype) * torch.ones_like(csm)
    converged = torch.zeros(lead, dtype=torch.bool, device=csm.device)
    its = 0
    for _ in range(max_iterations):
        old = g.clone()
        y = torch.linalg.solve(g, csm)                                               # C9
        b = torch.linalg.solve(g, y.conj().transpose(-1, -2)) + eye
        g = torch.matmul(g, _plus(b))                                                # C11
        g[converged] = old[converged]
        err = (g - old).abs().reshape(lead, -1).amax(dim=1)                          # C12
        converged = err < tolerance
        its += 1
        if bool(converged.all()):                                                    # host sync every iteration
            break
    return g, its


def pairwise_granger(coef, pairs=None):
    """connectivity.py:1161-1191, 2282-2340: expectation CSM, _power, Python loop over pairs."""
    nfft, s = coef.shape[-2], coef.shape[-1]
    csm = expectation_csm(coef)
    total_power = power(coef)[:, : nfft // 2 + 1]
    out = torch.full((csm.shape[0], nfft // 2 + 1, s, s), float("nan"), dtype=torch.float64, device=coef.device)
    iters = []
    for i, j in (pairs if pairs is not None else combinations(range(s), 2)):
        ix = torch.tensor([i, j], device=coef.device)
        sub = csm[..., ix[:, None], ix[None, :]]
        g, its = wilson(sub)
        iters.append(its)
        h0 = torch.fft.ifft(g, dim=-3).real[..., 0:1, :, :]                          # C13
        lam = TIKHONOV * (h0 * h0).mean()
        eye = torch.eye(2, dtype=h0.dtype, device=h0.device)
        h = torch.matmul(g, torch.linalg.solve(h0 + lam * eye, eye).to(g.dtype))[:, : nfft // 2 + 1]
        sigma = torch.matmul(h0[..., 0, :, :], h0[..., 0, :, :].transpose(-1, -2))
        var = torch.diagonal(sigma, dim1=-1, dim2=-2)[..., None]
        rot = var.transpose(-1, -2) - sigma ** 2 / var
        p = total_power[..., ix][..., None]
        intrinsic = p - rot[..., None, :, :] * (h.real ** 2 + h.imag ** 2)
        intrinsic[intrinsic == 0] = EPS
        gc = torch.log(p) - torch.log(intrinsic)
        gc[gc <= 0] = float("nan")
        out[..., ix[:, None], ix[None, :]] = gc
    idx = torch.arange(s, device=coef.device)
    out[..., idx, idx] = float("nan")
    return out, iters


def sample_step(x_host, tapers_host, n, step, nfft, fs, measures, device):
    """One replay pass from a HOST array to HOST results (like the reference under CuPy: xp.asarray on the way in,
    _asnumpy on the way out), timed with the wall clock around a device synchronisation."""
    torch.cuda.synchronize(device)
    t0 = time.perf_counter()
    x = torch.from_numpy(x_host).to(device)
    taps = torch.from_numpy(tapers_host).to(device)
    out = {}
    for name in measures:
        coef = multitaper_fft(x, taps, n, step, nfft, fs)        # from_multitaper is called once per method
        if name == "coherence_magnitude":
            out[name] = coherence_magnitude(coef).cpu().numpy()
        elif name == "pairwise_spectral_granger_prediction":
            gc, its = pairwise_granger(coef)
            out[name] = gc.cpu().numpy()
            out["_wilson_iterations"] = float(np.mean(its))
        else:
            raise ValueError(f"replay does not cover '{name}'")
        del coef
    torch.cuda.synchronize(device)
    return time.perf_counter() - t0, out
