"""NumPy float64 restatement of the reference hot path (TEST INFRASTRUCTURE).

This module is the *checker*, never the product: only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import it.  It restates, function by function,
what Eden-Kramer-Lab/spectral_connectivity computes on the path

    time series -> Multitaper.fft -> expectation of the cross-spectral matrix
    -> coherence / PLV / PLI family -> Wilson factorisation -> pairwise Granger

Every function cites the reference file:line it follows (paths relative to the
reference checkout).  Parity status: **pinned** -- ``tests/golden/make_golden.py``
imports the live reference in the build container, runs it on seeded inputs and
stores the outputs as fixtures; ``tests/test_oracle_golden.py`` checks this
module against those fixtures and against the known-answer vectors of the
reference's own tests (tests/test_connectivity.py:25-264,
tests/test_minimum_phase_decomposition.py:44-119, tests/test_transforms.py:39-59).

Third-party arithmetic on the path (not vendored by the reference):
scipy.fft (pocketfft) fft/ifft/next_fast_len, numpy.linalg.{solve,cholesky},
scipy.signal.windows.dpss as the DPSS stand-in (equal to the reference's
dpss_windows to <=1e-11).  Versions in the build image:
NumPy 2.3.5, SciPy 1.18.1 (the reference declares ranges only, pyproject.toml:42-47).
"""
from __future__ import annotations

from itertools import combinations

import numpy as np
from scipy.fft import fft, fftfreq, ifft, next_fast_len
from scipy.signal.windows import dpss as _scipy_dpss

EPS = np.finfo(float).eps
LOW_BIAS_EIGENVALUE = 0.9  # transforms.py:22
TIKHONOV = 1e-12  # connectivity.py:79

# axes of (window, trial, taper) reduced by each expectation_type, connectivity.py:67-75
EXPECTATION_AXES = {
    "time": (0,),
    "trials": (1,),
    "tapers": (2,),
    "time_trials": (0, 1),
    "time_tapers": (0, 2),
    "trials_tapers": (1, 2),
    "time_trials_tapers": (0, 1, 2),
}


# --------------------------------------------------------------------------- #
# window / frequency index arithmetic (must be bit exact)
# --------------------------------------------------------------------------- #
def window_geometry(n_samples, fs, duration=None, step_s=None, n_per_window=None,
                    n_per_step=None, n_fft=None):
    """Samples per window / per step / FFT length.

    transforms.py:1002-1023 (window: ``int(around(duration*fs))``),
    transforms.py:1051-1070 (step: ``int(step*fs)`` -- truncation, not rounding),
    transforms.py:1025-1036 (``next_fast_len`` default).
    """
    if duration is not None:
        n = int(np.around(duration * fs))
    elif n_per_window is not None:
        n = int(n_per_window)
    else:
        n = int(n_samples)
    if step_s is not None:
        step = int(step_s * fs)
    elif n_per_step is not None:
        step = int(n_per_step)
    else:
        step = n
    nfft = int(n_fft) if n_fft is not None else int(next_fast_len(n))
    return n, step, nfft


def n_windows(n_samples, n, step):
    """Window count, float arithmetic then floor (transforms.py:1363-1365)."""
    return int(np.floor((n_samples / step) - (n / step) + 1))


def frequencies(nfft, fs):
    """Two-sided bin frequencies (transforms.py:1038-1048)."""
    return fftfreq(nfft, 1.0 / fs)


def non_negative_frequencies(freqs):
    """First Nfft//2+1 bins with the Nyquist sign fix (connectivity.py:402-424)."""
    out = np.array(freqs[: len(freqs) // 2 + 1], dtype=float)
    if out.size and out[-1] < 0:
        out[-1] = abs(out[-1])
    return out


def window_times(n_samples, fs, n, step, start_time=0.0):
    """Window START times (transforms.py:1072-1091)."""
    w = n_windows(n_samples, n, step)
    return start_time + (np.arange(w) * step) / fs


# --------------------------------------------------------------------------- #
# tapers
# --------------------------------------------------------------------------- #
def default_n_tapers(time_halfbandwidth_product):
    """``floor(2*NW - 1)`` (transforms.py:979-993)."""
    return int(np.floor(2.0 * time_halfbandwidth_product - 1))


def dpss_tapers(n, time_halfbandwidth_product, n_tapers, fs, is_low_bias=True):
    """(n, K) tapers: unit l2 norm DPSS, reference sign convention, eigenvalue
    filter (> 0.9, else keep the best one), scaled by sqrt(fs).

    transforms.py:1408-1440 (_make_tapers), 1539-1613 (dpss_windows),
    1717-1745 (sign), 1758-1765 (low-bias filter).  SciPy's ``dpss`` is the
    nitime-compatible stand-in the reference's own test compares against
    (tests/test_transforms.py:271-284).
    """
    tapers, ratios = _scipy_dpss(n, time_halfbandwidth_product, int(n_tapers),
                                 norm=2, return_ratios=True)
    tapers = np.atleast_2d(tapers)
    ratios = np.atleast_1d(ratios)
    if is_low_bias:
        keep = ratios > LOW_BIAS_EIGENVALUE
        if not keep.any():
            keep = np.zeros_like(keep)
            keep[np.argmax(ratios)] = True
        tapers = tapers[keep]
    return tapers.T * np.sqrt(fs)


# --------------------------------------------------------------------------- #
# Multitaper.fft
# --------------------------------------------------------------------------- #
def sliding_windows(x, n, step):
    """(N,T,S) -> (W,T,S,n) copy (transforms.py:1311-1374)."""
    w = n_windows(x.shape[0], n, step)
    idx = (np.arange(w) * step)[:, None] + np.arange(n)[None, :]  # (W, n)
    return np.moveaxis(x[idx], 1, -1)  # (W,n,T,S) -> (W,T,S,n)


def detrend(windows, kind):
    """Per-window detrend along the last axis (transforms.py:1798-1915)."""
    if kind is None:
        return windows
    if kind in ("constant", "c"):
        return windows - windows.mean(axis=-1, keepdims=True)
    if kind in ("linear", "l"):
        n = windows.shape[-1]
        design = np.stack([np.arange(1, n + 1) / n, np.ones(n)], axis=1)  # (n,2), :1903-1906
        flat = windows.reshape(-1, n).T
        coef, *_ = np.linalg.lstsq(design, flat, rcond=None)
        return (flat - design @ coef).T.reshape(windows.shape)
    raise ValueError(kind)


def multitaper_fft(x, fs, tapers, n, step, nfft, detrend_type="constant"):
    """Fourier coefficients (W,T,K,Nfft,S), complex128, two-sided.

    transforms.py:1147-1171 (fft), 1377-1405 (_multitaper_fft): taper product,
    ``fft(n=Nfft)`` (zero-pads or crops), divide by fs.
    """
    x = np.asarray(x, dtype=float)
    if x.ndim == 1:
        x = x[:, None, None]
    elif x.ndim == 2:
        x = x[:, None, :]
    win = detrend(sliding_windows(x, n, step), detrend_type)  # (W,T,S,n)
    tapered = win[..., :, None] * tapers[None, None, None, :, :]  # (W,T,S,n,K)
    coef = fft(tapered, n=nfft, axis=-2) / fs  # (W,T,S,Nfft,K)
    return np.swapaxes(coef, 2, -1)  # (W,T,K,Nfft,S)


# --------------------------------------------------------------------------- #
# cross-spectral matrix + expectation
# --------------------------------------------------------------------------- #
def expectation(arr, expectation_type):
    """Mean over the (window, trial, taper) axes named (connectivity.py:67-75)."""
    return arr.mean(axis=EXPECTATION_AXES[expectation_type])


def n_observations(shape, expectation_type):
    """connectivity.py:594-610."""
    return int(np.prod([shape[a] for a in EXPECTATION_AXES[expectation_type]]))


def cross_spectral_matrix(coef):
    """Un-averaged X_i conj(X_j) per (w,t,k,f): (W,T,K,F,S,S)
    (connectivity.py:447-461, 1799-1822)."""
    return coef[..., :, None] * np.conj(coef[..., None, :])


def expected_csm(coef, expectation_type="trials_tapers", fcn=None, row_block=None):
    """E[fcn(X_i conj X_j)] (connectivity.py:463-526).

    ``row_block`` tiles the first signal index to bound the un-averaged tensor,
    the same purpose as the reference's ``blocks`` option (connectivity.py:490-524).
    """
    n_sig = coef.shape[-1]
    rb = n_sig if not row_block else int(row_block)
    rows = []
    for lo in range(0, n_sig, rb):
        hi = min(n_sig, lo + rb)
        block = coef[..., lo:hi, None] * np.conj(coef[..., None, :])
        if fcn is not None:
            block = fcn(block, lo)
        rows.append(expectation(block, expectation_type))
    return np.concatenate(rows, axis=-2)


def power(coef, expectation_type="trials_tapers"):
    """E[|X|^2], all bins (connectivity.py:441-445)."""
    return expectation((coef * np.conj(coef)).real, expectation_type)


def _nonneg(arr, axis):
    """connectivity.py:113-141: keep the first Nfft//2+1 bins."""
    nf = arr.shape[axis]
    return np.take(arr, np.arange(nf // 2 + 1), axis=axis)


def _pair_norm(p):
    norm = np.sqrt(p[..., :, None] * p[..., None, :])
    return np.maximum(norm, EPS)  # connectivity.py:649-652


def _nan_diagonal(a):
    s = a.shape[-1]
    a[..., np.arange(s), np.arange(s)] = np.nan
    return a


def coherency(coef, expectation_type="trials_tapers", row_block=None):
    """connectivity.py:632-657."""
    p = power(coef, expectation_type)
    c = expected_csm(coef, expectation_type, row_block=row_block) / _pair_norm(p)
    return _nonneg(_nan_diagonal(c), -3)


def coherence_magnitude(coef, expectation_type="trials_tapers", row_block=None):
    """connectivity.py:675-702."""
    return np.clip(np.abs(coherency(coef, expectation_type, row_block)) ** 2, 0, 1)


def coherence_phase(coef, expectation_type="trials_tapers", row_block=None):
    """connectivity.py:659-673."""
    return np.angle(coherency(coef, expectation_type, row_block))


def imaginary_coherence(coef, expectation_type="trials_tapers", row_block=None):
    """connectivity.py:704-743."""
    p = power(coef, expectation_type)
    c = expected_csm(coef, expectation_type, row_block=row_block).imag / _pair_norm(p)
    return _nonneg(np.clip(np.abs(c), 0, 1), -3)


def _imag_zero_diag(block, row_lo):
    """Imaginary part with the (i,i) entries forced to 0 (connectivity.py:970-978)."""
    im = block.imag.copy()
    rows = np.arange(block.shape[-2])
    cols = rows + row_lo
    ok = cols < block.shape[-1]
    im[..., rows[ok], cols[ok]] = 0
    return im


def phase_locking_value_complex(coef, expectation_type="trials_tapers", row_block=None):
    """E[x/|x|] (connectivity.py:897-903)."""
    with np.errstate(invalid="ignore", divide="ignore"):
        return _nonneg(expected_csm(coef, expectation_type,
                                    lambda b, lo: b / np.abs(b), row_block), -3)


def phase_locking_value(coef, expectation_type="trials_tapers", row_block=None):
    """connectivity.py:905-931."""
    return np.abs(phase_locking_value_complex(coef, expectation_type, row_block))


def phase_lag_index(coef, expectation_type="trials_tapers", row_block=None):
    """E[sign Im] (connectivity.py:933-982)."""
    return _nonneg(expected_csm(coef, expectation_type,
                                lambda b, lo: np.sign(_imag_zero_diag(b, lo)), row_block).real, -3)


def weighted_phase_lag_index(coef, expectation_type="trials_tapers", row_block=None):
    """E[Im]/E[|Im|], weights below eps -> 1 (connectivity.py:984-1028)."""
    w = expected_csm(coef, expectation_type, lambda b, lo: np.abs(_imag_zero_diag(b, lo)), row_block)
    w[w < EPS] = 1
    num = expected_csm(coef, expectation_type, _imag_zero_diag, row_block)
    return _nonneg(num / w, -3)


def debiased_squared_phase_lag_index(coef, expectation_type="trials_tapers", row_block=None):
    """connectivity.py:1030-1058."""
    n = n_observations(coef.shape, expectation_type)
    return (n * phase_lag_index(coef, expectation_type, row_block) ** 2 - 1.0) / (n - 1.0)


def debiased_squared_weighted_phase_lag_index(coef, expectation_type="trials_tapers",
                                              row_block=None):
    """connectivity.py:1060-1127 (zero weights -> NaN)."""
    n = n_observations(coef.shape, expectation_type)
    s_im = expected_csm(coef, expectation_type, _imag_zero_diag, row_block) * n
    s_sq = expected_csm(coef, expectation_type,
                        lambda b, lo: _imag_zero_diag(b, lo) ** 2, row_block) * n
    s_abs = expected_csm(coef, expectation_type,
                         lambda b, lo: np.abs(_imag_zero_diag(b, lo)), row_block) * n
    w = s_abs ** 2 - s_sq
    w[w == 0] = np.nan
    return _nonneg((s_im ** 2 - s_sq) / w, -3)


def pairwise_phase_consistency(coef, expectation_type="trials_tapers", row_block=None):
    """connectivity.py:1129-1159."""
    n = n_observations(coef.shape, expectation_type)
    s = phase_locking_value_complex(coef, expectation_type, row_block) * n
    return ((s * np.conj(s) - n) / (n ** 2 - n)).real


# --------------------------------------------------------------------------- #
# Wilson spectral factorisation
def phase_slope_index(coef, freqs, frequencies_of_interest=None, frequency_resolution=None,
                      expectation_type="trials_tapers"):
    """Imaginary part of sum_{f1<f2} conj(c[f1]) c[f2] over the (band-passed, subsampled) non-negative
    frequencies of the coherency (connectivity.py:1587-1650; _bandpass :2040-2073 keeps lo < f < hi,
    _get_independent_frequency_step :2076-2100, _inner_combination :1652-1676)."""
    c = coherency(coef, expectation_type)
    f = non_negative_frequencies(np.asarray(freqs))
    if frequencies_of_interest is not None:
        keep = (frequencies_of_interest[0] < f) & (f < frequencies_of_interest[1])
        c = c[..., keep, :, :]
    step = 1 if frequency_resolution is None else int(np.ceil(frequency_resolution / (freqs[1] - freqs[0])))
    c = c[..., np.arange(0, c.shape[-3], step), :, :]
    i1, i2 = np.array(list(combinations(range(c.shape[-3]), 2))).T
    return (np.conj(c[..., i1, :, :]) * c[..., i2, :, :]).sum(axis=-3).imag


# --------------------------------------------------------------------------- #
def _herm(a):
    return np.conj(np.swapaxes(a, -1, -2))


def wilson_initial(csm):
    """Cholesky factor (transposed) of the real lag-0 covariance, broadcast over
    frequency (minimum_phase_decomposition.py:48-93).  The reference's random
    fallback for a non-SPD lag-0 matrix is not reproducible and is not restated:
    this raises ``numpy.linalg.LinAlgError`` instead."""
    lag0 = ifft(csm, axis=-3)[..., 0:1, :, :].real
    return np.swapaxes(np.linalg.cholesky(lag0), -1, -2)


def plus_operator(b):
    """Causal projection (minimum_phase_decomposition.py:96-142)."""
    nf, s = b.shape[-3], b.shape[-1]
    c = ifft(b, axis=-3)
    c[..., 0, :, :] *= 0.5
    r, q = np.tril_indices(s, k=-1)
    c[..., 0, r, q] = 0
    c[..., (nf + 1) // 2:, :, :] = 0
    return fft(c, axis=-3)


def wilson(csm, tolerance=1e-8, max_iterations=60, return_iterations=False):
    """Minimum-phase factor G with S = G G^H (minimum_phase_decomposition.py:227-322).

    Leading index 0 is the unit of convergence: an index is frozen at its first
    iterate whose max |dG| falls below ``tolerance`` (:310-315)."""
    csm = np.asarray(csm)
    lead = csm.shape[0]
    s = csm.shape[-1]
    eye = np.eye(s)
    g = np.zeros(csm.shape, dtype=complex)
    g[...] = wilson_initial(csm)
    frozen = np.zeros(lead, dtype=bool)
    iters = np.zeros(lead, dtype=int)
    for _ in range(max_iterations):
        prev = g
        y = np.linalg.solve(prev, csm)
        b = np.linalg.solve(prev, _herm(y)) + eye  # :218-224
        g = prev @ plus_operator(b)
        g[frozen] = prev[frozen]
        iters[~frozen] += 1
        err = np.abs((g - prev).reshape(lead, -1)).max(axis=1)  # :177-181
        frozen = err < tolerance
        if frozen.all():
            break
    return (g, iters) if return_iterations else g


# --------------------------------------------------------------------------- #
# spectral Granger
# --------------------------------------------------------------------------- #
def transfer_function(g):
    """H = G (H0 + lam I)^-1 with H0 = Re ifft(G)[lag 0], lam = 1e-12*mean(H0^2)
    over ALL leading indices (connectivity.py:1712-1748)."""
    h0 = ifft(g, axis=-3).real[..., 0:1, :, :]
    lam = TIKHONOV * np.mean(h0 * h0)
    eye = np.eye(h0.shape[-1])
    return g @ np.linalg.solve(h0 + lam * eye, eye)


def noise_covariance(g):
    """H0 H0^T (connectivity.py:1679-1709)."""
    h0 = ifft(g, axis=-3).real[..., 0, :, :]
    return h0 @ np.swapaxes(h0, -1, -2)


def rotated_covariance(sigma):
    """R[a,b] = Sigma_bb - Sigma_ab^2 / Sigma_aa (connectivity.py:1825-1848)."""
    var = np.diagonal(sigma, axis1=-1, axis2=-2)[..., None]
    return np.swapaxes(var, -1, -2) - sigma ** 2 / var


def pairwise_granger(csm, total_power, pairs=None, tolerance=1e-8, max_iterations=60,
                     return_iterations=False):
    """(..., Fnn, S, S) with [i, j] = influence j -> i; NaN diagonal and NaN where
    the log-ratio is <= 0 (connectivity.py:1161-1191, 2282-2340, 1751-1779)."""
    nf = total_power.shape[-2]
    keep = np.arange(nf // 2 + 1)
    p = np.take(total_power, keep, axis=-2)
    s = csm.shape[-1]
    shape = list(csm.shape)
    shape[-3] = keep.size
    out = np.full(shape, np.nan)
    if pairs is None:
        pairs = combinations(range(s), 2)
    it_log = []
    for i, j in pairs:
        ix = np.array([i, j])
        sub = csm[..., ix[:, None], ix[None, :]]
        try:
            g, its = wilson(sub, tolerance, max_iterations, return_iterations=True)
        except np.linalg.LinAlgError:
            continue
        it_log.append(its)
        h = transfer_function(g)[..., keep, :, :]
        rot = rotated_covariance(noise_covariance(g))
        with np.errstate(invalid="ignore", divide="ignore"):
            intrinsic = p[..., ix][..., None] - rot[..., None, :, :] * np.abs(h) ** 2
            intrinsic[intrinsic == 0] = EPS
            gc = np.log(p[..., ix][..., None]) - np.log(intrinsic)
            gc[gc <= 0] = np.nan
        out[..., ix[:, None], ix[None, :]] = gc
    out[..., np.arange(s), np.arange(s)] = np.nan
    return (out, it_log) if return_iterations else out


# --------------------------------------------------------------------------- #
# MVAR family from the full-matrix Wilson factor
# --------------------------------------------------------------------------- #
def mvar_transfer_function(csm):
    """Non-negative-frequency H from the full S x S factor (connectivity.py:567-574)."""
    g = wilson(csm)
    return _nonneg(transfer_function(g), -3), noise_covariance(g)


def mvar_fourier_coefficients(h):
    """Tikhonov-regularised inverse of H per (w, f) (connectivity.py:580-588)."""
    lam = TIKHONOV * np.mean(np.real(np.conj(h) * h))
    eye = np.eye(h.shape[-1], dtype=h.dtype)
    return np.linalg.solve(h + lam * eye, eye)


def _noise_variance(sigma):
    return np.diagonal(sigma, axis1=-1, axis2=-2)[..., None, :, None]  # connectivity.py:1904-1925


def directed_transfer_function(h):
    """connectivity.py:1237-1266."""
    inflow = np.sqrt(np.sum(np.abs(h) ** 2, axis=-1, keepdims=True))
    return np.abs(h / inflow) ** 2


def directed_coherence(h, sigma):
    """connectivity.py:1268-1296."""
    nv = _noise_variance(sigma)
    inflow = np.sqrt(np.sum(nv * np.abs(h) ** 2, axis=-1, keepdims=True))
    return np.sqrt(nv) * np.abs(h) ** 2 / inflow


def partial_directed_coherence(a):
    """connectivity.py:1298-1343."""
    outflow = np.sqrt(np.sum(np.abs(a) ** 2, axis=-2, keepdims=True))
    return np.abs(a / outflow) ** 2


def generalized_partial_directed_coherence(a, sigma):
    """connectivity.py:1345-1380."""
    nv = _noise_variance(sigma)
    outflow = np.sqrt(np.sum(np.abs(a) ** 2 / nv, axis=-2, keepdims=True))
    return np.abs(a / np.sqrt(nv) / outflow) ** 2


def direct_directed_transfer_function(h, a):
    """connectivity.py:1382-1426 (inflow summed over sources AND frequencies)."""
    inflow = np.sqrt(np.sum(np.abs(h) ** 2, axis=(-1, -3), keepdims=True))
    return np.abs(h / inflow) * np.sqrt(partial_directed_coherence(a))


# --------------------------------------------------------------------------- #
# SVD-based measures
# --------------------------------------------------------------------------- #
def _obs_matrix(coef):
    """(W,T,K,F,S) -> (W,F,S,T*K) (connectivity.py:1953-1976)."""
    w, t, k, f, s = coef.shape
    return np.moveaxis(coef.reshape(w, t * k, f, s), 1, -1)


def canonical_coherence(coef, group_labels):
    """(W, Fnn, G, G) squared canonical coherence between signal groups, NaN diagonal, and the
    sorted labels (connectivity.py:745-820, 1979-2032)."""
    group_labels = np.asarray(group_labels)
    labels = np.unique(group_labels)
    nf = coef.shape[-2]
    half = coef[..., : nf // 2 + 1, :]
    whitened = []
    for lab in labels:
        u, _, vh = np.linalg.svd(_obs_matrix(half[..., group_labels == lab]), full_matrices=False)
        whitened.append(u @ vh)
    n_g = len(labels)
    out = np.full(half.shape[:1] + (half.shape[-2], n_g, n_g), np.nan)
    for a, b in combinations(range(n_g), 2):
        cross = whitened[a] @ np.conj(np.swapaxes(whitened[b], -1, -2))
        sv = np.linalg.svd(cross, compute_uv=False)[..., 0]
        out[..., a, b] = out[..., b, a] = np.abs(sv) ** 2
    return out, labels


def global_coherence(coef, max_rank=1):
    """The ``max_rank`` largest eigenvalues of the per-(window, frequency) cross-spectral matrix = singular values
    squared / n_observations, over ALL Nfft bins, and their eigenvectors (each defined up to a phase)
    (connectivity.py:822-895, 2245-2279).  Ordering as the reference produces it: descending from the dense SVD
    when max_rank >= n_signals - 1 (:2258-2266), ASCENDING from scipy's ``svds`` otherwise (:2267-2276)."""
    x = _obs_matrix(coef)
    u, sv, _ = np.linalg.svd(x, full_matrices=False)
    val, vec = sv[..., :max_rank] ** 2 / x.shape[-1], u[..., :, :max_rank]
    if max_rank < x.shape[-2] - 1:
        val, vec = val[..., ::-1], vec[..., :, ::-1]
    return val, vec


# --------------------------------------------------------------------------- #
# deterministic synthetic workloads
# --------------------------------------------------------------------------- #
CONFIGS = {
    1: dict(N=1000, T=4, S=8, fs=500.0, NW=2.0, duration=None),
    2: dict(N=10_000, T=16, S=64, fs=1000.0, NW=3.0, duration=1.0),
    3: dict(N=30_000, T=32, S=128, fs=1000.0, NW=4.0, duration=1.0),
    4: dict(N=60_000, T=64, S=256, fs=1000.0, NW=4.0, duration=1.0),
    5: dict(N=120_000, T=128, S=512, fs=2000.0, NW=5.0, duration=0.060),
}


def synthetic_series(n_samples, n_trials, n_signals, fs, seed, dtype=np.float64):
    """Noise + lag-1 even->odd channel coupling + shared 40 Hz line with a
    per-channel phase."""
    rng = np.random.default_rng(seed)
    x = rng.standard_normal((n_samples, n_trials, n_signals))
    n_odd = x[:, :, 1::2].shape[-1]
    x[1:, :, 1::2] += 0.5 * x[:-1, :, 0:2 * n_odd:2]
    t = np.arange(n_samples) / fs
    phase = 2 * np.pi * np.arange(n_signals) / n_signals
    x += 0.5 * np.sin(2 * np.pi * 40.0 * t[:, None, None] + phase[None, None, :])
    return x.astype(dtype, copy=False)
