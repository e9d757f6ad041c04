"""Where does the PLI-family error come from?  (a) the fp32 Fourier coefficients, (b) Im(x_i conj x_j) formed in
fp32, (c) the fp32 accumulation over observations.  Run on the GPU box."""
import sys
import numpy as np
import torch
sys.path.insert(0, '/root/repo')
from oracle import oracle as O
import spectral_connectivity_b200 as sc


def nerr(a, b):
    return float(np.nanmax(np.abs(a - b)) / np.nanmax(np.abs(b)))


def case(name, n_samples, T, S, fs, nw, dur, sub):
    x = O.synthetic_series(n_samples, T, S, fs, seed=9)
    m = sc.Multitaper(x, sampling_frequency=fs, time_halfbandwidth_product=nw, time_window_duration=dur)
    c = sc.Connectivity.from_multitaper(m)
    got = c.compute(["weighted_phase_lag_index", "debiased_squared_weighted_phase_lag_index", "phase_lag_index"])
    n, step, nfft = O.window_geometry(n_samples, fs, dur)
    k = O.default_n_tapers(nw)
    coef = O.multitaper_fft(x, fs, O.dpss_tapers(n, nw, k, fs), n, step, nfft)[..., :sub]
    fnn = nfft // 2 + 1
    dev_coef = m.fft()[:, :, :, :fnn, :sub].to(torch.complex128)      # device fp32 coefficients, promoted
    W = dev_coef.shape[0]
    X = dev_coef.reshape(W, -1, fnn, sub)                              # (W, TK, F, s)
    cs = X[..., :, None] * X[..., None, :].conj()                      # (W, TK, F, s, s)
    im = cs.imag
    idx = torch.arange(sub)
    im[..., idx, idx] = 0
    wpli64 = (im.mean(1) / im.abs().mean(1)).cpu().numpy()
    # fp32 Im, fp64 accumulation
    X32 = m.fft()[:, :, :, :fnn, :sub].reshape(W, -1, fnn, sub)
    im32 = (X32[..., :, None] * X32[..., None, :].conj()).imag
    im32[..., idx, idx] = 0
    im32 = im32.to(torch.float64)
    wpli_im32 = (im32.mean(1) / im32.abs().mean(1)).cpu().numpy()
    ref = O.weighted_phase_lag_index(coef)
    g = got["weighted_phase_lag_index"][..., :sub, :sub]
    off = ~np.eye(sub, dtype=bool)
    print(f"[{name}] wPLI: kernel vs oracle {nerr(g[..., off], ref[..., off]):.2e} | "
          f"fp64-from-device-coefs vs oracle {nerr(wpli64[..., off], ref[..., off]):.2e} | "
          f"kernel vs fp64-from-device-coefs {nerr(g[..., off], wpli64[..., off]):.2e} | "
          f"fp32-Im/fp64-acc vs fp64-from-device-coefs {nerr(wpli_im32[..., off], wpli64[..., off]):.2e}")
    refd = O.debiased_squared_weighted_phase_lag_index(coef)
    gd = got["debiased_squared_weighted_phase_lag_index"][..., :sub, :sub]
    nobs = X.shape[1]
    s_im, s_abs, s_sq = im.sum(1), im.abs().sum(1), (im ** 2).sum(1)
    d64 = ((s_im ** 2 - s_sq) / (s_abs ** 2 - s_sq)).cpu().numpy()
    print(f"[{name}] dwPLI: kernel vs oracle {nerr(gd[..., off], refd[..., off]):.2e} | fp64-from-device-coefs vs oracle "
          f"{nerr(d64[..., off], refd[..., off]):.2e} | kernel vs fp64-from-device-coefs {nerr(gd[..., off], d64[..., off]):.2e}")
    # where is the worst element?
    d = np.abs(g - ref); d[~np.isfinite(d)] = 0
    w = np.unravel_index(np.argmax(d), d.shape)
    print(f"[{name}] worst wPLI element {w}: got {g[w]:.7f} ref {ref[w]:.7f}; coherence there "
          f"{O.coherence_magnitude(coef)[w]:.4f}")
    refp = O.phase_lag_index(coef)
    gp = got["phase_lag_index"][..., :sub, :sub]
    print(f"[{name}] PLI kernel vs oracle {nerr(gp[..., off], refp[..., off]):.2e}; mismatching elements "
          f"{int((np.abs(gp - refp)[..., off] > 1e-6).sum())} of {gp[..., off].size}")


case("ragged S=70 T=3", 500, 3, 70, 250.0, 2, 0.4, 70)
case("ragged S=130 T=2", 500, 2, 130, 250.0, 2, 0.4, 64)
case("cfg3 window", 1000, 32, 128, 1000.0, 4, 1.0, 16)
