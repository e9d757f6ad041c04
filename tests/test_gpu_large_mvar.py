"""GPU parity tests of the large-matrix (S > 32) Wilson / MVAR path (csrc/wilson_blocked.cuh): blocked
Gauss-Jordan inversion, tiled c128 GEMMs, lag-0 Cholesky -- BASELINE config 5 (512-channel DTF).
Everything goes through the C ABI; the checker is the NumPy oracle / numpy.linalg."""
import numpy as np
import pytest
import torch
from conftest import assert_parity, golden

from oracle import oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def sc():
    import spectral_connectivity_b200 as sc
    assert torch.cuda.is_available()
    return sc


def _lib():
    from spectral_connectivity_b200 import _lib
    return _lib, _lib.load()


@pytest.mark.parametrize("s", [33, 64, 100, 257, 300])
def test_blocked_inverse_vs_numpy(sc, s):
    """sc_mvar_inverse for S > 32 = shift + blocked in-place Gauss-Jordan with partial pivoting + unscramble."""
    L, lib = _lib()
    rng = np.random.default_rng(s)
    n = 3
    h = rng.standard_normal((n, s, s)) + 1j * rng.standard_normal((n, s, s))
    h[1, :, 0] *= 1e-3  # forces pivoting in the first panel
    h[2] = np.triu(h[2]) + np.eye(s) * 3  # triangular: no interchanges at all
    lam = 0.25
    ht = torch.from_numpy(h).cuda()
    out = torch.empty_like(ht)
    wsb = lib.sc_mvar_workspace_bytes(n, 1, s)
    assert wsb > 0
    ws = torch.empty(wsb, dtype=torch.uint8, device="cuda")
    L.check(lib.sc_mvar_inverse(L.ptr(ht), lam, None, n, s, L.ptr(out), L.ptr(ws), wsb, L.stream_ptr()), "sc_mvar_inverse")
    ref = np.linalg.inv(h + lam * np.eye(s))
    assert_parity(out.cpu().numpy(), ref, 1e-10, f"blocked inverse S={s}")
    # too small a workspace is refused, not overrun
    rc = lib.sc_mvar_inverse(L.ptr(ht), lam, None, n, s, L.ptr(out), L.ptr(ws), 16, L.stream_ptr())
    assert rc == -3


@pytest.mark.parametrize("s,nfo", [(40, 3), (130, 2)])
def test_blocked_transfer_vs_numpy(sc, s, nfo):
    L, lib = _lib()
    rng = np.random.default_rng(7 + s)
    nb, nf = 2, 4
    g = rng.standard_normal((nb, nf, s, s)) + 1j * rng.standard_normal((nb, nf, s, s))
    h0 = rng.standard_normal((nb, s, s)) + 4 * np.eye(s)
    lam = 1e-3
    gt, h0t = torch.from_numpy(g).cuda(), torch.from_numpy(h0).cuda()
    h = torch.empty((nb, nfo, s, s), dtype=torch.complex128, device="cuda")
    sig = torch.empty((nb, s, s), dtype=torch.float64, device="cuda")
    wsb = lib.sc_mvar_workspace_bytes(nb, 2, s)
    ws = torch.empty(wsb, dtype=torch.uint8, device="cuda")
    L.check(lib.sc_mvar_transfer(L.ptr(gt), L.ptr(h0t), lam, None, nb, nf, nfo, s, L.ptr(h), L.ptr(sig), L.ptr(ws), wsb,
                                 L.stream_ptr()), "sc_mvar_transfer")
    ref = g[:, :nfo] @ np.linalg.inv(h0 + lam * np.eye(s))[:, None]
    assert_parity(h.cpu().numpy(), ref, 1e-10, "H = G (H0 + lam I)^-1")
    assert_parity(sig.cpu().numpy(), h0 @ np.swapaxes(h0, -1, -2), 1e-12, "Sigma")


def _csm(n_signals, n_trials, n_win, n=64, fs=200.0, seed=5):
    x = O.synthetic_series(n * n_win, n_trials, n_signals, fs, seed=seed).astype(np.float32)
    taps = O.dpss_tapers(n, 3, 5, fs)
    coef = O.multitaper_fft(x.astype(np.float64), fs, taps, n, n, n)
    return x, O.expected_csm(coef, row_block=8)


@pytest.mark.parametrize("s,n_trials", [(40, 24), (70, 40)])
def test_wilson_blocked_vs_oracle(sc, s, n_trials):
    """minimum_phase_decomposition on S > 32 (two-sided path): same factor and iteration counts as the oracle."""
    _, csm = _csm(s, n_trials, 2)
    got, iters, flags = sc.minimum_phase_decomposition(csm, return_info=True)
    ref, ref_it = O.wilson(csm, return_iterations=True)
    assert not flags.any()
    assert np.array_equal(iters, ref_it), (iters, ref_it)
    assert_parity(got, ref, 1e-7, f"blocked Wilson S={s}")
    # (G G^H reproduces S only to ~10 % here, on the device exactly as in the oracle: the reference's stopping
    # rule looks at max |dG| and its lag-0 treatment leaves a slowly decaying residual)
    rec = got @ np.conj(np.swapaxes(got, -1, -2))
    rec_ref = ref @ np.conj(np.swapaxes(ref, -1, -2))
    assert np.abs(rec - rec_ref).max() / np.abs(csm).max() < 1e-6


MVAR = ["directed_transfer_function", "directed_coherence", "partial_directed_coherence",
        "generalized_partial_directed_coherence", "direct_directed_transfer_function"]


def test_mvar_large_vs_oracle_on_device_csm(sc):
    """Real-series (Hermitian half-spectrum) path, S = 48: every MVAR quantity against the oracle evaluated on
    the device's own expected CSM (isolates the blocked fp64 algebra from the fp32 CSM rounding), and the
    end-to-end result against the oracle on the same fp32 samples."""
    fs, n, s = 200.0, 64, 48
    x, csm_ref = _csm(s, 30, 3, n=n, fs=fs)
    m = sc.Multitaper(x, sampling_frequency=fs, time_halfbandwidth_product=3, time_window_duration=n / fs)
    c = sc.Connectivity.from_multitaper(m)
    csm_dev = np.asarray(c._expectation_cross_spectral_matrix()).astype(np.complex128)
    assert csm_dev.shape == csm_ref.shape
    g = O.wilson(csm_dev)
    keep = np.arange(n // 2 + 1)
    h = O.transfer_function(g)[..., keep, :, :]
    sigma = O.noise_covariance(g)
    a = O.mvar_fourier_coefficients(h)
    assert_parity(c._minimum_phase_factor, g[..., keep, :, :], 1e-6, "G")
    assert_parity(c._transfer_function, h, 1e-6, "H")
    assert_parity(c._noise_covariance, sigma, 1e-6, "Sigma")
    assert_parity(c._MVAR_Fourier_coefficients, a, 1e-6, "A")
    refs = {"directed_transfer_function": O.directed_transfer_function(h),
            "directed_coherence": O.directed_coherence(h, sigma),
            "partial_directed_coherence": O.partial_directed_coherence(a),
            "generalized_partial_directed_coherence": O.generalized_partial_directed_coherence(a, sigma),
            "direct_directed_transfer_function": O.direct_directed_transfer_function(h, a)}
    for name in MVAR:
        assert_parity(getattr(c, name)(), refs[name], 2e-6, name)
    assert int(c.last_wilson_flags.sum()) == 0
    # end to end against the oracle's own float64 CSM of the same samples (fp32 CSM rounding is amplified by
    # the conditioning of a 48 x 48 spectral matrix, hence the looser bound)
    h_ref, _ = O.mvar_transfer_function(csm_ref)
    assert_parity(c.directed_transfer_function(), O.directed_transfer_function(h_ref), 1e-5, "DTF end to end")


def test_mvar_streamed_equals_cached(sc, monkeypatch):
    """Above _MVAR_CACHE_BYTES the measures stream over window chunks (config 5 cannot keep 256 GB of factors)."""
    fs, n, s = 200.0, 64, 40
    x, _ = _csm(s, 24, 4, n=n, fs=fs)
    mk = lambda: sc.Connectivity.from_multitaper(
        sc.Multitaper(x, sampling_frequency=fs, time_halfbandwidth_product=3, time_window_duration=n / fs),
        max_chunk_bytes=1 << 20)
    cached = mk().partial_directed_coherence()
    c2 = mk()
    monkeypatch.setattr(type(c2), "_MVAR_CACHE_BYTES", 3 * 8 * 33 * s * s * 16 // 2)
    streamed = c2.partial_directed_coherence()
    assert getattr(c2, "_mvar_cache", None) is None
    assert_parity(streamed, cached, 1e-6, "streamed vs cached PDC")
    assert c2.last_wilson_iterations.numel() == 4


def _series_512(n=120, n_trials=128, s=512, seed=55):
    """Same recipe as tests/golden/make_golden.py::series_512 (white noise + lag-1 coupling, float32)."""
    rng = np.random.default_rng(seed)
    x = rng.standard_normal((n, n_trials, s)).astype(np.float32)
    x[1:, :, 1::2] += 0.5 * x[:-1, :, 0::2]
    return x


def test_dtf_512_channels_golden(sc):
    """BASELINE config 5 geometry (512 channels x 128 trials, 120-sample windows @ 2 kHz, 9 tapers), one window:
    directed_transfer_function against the LIVE REFERENCE's result (tests/golden/dtf512.npz, ~15 CPU-minutes
    there), plus the size-independent properties: converged, rows sum to 1, range [0, 1]."""
    import warnings
    x = _series_512()
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")  # "fewer time points than signals" (transforms.py:733-745)
        m = sc.Multitaper(x, sampling_frequency=2000.0, time_halfbandwidth_product=5, time_window_duration=0.060)
    c = sc.Connectivity.from_multitaper(m, output="torch")
    dtf = c.directed_transfer_function()
    assert dtf.shape == (1, 61, 512, 512)
    assert int(c.last_wilson_flags.sum()) == 0, (c.last_wilson_flags, c.last_wilson_iterations)
    assert 5 < int(c.last_wilson_iterations[0]) < 40
    assert bool(torch.isfinite(dtf).all())
    assert float(dtf.min()) >= 0 and float(dtf.max()) <= 1 + 1e-6
    assert float((dtf.sum(-1) - 1).abs().max()) < 1e-5
    g = golden("dtf512.npz")
    got = dtf.cpu().numpy()
    assert_parity(got[0, ::6, :32, :], g["rows"], 1e-5, "DTF 512 channels vs live reference")
    assert_parity(got.mean(axis=-2)[0, ::6], g["col_mean"], 1e-5, "DTF column means")


def test_wilson_diverging_problem_is_flagged_like_the_reference(sc):
    """On the SURVEY.md 8(d) recipe (common 40 Hz line) at 512 channels x 1152 observations the reference's
    iteration itself diverges (oracle: max |dG| grows to ~75 by iteration 4); the device reports
    NOT_CONVERGED for such windows instead of hanging or crashing."""
    import warnings
    fs, n, s, n_trials = 2000.0, 120, 512, 128
    x = O.synthetic_series(n, n_trials, s, fs, seed=55).astype(np.float32)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        m = sc.Multitaper(x, sampling_frequency=fs, time_halfbandwidth_product=5, time_window_duration=n / fs)
    c = sc.Connectivity.from_multitaper(m, output="torch")
    c._mvar(max_iterations=12)
    assert int(c.last_wilson_flags[0]) == 1 and int(c.last_wilson_iterations[0]) == 12


@pytest.mark.parametrize("s,n_trials", [(70, 12), (130, 20)])
def test_global_coherence_large_vs_oracle(sc, s, n_trials):
    """global_coherence above 64 signals: repeated squaring with the tiled c128 GEMM vs the oracle's SVD."""
    fs, n = 200.0, 32
    x = O.synthetic_series(2 * n, n_trials, s, fs, seed=77).astype(np.float32)
    m = sc.Multitaper(x, sampling_frequency=fs, time_halfbandwidth_product=2, time_window_duration=n / fs)
    c = sc.Connectivity.from_multitaper(m)
    val, vec = c.global_coherence()
    taps = O.dpss_tapers(n, 2, O.default_n_tapers(2), fs)
    coef = O.multitaper_fft(x.astype(np.float64), fs, taps, n, n, n)
    ref_val, ref_vec = O.global_coherence(coef)
    assert val.shape == ref_val.shape and vec.shape == ref_vec.shape
    assert_parity(val, ref_val, 1e-5, f"global coherence S={s}")
    overlap = np.abs(np.sum(np.conj(vec) * ref_vec, axis=-2))
    assert np.allclose(overlap, 1.0, atol=1e-3)
