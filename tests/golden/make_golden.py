"""Generate golden fixtures by running the LIVE reference (build container only).

Usage (from the repo root, in the build container where /root/reference exists):

    python tests/golden/make_golden.py

Imports the reference's hot-path modules through the shim documented in
SURVEY.md (appendix) -- the package's ``__init__`` needs xarray, the three
modules below only need numpy/scipy -- runs them on seeded inputs and writes
``tests/golden/*.npz``.  The fixtures travel to the GPU box; the reference does
not.  Nothing here is imported by the product.
"""
import os
import sys
import types

import numpy as np

REF = os.environ.get("SC_REFERENCE_PATH", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))


def load_reference():
    pkg = types.ModuleType("spectral_connectivity")
    pkg.__path__ = [os.path.join(REF, "spectral_connectivity")]
    sys.modules["spectral_connectivity"] = pkg
    import spectral_connectivity.connectivity as C
    import spectral_connectivity.minimum_phase_decomposition as M
    import spectral_connectivity.transforms as T
    return T, C, M


def series(seed, n, t, s, fs):
    sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
    from oracle.oracle import synthetic_series
    return synthetic_series(n, t, s, fs, seed)


def psi_section(T, C):
    """phase_slope_index (connectivity.py:1587-1650) on the section-4 input: whole band, a band of
    interest, and a band with a frequency resolution (independent-frequency subsampling)."""
    psi = {}
    x = series(7, 300, 4, 4, 100.0)
    m = T.Multitaper(x, sampling_frequency=100.0, time_halfbandwidth_product=3, time_window_duration=1.0)
    c = C.Connectivity.from_multitaper(m)
    psi["x"] = x
    psi["meta"] = np.array([100.0, 3.0, 1.0])
    psi["all"] = np.asarray(c.phase_slope_index())
    psi["band"] = np.asarray(c.phase_slope_index(frequencies_of_interest=[5.0, 30.0]))
    psi["band_res"] = np.asarray(c.phase_slope_index(frequencies_of_interest=[2.0, 45.0], frequency_resolution=3.5))
    np.savez_compressed(os.path.join(HERE, "psi.npz"), **psi)


def series_512(n=120, n_trials=128, s=512, seed=55):
    """White noise + lag-1 even->odd coupling (no common sinusoid: with 1152 observations for 512 channels the
    reference's Wilson iteration diverges on the SURVEY.md 8(d) recipe, whose 40 Hz line makes the CSM
    ill-conditioned).  float32-representable so that the device sees the same samples."""
    rng = np.random.default_rng(seed)
    x = rng.standard_normal((n, n_trials, s)).astype(np.float32)
    x[1:, :, 1::2] += 0.5 * x[:-1, :, 0::2]
    return x


def dtf512_section(T, C):
    """BASELINE config-5 geometry on ONE window: 512 channels x 128 trials, 120 samples @ 2 kHz, NW = 5 (9 tapers);
    directed_transfer_function through the live reference (blocks=32 keeps the un-averaged CSM at 0.6 GB per
    block; two full-matrix Wilson factorisations of 120 x 512 x 512).  Takes ~15 minutes; only a slice is kept."""
    x = series_512().astype(np.float64)
    m = T.Multitaper(x, sampling_frequency=2000.0, time_halfbandwidth_product=5, time_window_duration=0.060)
    c = C.Connectivity.from_multitaper(m, blocks=32)
    dtf = np.asarray(c.directed_transfer_function())
    assert dtf.shape == (1, 61, 512, 512), dtf.shape
    np.savez_compressed(os.path.join(HERE, "dtf512.npz"), rows=dtf[0, ::6, :32, :],
                        row_sums=dtf.sum(axis=-1)[0], col_mean=dtf.mean(axis=-2)[0, ::6])


def baseline_configs_section(T, C):
    """BASELINE.json configs[1] in full (64 ch x 16 trials x 10 s @ 1 kHz, NW 3: power + coherency) and configs[2]
    on 3 of its 30 windows (128 ch x 32 trials @ 1 kHz, NW 4: expected CSM) through the
    live reference; float32-representable input (SURVEY.md 8d recipe), only slices are kept."""
    out = {}
    x2 = series(20261017 + 2, 10_000, 16, 64, 1000.0).astype(np.float32).astype(np.float64)
    m = T.Multitaper(x2, sampling_frequency=1000.0, time_halfbandwidth_product=3, time_window_duration=1.0)
    c = C.Connectivity.from_multitaper(m, blocks=8)
    power = np.asarray(c.power())
    coh = np.asarray(c.coherency())
    assert power.shape == (10, 501, 64) and coh.shape == (10, 501, 64, 64)
    out["cfg2_power"] = power[::3, ::7]
    out["cfg2_coherency"] = coh[::3, ::25, :8, :]
    x3 = series(20261017 + 3, 3_000, 32, 128, 1000.0).astype(np.float32).astype(np.float64)
    m = T.Multitaper(x3, sampling_frequency=1000.0, time_halfbandwidth_product=4, time_window_duration=1.0)
    c = C.Connectivity.from_multitaper(m, blocks=16)
    csm = np.asarray(c._expectation_cross_spectral_matrix())
    assert csm.shape == (3, 1000, 128, 128)
    out["cfg3_csm"] = csm[:, ::100, :8, :]
    # (weighted_phase_lag_index cannot be generated here: with `blocks` the reference's own diagonal masking
    # raises IndexError (connectivity.py:1015), without blocks the un-averaged CSM of this geometry is 176 GB)
    np.savez_compressed(os.path.join(HERE, "baseline_configs.npz"), **out)


CFG4_PAIRS = [(0, 1), (2, 5), (10, 200), (3, 255), (100, 101), (17, 18)]


def config4_section(T, C):
    """BASELINE.json configs[3] (the headline workload: 256 ch x 64 trials @ 1 kHz, NW 4 -> 7 tapers, 1 s windows)
    on ONE window through the live reference: coherence_magnitude and pairwise spectral Granger prediction among
    the 12 channels that the six pairs below touch.  Both are pairwise measures (a pair's value depends on its two
    channels only), and the reference cannot evaluate them on all 256 channels here: its subset-Granger path
    allocates the un-averaged (1, 64, 7, 1000, 256, 256) CSM = 438 GB (connectivity.py:551)."""
    x = series(20261017 + 4, 1_000, 64, 256, 1000.0).astype(np.float32).astype(np.float64)
    ij = np.array(CFG4_PAIRS)
    chans = np.unique(ij)
    m = T.Multitaper(x[:, :, chans], sampling_frequency=1000.0, time_halfbandwidth_product=4, time_window_duration=1.0)
    c = C.Connectivity.from_multitaper(m)
    coh = np.asarray(c.coherence_magnitude())
    gc = np.asarray(c.pairwise_spectral_granger_prediction())
    assert coh.shape == (1, 501, chans.size, chans.size) and gc.shape == coh.shape
    np.savez_compressed(os.path.join(HERE, "config4.npz"), channels=chans, pairs=ij, coherence=coh[0, ::5],
                        granger=gc[0])


def round2_section(T, C):
    """Round-2 fixtures, all from the live reference:
    * ``cfg3_*``: BASELINE.json configs[2]'s own measures (weighted / plain / debiased phase lag index) on window 0 of
      its recording, 16 channels (pairwise measures: a channel subset equals that block of the full result; the
      un-averaged (1,32,7,1000,16,16) CSM is 0.9 GB, all 128 channels would need 59 GB per pass);
    * ``cfg5_canonical``: configs[4]'s canonical_coherence at ITS grouping (8 groups x 64 channels, 1152
      observations) on one 60 ms window of the 512-channel recipe of ``series_512``;
    * canonical coherence with a 70-signal group and with a rank-deficient group (more signals than observations),
      global coherence with max_rank = 1, 2, 3 (dense-SVD and svds branches);
    * pairwise Granger and DTF for EVERY expectation type (the reference treats ``csm.shape[0]`` as the unit of
      convergence whatever that axis is, minimum_phase_decomposition.py:290, 310-315)."""
    out = {}
    x3 = series(20261017 + 3, 1_000, 32, 128, 1000.0).astype(np.float32).astype(np.float64)
    m = T.Multitaper(x3[:, :, :16], sampling_frequency=1000.0, time_halfbandwidth_product=4, time_window_duration=1.0)
    c = C.Connectivity.from_multitaper(m)
    for meth in ("weighted_phase_lag_index", "phase_lag_index", "debiased_squared_weighted_phase_lag_index",
                 "debiased_squared_phase_lag_index"):
        out[f"cfg3_{meth}"] = np.asarray(getattr(c, meth)())
    x5 = series_512().astype(np.float64)
    m = T.Multitaper(x5, sampling_frequency=2000.0, time_halfbandwidth_product=5, time_window_duration=0.060)
    c = C.Connectivity.from_multitaper(m)
    cc, lab = c.canonical_coherence(np.arange(512) // 64)
    out["cfg5_canonical"] = np.asarray(cc)
    out["cfg5_canonical_labels"] = np.asarray(lab)
    # groups of 70 + 10 signals, 8 trials x 3 tapers = 24 observations (< 70: rank deficient) and 40 x 3 = 120 (full rank)
    for tag, n_trials in (("rankdef", 8), ("big", 40)):
        xg = series(31, 200, n_trials, 80, 100.0)
        m = T.Multitaper(xg, sampling_frequency=100.0, time_halfbandwidth_product=2, time_window_duration=1.0)
        c = C.Connectivity.from_multitaper(m)
        labels = np.where(np.arange(80) < 70, 0, 1)
        cc, lab = c.canonical_coherence(labels)
        out[f"canon_{tag}"] = np.asarray(cc)
    x6 = series(9, 300, 5, 6, 100.0)
    m6 = T.Multitaper(x6, sampling_frequency=100.0, time_halfbandwidth_product=3, time_window_duration=1.0)
    c6 = C.Connectivity.from_multitaper(m6)
    for rank in (1, 2, 3):
        val, vec = c6.global_coherence(max_rank=rank)
        out[f"global_rank{rank}_values"] = np.asarray(val)
        out[f"global_rank{rank}_vectors"] = np.asarray(vec)
    x4 = series(7, 300, 4, 4, 100.0)
    m4 = T.Multitaper(x4, sampling_frequency=100.0, time_halfbandwidth_product=3, time_window_duration=1.0)
    coef = np.asarray(m4.fft())
    for et in ["trials_tapers", "trials", "tapers", "time", "time_trials", "time_tapers", "time_trials_tapers"]:
        c = C.Connectivity(coef, expectation_type=et, frequencies=m4.frequencies, time=m4.time)
        out[f"granger__{et}"] = np.asarray(c.pairwise_spectral_granger_prediction())
        out[f"dtf__{et}"] = np.asarray(c.directed_transfer_function())
    np.savez_compressed(os.path.join(HERE, "round2.npz"), **out)
    for k, v in out.items():
        print(k, v.shape)


def delay_section(T, C):
    """delay / group_delay (connectivity.py:1428-1585) plus the statistics helpers they use
    (statistics.py:21-59, 147-203).  Channel 1 is a 3-sample delayed copy of channel 0 plus noise, so the pair
    is strongly coherent -- and the reference still finds nothing significant (n_obs2 = 0 default, see
    spectral_connectivity_b200/_statistics.py): the fixtures record exactly that."""
    import warnings
    import spectral_connectivity.statistics as S
    x = series(7, 600, 8, 4, 200.0)
    x[3:, :, 1] += 0.8 * x[:-3, :, 0]
    m = T.Multitaper(x, sampling_frequency=200.0, time_halfbandwidth_product=3, time_window_duration=1.0)
    c = C.Connectivity.from_multitaper(m)
    d = {"x": x, "meta": np.array([200.0, 3.0, 1.0])}
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        d["delay_band"] = np.asarray(c.delay(frequencies_of_interest=[5.0, 60.0]))
        d["delay_res"] = np.asarray(c.delay(frequencies_of_interest=[5.0, 60.0], frequency_resolution=4.0, n_range=2))
        for k, a in zip(("gd_delay", "gd_slope", "gd_r"), c.group_delay(frequencies_of_interest=[5.0, 60.0])):
            d[k] = np.asarray(a)
        rng = np.random.default_rng(3)
        coh1 = 0.9 * rng.random((5, 40, 6)) * np.exp(2j * np.pi * rng.random((5, 40, 6)))
        coh2 = 0.9 * rng.random((5, 40, 6)) * np.exp(2j * np.pi * rng.random((5, 40, 6)))
        d["stat_coh1"], d["stat_coh2"] = coh1, coh2
        d["stat_z_default"] = S.coherence_fisher_z_transform(coh1.copy(), 50)
        d["stat_z_two"] = S.coherence_fisher_z_transform(coh1.copy(), 50, coh2.copy(), 70)
        p = S.get_normal_distribution_p_values(d["stat_z_two"])
        d["stat_p"] = p
        d["stat_bh"] = S.Benjamini_Hochberg_procedure(p, alpha=0.2)
        d["stat_bonf"] = S.Bonferroni_correction(p, alpha=0.2)
        d["stat_groups"] = np.apply_along_axis(C._find_largest_independent_group, -2, p < 0.3, 2, 3)
    np.savez_compressed(os.path.join(HERE, "delay.npz"), **d)


def main():
    T, C, M = load_reference()
    if len(sys.argv) > 1 and sys.argv[1] == "delay":
        delay_section(T, C)
        return
    if len(sys.argv) > 1 and sys.argv[1] == "round2":
        round2_section(T, C)
        return
    if len(sys.argv) > 1 and sys.argv[1] == "config4":
        config4_section(T, C)
        return
    if len(sys.argv) > 1 and sys.argv[1] == "configs":
        baseline_configs_section(T, C)
        return
    if len(sys.argv) > 1 and sys.argv[1] == "psi":
        psi_section(T, C)
        return
    if len(sys.argv) > 1 and sys.argv[1] == "dtf512":
        dtf512_section(T, C)
        return
    np.random.seed(42)

    # ---- 1. index arithmetic (bit exact) ---------------------------------
    rows = []
    cases = [
        (1000, 500.0, None, None), (10_000, 1000.0, 1.0, None), (10_000, 1000.0, 1.0, 0.5),
        (120_000 // 50, 2000.0, 0.060, None), (120_000 // 50, 2000.0, 0.060, 0.030),
        (5000, 1000.0, 0.3, 0.1), (5000, 1500.0, 0.1234, 0.0371), (4097, 256.0, 1.001, 0.333),
        (3000, 1000.0 / 3.0, 0.75, 0.3), (777, 123.0, 0.5, 0.07), (2048, 1024.0, 0.0625, 0.03125),
    ]
    geom = {}
    for ci, (n, fs, dur, step) in enumerate(cases):
        m = T.Multitaper(np.zeros((n, 1, 1)), sampling_frequency=fs, time_window_duration=dur,
                         time_window_step=step, time_halfbandwidth_product=2)
        rows.append([n, fs, -1 if dur is None else dur, -1 if step is None else step,
                     m.n_time_samples_per_window, m.n_time_samples_per_step, m.n_fft_samples,
                     len(m.time), m.n_tapers])
        geom[f"time_{ci}"] = np.asarray(m.time)
        geom[f"freq_{ci}"] = np.asarray(m.frequencies)
    np.savez_compressed(os.path.join(HERE, "geometry.npz"), table=np.array(rows, dtype=float), **geom)

    # ---- 2. tapers --------------------------------------------------------
    tap = {}
    for n, nw, k in [(64, 2.0, 3), (120, 5.0, 9), (1000, 3.0, 5), (1000, 4.0, 7), (257, 2.5, 4),
                     (50, 1.0, 1), (80, 2.0, 3)]:
        t_, ev = T.dpss_windows(n, nw, k, is_low_bias=False)
        tap[f"tapers_{n}_{nw}_{k}"] = np.asarray(t_)
        tap[f"eig_{n}_{nw}_{k}"] = np.asarray(ev)
    # a case where the low-bias filter drops tapers
    t_, ev = T.dpss_windows(64, 1.5, 4, is_low_bias=True)
    tap["lowbias_64_1.5_4"] = np.asarray(t_)
    np.savez_compressed(os.path.join(HERE, "tapers.npz"), **tap)

    # ---- 3. Multitaper.fft -------------------------------------------------
    mt = {}
    mt_cases = {
        "whole": dict(n=256, t=2, s=3, fs=200.0, kw=dict(time_halfbandwidth_product=2)),
        "sliding": dict(n=400, t=3, s=4, fs=200.0,
                        kw=dict(time_halfbandwidth_product=2, time_window_duration=0.4,
                                time_window_step=0.2)),
        "linear": dict(n=300, t=2, s=2, fs=100.0,
                       kw=dict(time_halfbandwidth_product=3, time_window_duration=1.0,
                               detrend_type="linear")),
        "nodetrend": dict(n=300, t=2, s=2, fs=100.0,
                          kw=dict(time_halfbandwidth_product=3, time_window_duration=1.0,
                                  detrend_type=None)),
        "crop": dict(n=200, t=2, s=3, fs=100.0,
                     kw=dict(time_halfbandwidth_product=2, time_window_duration=1.0,
                             n_fft_samples=64)),
        "pad": dict(n=200, t=2, s=3, fs=100.0,
                    kw=dict(time_halfbandwidth_product=2, time_window_duration=1.0,
                            n_fft_samples=128)),
        "odd": dict(n=189, t=2, s=5, fs=63.0,
                    kw=dict(time_halfbandwidth_product=2, time_window_duration=1.0)),
        "prime": dict(n=202, t=1, s=2, fs=101.0,
                      kw=dict(time_halfbandwidth_product=2, time_window_duration=1.0,
                              n_fft_samples=101)),
    }
    for name, c in mt_cases.items():
        x = series(100 + len(mt), c["n"], c["t"], c["s"], c["fs"])
        m = T.Multitaper(x, sampling_frequency=c["fs"], **c["kw"])
        mt[f"{name}_x"] = x
        mt[f"{name}_fft"] = np.asarray(m.fft())
        mt[f"{name}_tapers"] = np.asarray(m.tapers)
        mt[f"{name}_meta"] = np.array([m.n_time_samples_per_window, m.n_time_samples_per_step,
                                       m.n_fft_samples, c["fs"]], dtype=float)
    np.savez_compressed(os.path.join(HERE, "multitaper_fft.npz"), **mt)

    # ---- 4. connectivity measures -----------------------------------------
    conn = {}
    x = series(7, 300, 4, 4, 100.0)
    m = T.Multitaper(x, sampling_frequency=100.0, time_halfbandwidth_product=3,
                     time_window_duration=1.0)
    coef = np.asarray(m.fft())
    conn["x"] = x
    conn["coef"] = coef
    conn["meta"] = np.array([100.0, 3.0, 1.0])
    methods = ["power", "coherency", "coherence_magnitude", "coherence_phase",
               "imaginary_coherence", "phase_locking_value", "phase_lag_index",
               "weighted_phase_lag_index", "debiased_squared_phase_lag_index",
               "debiased_squared_weighted_phase_lag_index", "pairwise_phase_consistency",
               "pairwise_spectral_granger_prediction"]
    for et in ["trials_tapers", "trials", "tapers", "time", "time_trials", "time_tapers",
               "time_trials_tapers"]:
        c = C.Connectivity(coef, expectation_type=et, frequencies=m.frequencies, time=m.time)
        conn[f"{et}__csm"] = np.asarray(c._expectation_cross_spectral_matrix())
        for meth in methods:
            if meth == "pairwise_spectral_granger_prediction" and et != "trials_tapers":
                continue
            conn[f"{et}__{meth}"] = np.asarray(getattr(c, meth)())
    np.savez_compressed(os.path.join(HERE, "connectivity.npz"), **conn)

    # ---- 5. Wilson ----------------------------------------------------------
    wil = {}
    c = C.Connectivity(coef, frequencies=m.frequencies, time=m.time)
    csm = np.asarray(c._expectation_cross_spectral_matrix())
    sub = csm[..., np.array([[0], [1]]), np.array([[0, 1]])]
    wil["csm2"] = sub
    wil["g2"] = np.asarray(M.minimum_phase_decomposition(sub))
    sub3 = csm[..., np.array([[0], [2], [3]]), np.array([[0, 2, 3]])]
    wil["csm3"] = sub3
    wil["g3"] = np.asarray(M.minimum_phase_decomposition(sub3))
    wil["csm4"] = csm
    wil["g4"] = np.asarray(M.minimum_phase_decomposition(csm))
    np.savez_compressed(os.path.join(HERE, "wilson.npz"), **wil)
    # ---- 6. MVAR family (full-matrix Wilson) -----------------------------------
    mv = {}
    c = C.Connectivity(coef, frequencies=m.frequencies, time=m.time)
    mv["transfer_function"] = np.asarray(c._transfer_function)
    mv["noise_covariance"] = np.asarray(c._noise_covariance)
    mv["mvar_fourier_coefficients"] = np.asarray(c._MVAR_Fourier_coefficients)
    for meth in ["directed_transfer_function", "directed_coherence", "partial_directed_coherence",
                 "generalized_partial_directed_coherence", "direct_directed_transfer_function"]:
        mv[meth] = np.asarray(getattr(c, meth)())
    np.savez_compressed(os.path.join(HERE, "mvar.npz"), **mv)

    # ---- 7. SVD-based measures ---------------------------------------------------
    sv = {}
    x6 = series(9, 300, 5, 6, 100.0)
    m6 = T.Multitaper(x6, sampling_frequency=100.0, time_halfbandwidth_product=3, time_window_duration=1.0)
    c6 = C.Connectivity.from_multitaper(m6)
    labels = np.array([2, 0, 0, 1, 2, 1])
    cc, lab = c6.canonical_coherence(labels)
    gc_, gv_ = c6.global_coherence(max_rank=5)   # max_rank >= n_signals - 1: dense SVD branch (:2258-2266)
    sv["x"] = x6
    sv["labels"] = labels
    sv["canonical_coherence"] = np.asarray(cc)
    sv["canonical_labels"] = np.asarray(lab)
    sv["global_coherence"] = np.asarray(gc_)[..., :1]
    sv["global_vectors"] = np.asarray(gv_)[..., :1]
    np.savez_compressed(os.path.join(HERE, "svd_measures.npz"), **sv)
    psi_section(T, C)
    delay_section(T, C)
    print("golden fixtures written to", HERE)
    for f in sorted(os.listdir(HERE)):
        if f.endswith(".npz"):
            print(f, os.path.getsize(os.path.join(HERE, f)))


if __name__ == "__main__":
    main()
