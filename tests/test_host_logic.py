"""CPU tests: host-side logic of the drop-in classes and the C-ABI surface (no compute)."""
import ctypes
import os
import re
import warnings

import numpy as np
import pytest
from conftest import ROOT, assert_parity, golden

from spectral_connectivity_b200 import _lib
from spectral_connectivity_b200._dpss import dpss_windows, make_tapers
from spectral_connectivity_b200.transforms import EXPECTATION_AXES, Multitaper, expectation_map


def test_geometry_bit_exact():
    g = golden("geometry.npz")
    for ci, row in enumerate(g["table"]):
        n_samples, fs, dur, step = int(row[0]), row[1], row[2], row[3]
        m = Multitaper(np.zeros((n_samples, 1, 1)), sampling_frequency=fs,
                       time_window_duration=None if dur < 0 else dur,
                       time_window_step=None if step < 0 else step, time_halfbandwidth_product=2)
        assert m.n_time_samples_per_window == int(row[4])
        assert m.n_time_samples_per_step == int(row[5])
        assert m.n_fft_samples == int(row[6])
        assert m.n_time_windows == int(row[7]) == len(m.time)
        assert m.n_tapers == int(row[8])
        assert np.array_equal(m.time, g[f"time_{ci}"])
        assert np.array_equal(m.frequencies, g[f"freq_{ci}"])


def test_tapers_match_reference():
    g = golden("tapers.npz")
    for key in g.files:
        if key.startswith("tapers_"):
            _, n, nw, k = key.split("_")
            t, e = dpss_windows(int(n), float(nw), int(k), is_low_bias=False)
            assert_parity(t, g[key], 1e-10, key)
            assert_parity(e, g["eig_" + key[len("tapers_"):]], 1e-10, key)
            assert np.allclose((t ** 2).sum(axis=1), 1.0)
    t, _ = dpss_windows(64, 1.5, 4, is_low_bias=True)
    assert_parity(t, g["lowbias_64_1.5_4"], 1e-10, "lowbias")
    assert make_tapers(64, 100.0, 2.0, 3).shape == (64, 3)


def test_multitaper_properties():
    # mirrors reference tests/test_transforms.py:62-229
    m = Multitaper(np.zeros((100, 1, 1)), sampling_frequency=1000, time_halfbandwidth_product=3)
    assert m.n_tapers == 5 and m.n_time_samples_per_window == 100 and m.n_fft_samples == 100
    assert m.time_window_duration == 0.1 and m.time_window_step == 0.1
    assert m.nyquist_frequency == 500 and m.n_signals == 1 and m.n_trials == 1
    assert m.frequency_resolution == pytest.approx(60.0)
    m = Multitaper(np.zeros((100, 1, 1)), n_fft_samples=5)
    assert m.n_fft_samples == 5 and len(m.frequencies) == 5
    m = Multitaper(np.zeros((100, 1, 1)), sampling_frequency=1000, time_window_duration=0.02,
                   time_window_step=0.01, start_time=2.0)
    assert m.n_time_samples_per_step == 10 and m.n_time_windows == 9
    assert np.allclose(m.time, 2.0 + np.arange(9) * 0.01)
    m = Multitaper(np.zeros((23, 1, 1)), n_time_samples_per_window=8, n_time_samples_per_step=3, n_tapers=2)
    assert m.n_time_windows == 6 and m.n_tapers == 2
    custom = np.ones((8, 2))
    m = Multitaper(np.zeros((23, 1, 1)), n_time_samples_per_window=8, tapers=custom)
    assert np.array_equal(m.tapers, custom)


def test_multitaper_validation():
    with pytest.raises(ValueError, match="3D"):
        Multitaper(np.zeros(10))
    with pytest.raises(ValueError, match="3D"):
        Multitaper(np.zeros((10, 2)))
    with pytest.raises(ValueError, match="sampling_frequency"):
        Multitaper(np.zeros((10, 1, 1)), sampling_frequency=0)
    with pytest.raises(ValueError, match="time_halfbandwidth_product"):
        Multitaper(np.zeros((10, 1, 1)), time_halfbandwidth_product=0.5)
    with pytest.raises(ValueError, match="time_window_duration"):
        Multitaper(np.zeros((10, 1, 1)), time_window_duration=-1)
    with pytest.raises(ValueError, match="time_window_step"):
        Multitaper(np.zeros((10, 1, 1)), time_window_step=0)
    with pytest.raises(ValueError, match="trend"):
        Multitaper(np.zeros((10, 1, 1)), detrend_type="quadratic")
    with pytest.warns(UserWarning, match="NaN"):
        Multitaper(np.full((10, 1, 1), np.nan))
    with pytest.warns(UserWarning, match="transposed"):
        Multitaper(np.zeros((3, 1, 8)))
    with pytest.warns(UserWarning, match="unusually large"):
        Multitaper(np.zeros((100, 1, 1)), time_halfbandwidth_product=11)
    with pytest.warns(UserWarning, match="gaps"):
        Multitaper(np.zeros((100, 1, 1)), sampling_frequency=100, time_window_duration=0.1, time_window_step=0.2)


@pytest.mark.parametrize("et", list(EXPECTATION_AXES))
def test_expectation_map_matches_numpy_mean(et):
    w, t, k = 3, 4, 5
    mapping, kept, nb, nr = expectation_map((w, t, k), et)
    vals = np.random.default_rng(0).standard_normal((w, t, k))
    acc = np.zeros(nb)
    cnt = np.zeros(nb)
    seen = set()
    for iw in range(w):
        for it in range(t):
            for ik in range(k):
                b = iw * mapping[0] + it * mapping[1] + ik * mapping[2]
                r = iw * mapping[3] + it * mapping[4] + ik * mapping[5]
                assert 0 <= b < nb and 0 <= r < nr and (b, r) not in seen
                seen.add((b, r))
                acc[b] += vals[iw, it, ik]
                cnt[b] += 1
    assert len(seen) == w * t * k and nb * nr == w * t * k
    ref = vals.mean(axis=EXPECTATION_AXES[et])
    assert np.allclose((acc / cnt).reshape(kept if kept else ()), ref)


def test_connectivity_validation_without_gpu():
    from spectral_connectivity_b200 import Connectivity
    with pytest.raises(ValueError, match="5-dimensional"):
        Connectivity(np.zeros((2, 2, 2), dtype=complex))
    with pytest.raises(ValueError, match="Did you mean 'trials_tapers'"):
        Connectivity(np.zeros((1, 1, 1, 1, 2), dtype=complex), expectation_type="tapers_trials")
    with pytest.raises(ValueError, match="Invalid expectation_type"):
        Connectivity(np.zeros((1, 1, 1, 1, 2), dtype=complex), expectation_type="bogus")


def test_abi_exports_every_declared_symbol():
    """The shared library loads and exports every function include/sc_b200.h declares."""
    header = open(os.path.join(ROOT, "include", "sc_b200.h")).read()
    body = re.sub(r"/\*.*?\*/", "", header, flags=re.S)
    declared = set(re.findall(r"\b(sc_[a-z0-9_]+)\s*\(", body))
    assert {"sc_mt_fft", "sc_csm", "sc_power", "sc_pairwise_epilogue", "sc_wilson2",
            "sc_granger_pairwise", "sc_repack_coefficients"} <= declared
    assert os.path.exists(_lib.LIB_PATH), "build the library first (__graft_entry__.build())"
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in sc_b200.h but not exported"
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    loaded = _lib.load()
    assert loaded.sc_version() >= 100
    assert loaded.sc_wilson_workspace_bytes(1000) > 0
    assert loaded.sc_mt_fft_workspace_bytes(1000, 1000) == 0
    assert loaded.sc_mt_fft_workspace_bytes(60000, 60000) > 0


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "spectral_connectivity_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f
                assert "/root/reference" not in src, f


def test_bench_keeps_native_banners_off_stdout(tmp_path):
    """bench.py prints ONE JSON line on stdout: banners that native libraries write to file descriptor 1 while the
    process group initialises are routed to stderr, and the descriptor is restored even on an exception."""
    import subprocess
    import sys
    script = tmp_path / "fd.py"
    script.write_text(
        "import ctypes, json, sys\n"
        f"sys.path.insert(0, {str(ROOT)!r})\n"
        "import bench\n"
        "libc = ctypes.CDLL(None)\n"
        "with bench.native_stdout_to_stderr():\n"
        "    libc.puts(b'NATIVE BANNER')\n"
        "    libc.fflush(None)\n"
        "try:\n"
        "    with bench.native_stdout_to_stderr():\n"
        "        raise RuntimeError('boom')\n"
        "except RuntimeError:\n"
        "    pass\n"
        "print(json.dumps({'ok': 1}))\n")
    res = subprocess.run([sys.executable, str(script)], capture_output=True, text=True, timeout=120)
    assert res.returncode == 0, res.stderr
    assert res.stdout.strip() == '{"ok": 1}'
    assert "NATIVE BANNER" in res.stderr


def test_statistics_helpers_match_reference():
    """_statistics.py (host bookkeeping behind delay / group_delay) against the live reference's statistics.py and
    connectivity.py:2102-2243 (tests/golden/make_golden.py: delay_section) -- float64 in, identical values out."""
    from spectral_connectivity_b200 import _statistics as st
    g = golden("delay.npz")
    z_default = st.fisher_z(g["stat_coh1"].copy(), 50)
    assert np.isnan(z_default).all() and np.isnan(g["stat_z_default"]).all()   # the reference's degenerate default
    z = st.fisher_z(g["stat_coh1"].copy(), 50, g["stat_coh2"].copy(), 70)
    assert_parity(z, g["stat_z_two"], 1e-12, "fisher z, two groups")
    p = st.upper_tail_p(z)
    assert_parity(p, g["stat_p"], 1e-12, "upper-tail p")
    assert np.array_equal(st.benjamini_hochberg(g["stat_p"], alpha=0.2), g["stat_bh"])
    assert np.array_equal(st.bonferroni(g["stat_p"], alpha=0.2), g["stat_bonf"])
    assert np.array_equal(np.apply_along_axis(st._independent_run, -2, g["stat_p"] < 0.3, 2, 3), g["stat_groups"])
    # statistics.py:43-47 (the reference's own documented example)
    assert st.benjamini_hochberg(np.array([0.001, 0.02, 0.04, 0.3, 0.8])).tolist() == [True, True, False, False, False]


def test_delay_group_delay_host_half_matches_reference():
    """Host half of delay / group_delay (band-pass, significance mask, masked phase arithmetic) fed with the ORACLE's
    coherency instead of the device's: equal to the live reference's outputs (tests/golden/delay.npz)."""
    import types
    from oracle import oracle as O
    from spectral_connectivity_b200.connectivity import Connectivity
    g = golden("delay.npz")
    x = g["x"]
    fs, nw, dur = g["meta"]
    n, step, nfft = O.window_geometry(x.shape[0], fs, duration=dur)
    coef = O.multitaper_fft(x, fs, O.dpss_tapers(n, nw, O.default_n_tapers(nw), fs), n, step, nfft)
    fnn = nfft // 2 + 1
    coh = O.coherency(coef)[..., :fnn, :, :]

    class Host:
        frequencies = O.frequencies(nfft, fs)[:fnn]
        n_observations = O.n_observations(coef.shape, "trials_tapers")
        coherency = staticmethod(lambda: coh.astype(np.complex64))
    h = Host()
    h._significant_band_phase = types.MethodType(Connectivity._significant_band_phase, h)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        d1 = Connectivity.delay(h, frequencies_of_interest=[5.0, 60.0])
        d2 = Connectivity.delay(h, frequencies_of_interest=[5.0, 60.0], frequency_resolution=4.0, n_range=2)
        gd = Connectivity.group_delay(h, frequencies_of_interest=[5.0, 60.0])
    for got, key in [(d1, "delay_band"), (d2, "delay_res"), (gd[0], "gd_delay"), (gd[1], "gd_slope"), (gd[2], "gd_r")]:
        assert got.shape == g[key].shape and np.array_equal(np.isnan(got), np.isnan(g[key])), key
        assert_parity(np.nan_to_num(got), np.nan_to_num(g[key]), 1e-5, key)


def test_granger_group_enumeration_model():
    """Integer model of the grouped problem order of granger_herm_kernel (groups = (row i, four aligned columns), members
    j > i): the closed-form prefix count, the binary-search decode and the member ranges cover every pair (i < j) of
    combinations(range(S), 2) exactly once, and the pair index formula matches the lexicographic order -- for every
    S % 4 == 0 the kernel accepts up to 1024 (csrc/granger_herm.cu: groups_before / decode in begin_group)."""
    from itertools import combinations

    def groups_before(i, s4):
        q4, r4 = i >> 2, i & 3
        return i * s4 - (2 * q4 * (q4 - 1) + q4 * (r4 + 1))

    for S in list(range(8, 132, 4)) + [256, 512, 1024]:
        s4 = S // 4
        per_window = groups_before(S - 1, s4)
        assert per_window == sum(s4 - (r + 1) // 4 for r in range(S - 1))
        order = {p: k for k, p in enumerate(combinations(range(S), 2))} if S <= 128 else None
        seen = 0
        starts = [groups_before(i, s4) for i in range(S)]
        assert all(b > a for a, b in zip(starts[:-1], starts[1:]))          # strictly increasing: the search is exact
        for g in ([*range(per_window)] if S <= 128 else [0, 1, per_window // 2, per_window - 2, per_window - 1]):
            lo, hi = 0, S - 2
            while lo < hi:
                mid = (lo + hi + 1) >> 1
                if groups_before(mid, s4) <= g:
                    lo = mid
                else:
                    hi = mid - 1
            i = lo
            j0 = 4 * ((i + 1) // 4 + (g - groups_before(i, s4)))
            jlo = j0 if j0 > i else i + 1
            assert 0 <= i < S - 1 and j0 % 4 == 0 and i < jlo <= j0 + 3 < S
            for j in range(jlo, j0 + 4):
                pk = i * (2 * S - i - 1) // 2 + (j - i - 1)
                if order is not None:
                    assert order[(i, j)] == pk
                seen += 1
        if S <= 128:
            assert seen == S * (S - 1) // 2


def test_unpack_upper_numpy_roundtrip():
    """unpack_upper (host side of compute(packed=...)): the packed index of (i, j >= i) is i S - i (i - 1) / 2 + j - i,
    the order sc_pack_upper writes (csrc/csm.cu) and numpy.triu_indices enumerates."""
    from spectral_connectivity_b200.connectivity import unpack_upper
    rng = np.random.default_rng(0)
    for n_sig in (1, 2, 5, 37):
        full = rng.normal(size=(2, 3, n_sig, n_sig)).astype(np.float32)
        full = full + np.swapaxes(full, -1, -2)
        packed = np.stack([[full[a, b][np.triu_indices(n_sig)] for b in range(3)] for a in range(2)])
        for i in range(n_sig):
            for j in range(i, n_sig):
                assert packed[0, 0, i * n_sig - i * (i - 1) // 2 + j - i] == full[0, 0, i, j]
        assert np.array_equal(unpack_upper(packed, n_sig), full)
