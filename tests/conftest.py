"""pytest configuration: registers the ``gpu`` marker and shared helpers."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    # the C-ABI library is a build artefact (git-ignored): build it once if the tree is fresh
    lib = os.path.join(ROOT, "spectral_connectivity_b200", "libsc_b200.so")
    if not os.path.exists(lib):
        import shutil
        import subprocess
        if shutil.which("nvcc") or os.path.exists("/usr/local/cuda/bin/nvcc"):
            subprocess.run(["make", "-C", os.path.join(ROOT, "spectral_connectivity_b200", "csrc"), "-j8"],
                           check=False, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)


def golden(name):
    return np.load(os.path.join(GOLDEN, name))


def assert_parity(got, ref, tol=1e-5, what=""):
    """The parity contract of SURVEY.md section 8(d): identical shape and NaN mask,
    scale-normalised max error <= tol and allclose(rtol=tol, atol=tol*max|ref|)."""
    got = np.asarray(got)
    ref = np.asarray(ref)
    assert got.shape == ref.shape, f"{what}: shape {got.shape} != {ref.shape}"
    nan_g, nan_r = np.isnan(got), np.isnan(ref)
    assert np.array_equal(nan_g, nan_r), f"{what}: NaN masks differ ({nan_g.sum()} vs {nan_r.sum()})"
    if ref.size == 0 or nan_r.all():
        return
    scale = np.nanmax(np.abs(ref))
    scale = scale if scale > 0 else 1.0
    err = np.nanmax(np.abs(got - ref)) / scale
    log = os.environ.get("SC_PARITY_LOG")   # optional: record the achieved error of every parity check (margins)
    if log:
        with open(log, "a") as fh:
            fh.write(f"{what}\t{err:.3e}\t{tol:.1e}\n")
    assert err <= tol, f"{what}: scale-normalised error {err:.3e} > {tol}"
    ok = np.isclose(got, ref, rtol=tol, atol=tol * scale, equal_nan=True)
    assert ok.all(), f"{what}: {np.count_nonzero(~ok)} elements outside rtol/atol {tol}"


@pytest.fixture(autouse=True)
def _seed():
    np.random.seed(42)
