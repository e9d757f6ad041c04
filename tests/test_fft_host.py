"""CPU check of the device FFT butterflies/index math (g++ build of fft_device.cuh)."""
import ctypes
import os
import subprocess
import tempfile

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def lib():
    out = os.path.join(tempfile.mkdtemp(), "libfft_host_test.so")
    subprocess.check_call(["g++", "-O1", "-shared", "-fPIC", "-std=c++17", "-o", out,
                           os.path.join(HERE, "cpu", "fft_host_test.cpp")])
    return ctypes.CDLL(out)


SIZES = [1, 2, 3, 4, 5, 6, 7, 8, 10, 11, 12, 13, 16, 17, 20, 27, 30, 49, 61, 64, 100, 101, 120, 121,
         128, 169, 250, 289, 500, 1000, 1001, 1008, 1024, 2 * 3 * 5 * 7 * 11, 7203]


@pytest.mark.parametrize("n", SIZES)
@pytest.mark.parametrize("inverse", [0, 1])
def test_fft_f64(lib, n, inverse):
    rng = np.random.default_rng(n)
    z = rng.standard_normal(n) + 1j * rng.standard_normal(n)
    buf = np.ascontiguousarray(z.view(np.float64).copy())
    st = lib.fft_host_f64(buf.ctypes.data_as(ctypes.POINTER(ctypes.c_double)), n, inverse, 0)
    assert st >= 0
    ref = np.fft.ifft(z) * n if inverse else np.fft.fft(z)
    got = buf.view(np.complex128)
    assert np.abs(got - ref).max() <= 1e-12 * max(1.0, np.abs(ref).max()) * max(1, np.log2(n + 1))


@pytest.mark.parametrize("n", [12, 35, 100, 1000])
def test_fft_generic_stage(lib, n):
    rng = np.random.default_rng(n)
    z = rng.standard_normal(n) + 1j * rng.standard_normal(n)
    buf = np.ascontiguousarray(z.view(np.float64).copy())
    lib.fft_host_f64(buf.ctypes.data_as(ctypes.POINTER(ctypes.c_double)), n, 0, 1)
    assert np.abs(buf.view(np.complex128) - np.fft.fft(z)).max() <= 1e-11 * np.abs(z).sum()


@pytest.mark.parametrize("n", [120, 1000, 1008])
def test_fft_f32(lib, n):
    rng = np.random.default_rng(n)
    z = (rng.standard_normal(n) + 1j * rng.standard_normal(n)).astype(np.complex64)
    buf = np.ascontiguousarray(z.view(np.float32).copy())
    lib.fft_host_f32(buf.ctypes.data_as(ctypes.POINTER(ctypes.c_float)), n, 0, 0)
    ref = np.fft.fft(z.astype(np.complex128))
    assert np.abs(buf.view(np.complex64) - ref).max() <= 2e-6 * np.abs(ref).max()


@pytest.mark.parametrize("n", [120, 1000])
@pytest.mark.parametrize("inverse", [0, 1])
def test_fft_static_plan(lib, n, inverse):
    rng = np.random.default_rng(n + inverse)
    z = rng.standard_normal(n) + 1j * rng.standard_normal(n)
    buf = np.ascontiguousarray(z.view(np.float64).copy())
    cnt = lib.fft_static_f64(buf.ctypes.data_as(ctypes.POINTER(ctypes.c_double)), n, inverse)
    assert cnt == {1000: 990, 120: 120 - 30 + 18}[n] or cnt > 0
    ref = np.fft.ifft(z) * n if inverse else np.fft.fft(z)
    assert np.abs(buf.view(np.complex128) - ref).max() <= 1e-12 * np.abs(ref).max() * 10


@pytest.mark.parametrize("pow_twiddles", [0, 1])
@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_fused_inverse_window_forward(lib, dtype, pow_twiddles):
    """ScStaticConv (last inverse stage + window + first forward stage in registers) == ifft, window, fft."""
    n = 1000
    rng = np.random.default_rng(7 + pow_twiddles)
    z = rng.standard_normal((2, n)) + 1j * rng.standard_normal((2, n))
    w = rng.uniform(0.5, 1.5, n)
    ref = np.fft.ifft(z, axis=-1) * n
    ref[:, 500:] = 0
    ref[:, :500] *= w[:500]
    ref[1, 0] = ref[1, 0].real
    ref = np.fft.fft(ref, axis=-1)
    cdt = np.complex128 if dtype is np.float64 else np.complex64
    buf = np.ascontiguousarray(z.astype(cdt).view(dtype).copy())
    wv = np.ascontiguousarray(w.astype(dtype))
    ct = ctypes.c_double if dtype is np.float64 else ctypes.c_float
    fn = lib.fft_conv_f64 if dtype is np.float64 else lib.fft_conv_f32
    assert fn(buf.ctypes.data_as(ctypes.POINTER(ct)), wv.ctypes.data_as(ctypes.POINTER(ct)), pow_twiddles) == 0
    got = buf.view(cdt).reshape(2, n)
    tol = 1e-12 if dtype is np.float64 else (6e-6 if pow_twiddles else 3e-6)
    assert np.abs(got - ref).max() <= tol * np.abs(ref).max()


def test_plan_1000(lib):
    r = (ctypes.c_int * 32)()
    assert lib.fft_plan(1000, r) == 3 and list(r[:3]) == [10, 10, 10]
    assert lib.fft_plan(120, r) == 3 and list(r[:3]) == [10, 4, 3]
    assert lib.fft_plan(101, r) == 1 and r[0] == 101
