"""Host model of the defect-correction form of the Wilson iteration used by the pairwise Granger kernel
(csrc/granger_herm.cu: herm_iteration_defect): B = G^-1 S G^-H + I = 2I + E and the plus operator
(minimum_phase_decomposition.py:96-142) is linear with [2I]+ = I, hence G [B]+ = G + G [E]+, and only the small
defect E has to go through the FFTs -- in float32 on the device."""
import numpy as np
from scipy.fft import fft, ifft

from oracle import oracle as O


def _csm2(seed=3, n=64, n_trials=12, fs=200.0):
    x = O.synthetic_series(n, n_trials, 2, fs, seed=seed)
    coef = O.multitaper_fft(x, fs, O.dpss_tapers(n, 2, 3, fs), n, n, n)
    return O.expected_csm(coef)


def _plus_f32(e):
    """plus operator with the FFTs in float32 (complex64), as the device does for the defect"""
    c = ifft(e.astype(np.complex64), axis=-3)
    nf, s = e.shape[-3], e.shape[-1]
    c[..., 0, :, :] *= 0.5
    r, q = np.tril_indices(s, k=-1)
    c[..., 0, r, q] = 0
    c[..., (nf + 1) // 2:, :, :] = 0
    return fft(c, axis=-3).astype(np.complex128)


def test_plus_operator_is_linear_and_maps_2i_to_i():
    rng = np.random.default_rng(0)
    e = 1e-3 * (rng.standard_normal((2, 32, 2, 2)) + 1j * rng.standard_normal((2, 32, 2, 2)))
    eye = np.eye(2)
    assert np.allclose(O.plus_operator(2 * eye + e), eye + O.plus_operator(e), rtol=0, atol=1e-15)


def test_defect_iterations_reproduce_the_reference_iteration():
    """Run the reference iteration to the point where the device hands over to fp64 (update < 4e-4), then continue
    (a) with the plain fp64 iteration and (b) with the defect form whose projection runs in float32: same iteration
    count, factors equal to ~1e-10 absolute, two orders below the 1e-8 stopping tolerance."""
    csm = _csm2()
    eye = np.eye(2)
    g = np.zeros(csm.shape, dtype=complex)
    g[...] = O.wilson_initial(csm)
    for _ in range(60):                                    # "fp32 phase" stand-in: plain iterations
        b = np.linalg.solve(g, O._herm(np.linalg.solve(g, csm))) + eye
        new = g @ O.plus_operator(b)
        done = np.abs(new - g).max() < 4e-4 * np.abs(g).max()
        g = new
        if done:
            break
    ga, gb = g.copy(), g.copy()
    ita = itb = 0
    for _ in range(60):
        b = np.linalg.solve(ga, O._herm(np.linalg.solve(ga, csm))) + eye
        new = ga @ O.plus_operator(b)
        ita += 1
        err = np.abs(new - ga).max()
        ga = new
        if err < 1e-8:
            break
    for _ in range(60):
        e = np.linalg.solve(gb, O._herm(np.linalg.solve(gb, csm))) - eye      # defect, formed in float64
        dg = gb @ _plus_f32(e)                                               # projection in float32
        itb += 1
        gb = gb + dg
        if np.abs(dg).max() < 1e-8:
            break
    assert ita == itb
    assert np.abs(ga - gb).max() < 1e-9                       # absolute, 10x below the 1e-8 tolerance (measured 1.4e-10)
    assert np.abs(ga - gb).max() / np.abs(ga).max() < 1e-8    # relative (measured 1.5e-9)
