"""Host model of the repeated-squaring eigen-solver behind global_coherence for more than 64 signals
(csrc/wilson_general.cu: gc_init / gc_norm / gc_extract kernels).  It pins the numerical finding recorded in
DESIGN.md: an imaginary perturbation (1 + i b) of the dominant projector doubles with every squaring, so the
trace normalisation has to use the COMPLEX trace."""
import numpy as np


def top_eigen_by_squaring(c, n_squarings, complex_trace):
    p = c.astype(np.complex128)
    p = p / (np.trace(p) if complex_trace else np.trace(p).real)
    for _ in range(n_squarings):
        q = p @ p
        t = np.trace(q)
        if not complex_trace:
            t = t.real
        q = q / t
        if 1.0 - np.real(t) < 1e-13:
            p = q
            break
        p = q
    v = p[:, np.argmax(np.real(np.diag(p)))]
    return np.real(np.conj(v) @ c @ v) / np.real(np.conj(v) @ v), p


def _noisy_csm(seed=0, s=40, r=30):
    rng = np.random.default_rng(seed)
    x = rng.standard_normal((s, r)) + 1j * rng.standard_normal((s, r))
    c = (x @ np.conj(x.T) / r).astype(np.complex64).astype(np.complex128)
    # fp32-born matrices are not exactly Hermitian: the device's normalised trace came out as 1 + 2e-8j, i.e. the
    # input carries a component (1 + i b) x matrix with b ~ 2e-8
    return c * (1 + 2e-8j)


def test_complex_trace_normalisation_is_stable():
    c = _noisy_csm()
    lam_ref = np.linalg.eigvalsh(0.5 * (c + np.conj(c.T)))[-1]
    lam, p = top_eigen_by_squaring(c, 32, complex_trace=True)
    assert abs(lam - lam_ref) / lam_ref < 1e-9
    assert abs(np.sum(np.abs(p) ** 2) - 1.0) < 1e-9              # a rank-one projector: Frobenius norm 1


def test_real_trace_normalisation_runs_away():
    """The failure mode measured on the device before the fix: the imaginary component doubles per squaring."""
    c = _noisy_csm()
    growth = []
    for n in (12, 16, 20):
        _, p = top_eigen_by_squaring(c, n, complex_trace=False)
        growth.append(abs(np.sum(np.abs(p) ** 2) - 1.0))
    assert growth[1] > 50 * growth[0] and growth[2] > 50 * growth[1]   # x4 per squaring, x256 per four
