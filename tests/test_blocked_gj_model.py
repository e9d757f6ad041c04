"""Host models of the index logic inside csrc/wilson_blocked.cuh (the blocked in-place Gauss-Jordan inversion of
the large-matrix Wilson / MVAR path), checked with NumPy.  The CUDA kernels are tested on the GPU
(tests/test_gpu_large_mvar.py); these models pin the ALGORITHM the kernels implement: panel elimination producing
T_J, the row interchanges as a gather/scatter, the rank-nb update and the final column permutation traced per
column."""
import numpy as np
import pytest


def blocked_gj_inverse(a, nb):
    """zb_panel_kernel + zb_update_kernel + zb_unscramble_kernel in NumPy."""
    m = a.astype(complex).copy()
    s = m.shape[0]
    ipiv = np.arange(s)
    for j0 in range(0, s, nb):
        w = min(nb, s - j0)
        pan = m[:, j0:j0 + w].copy()
        for c in range(w):                      # unblocked in-place GJ on the panel, pivot rows >= j0 + c
            p = j0 + c
            piv = p + int(np.argmax(np.abs(pan[p:, c]) ** 2))
            ipiv[p] = piv
            if piv != p:
                pan[[p, piv]] = pan[[piv, p]]
            pinv = 1.0 / pan[p, c]
            f = pan[:, c].copy()
            rowp = pan[p] * pinv
            rowp[c] = pinv
            for r in range(s):
                if r == p:
                    pan[r] = rowp
                else:
                    t = pan[r] - f[r] * rowp
                    t[c] = -f[r] * pinv
                    pan[r] = t
        other = np.ones(s, bool)
        other[j0:j0 + w] = False
        rb, scatter = gather_scatter_plan(ipiv[j0:j0 + w], j0, w)
        old = m.copy()
        row_block = old[rb][:, other]           # RB = (P M)[J, other columns]
        for dst, src in scatter:                # displaced rows always come from ORIGINAL panel rows
            assert j0 <= src < j0 + w and not (j0 <= dst < j0 + w)
            m[dst, other] = old[src, other]
        mo = m[:, other]
        mo[j0:j0 + w] = 0
        m[:, other] = mo + pan @ row_block      # rank-nb update
        m[:, j0:j0 + w] = pan
    return m[:, trace_permutation(ipiv)]


def gather_scatter_plan(piv, j0, w):
    """Thread 0 of zb_panel_kernel: net effect of the w row interchanges (k <-> piv[k]) as `jsrc` (original row that
    ends at panel position k) and the list of outside rows that receive an original panel row."""
    jsrc = [j0 + k for k in range(w)]
    orow, osrc = [], []
    for k in range(w):
        pv = int(piv[k])
        if pv == j0 + k:
            continue
        if pv < j0 + w:
            jsrc[k], jsrc[pv - j0] = jsrc[pv - j0], jsrc[k]
        else:
            idx = orow.index(pv) if pv in orow else len(orow)
            cur = osrc[idx] if idx < len(orow) else pv
            if idx == len(orow):
                orow.append(pv)
                osrc.append(None)
            osrc[idx] = jsrc[k]
            jsrc[k] = cur
    return jsrc, list(zip(orow, osrc))


def trace_permutation(ipiv):
    """Last panel of zb_panel_kernel: undoing the row interchanges = a column permutation; each column is traced
    through the interchanges independently (s_k only moves k and ipiv[k] >= k)."""
    s = len(ipiv)
    perm = np.empty(s, int)
    for c in range(s):
        v = c
        for k in range(s):
            pv = ipiv[k]
            v = pv if v == k else (k if v == pv else v)
        perm[c] = v
    return perm


@pytest.mark.parametrize("s,nb", [(5, 2), (33, 16), (37, 16), (64, 32), (70, 24), (100, 32)])
def test_blocked_gauss_jordan_model(s, nb):
    rng = np.random.default_rng(s)
    a = rng.standard_normal((s, s)) + 1j * rng.standard_normal((s, s))
    a[:, 0] *= 1e-3                             # forces interchanges in the first panel
    x = blocked_gj_inverse(a, nb)
    assert np.abs(x - np.linalg.inv(a)).max() / np.abs(x).max() < 1e-11


def test_trace_permutation_equals_sequential_unscramble():
    rng = np.random.default_rng(1)
    for _ in range(50):
        s = int(rng.integers(2, 60))
        ipiv = np.array([rng.integers(k, s) for k in range(s)])
        cols = np.arange(s)
        for k in range(s - 1, -1, -1):
            cols[[k, ipiv[k]]] = cols[[ipiv[k], k]]
        assert np.array_equal(trace_permutation(ipiv), cols)


def test_gather_scatter_plan_equals_sequential_swaps():
    rng = np.random.default_rng(2)
    for _ in range(500):
        s = int(rng.integers(5, 40))
        w = int(rng.integers(1, min(8, s) + 1))
        j0 = int(rng.integers(0, s - w + 1))
        piv = []
        for k in range(w):
            lo = j0 + k
            reuse = [q for q in piv if q >= lo]
            piv.append(int(rng.choice(reuse)) if reuse and rng.random() < 0.4 else int(rng.integers(lo, s)))
        m = rng.standard_normal((s, 3))
        ref = m.copy()
        for k in range(w):
            if piv[k] != j0 + k:
                ref[[j0 + k, piv[k]]] = ref[[piv[k], j0 + k]]
        jsrc, scatter = gather_scatter_plan(piv, j0, w)
        assert np.array_equal(m[jsrc], ref[j0:j0 + w])
        out = m.copy()
        for dst, src in scatter:
            out[dst] = m[src]
        outside = np.ones(s, bool)
        outside[j0:j0 + w] = False
        assert np.array_equal(out[outside], ref[outside])
