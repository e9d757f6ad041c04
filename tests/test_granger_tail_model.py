"""CPU model of the closed-form late tail of the pairwise Granger kernel (csrc/granger_herm.cu, tail recursion).

The reference halves every lag-0 coefficient of the causal factor and then zeroes its lower triangle
(minimum_phase_decomposition.py:132-138), so once the frequency-dependent modes have converged every further Wilson
iteration multiplies G by a constant upper-triangular 2x2 matrix P_j = I + upper_half(C_j^-1 M C_j^-T - I), C_{j+1} = C_j P_j.
The kernel runs that recursion exactly until its diagonal steps are tiny and then uses: the diagonal recursions are Newton
square roots (quadratic), the off-diagonal defect halves exactly and drags ca along by pa = -(3/2) pb_prev^2.  This test states both forms in float64 and checks that
they stop at the same iterate with the same accumulated factor."""
import numpy as np

TOL = 1e-8
JUMP_PA, JUMP_PD = 1e-6, 1e-9          # csrc/granger_herm.cu: SC_TAIL_JUMP_PA / SC_TAIL_JUMP_PD


def tail(m00, m01, m11, e, n, max_iter=60, jump=False):
    """(steps, conv, T) of the tail recursion started at C = P0 = I + e / 2 (upper half); n = column maxima of G."""
    ca, cb, cd = 1.0 + 0.5 * e[0], 0.5 * e[1], 1.0 + 0.5 * e[2]
    ta, tb, td = 1.0, 0.0, 1.0
    n00, n10, n01, n11 = n
    it, conv = 0, False
    while it < max_iter:
        ia, idd = 1.0 / ca, 1.0 / cd
        ib = -cb * ia * idd
        r00, r01, r11 = ia * m00 + ib * m01, ia * m01 + ib * m11, idd * m11
        e00, e01, e11 = r00 * ia + r01 * ib - 1.0, r01 * idd, r11 * idd - 1.0
        pa, pb, pd = 0.5 * e00, 0.5 * e01, 0.5 * e11
        d0 = max(n00, n10) * abs(pa)
        d1 = max(n00 * abs(pb) + n01 * abs(pd), n10 * abs(pb) + n11 * abs(pd))
        tb = ta * pb + tb * (1.0 + pd); ta *= 1.0 + pa; td *= 1.0 + pd
        cb = ca * pb + cb * (1.0 + pd); ca *= 1.0 + pa; cd *= 1.0 + pd
        it += 1
        if max(d0, d1) < TOL:
            conv = True
            break
        if jump and abs(pa) < JUMP_PA and abs(pd) < JUMP_PD:
            pbm, pam, nmax = pb, pa, max(n00, n10)
            while it < max_iter:
                pbm *= 0.5; pam = -6.0 * pbm * pbm      # = -(3/2) pb_prev^2: the drag of ca behind q(r)
                tb = ta * pbm + tb; ta = ta * pam + ta
                it += 1
                if nmax * max(abs(pam), abs(pbm)) < TOL:
                    conv = True
                    break
            break
    return it, conv, np.array([ta, tb, td])


def test_late_tail_closed_form_matches_the_recursion():
    rng = np.random.default_rng(5)
    n_equal, worst = 0, 0.0
    for _ in range(2000):
        e = rng.uniform(-1, 1, 3) * 10.0 ** rng.uniform(-4, -1.3)          # lag-0 residual at tail entry
        m00, m01, m11 = 1.0 + e[0], e[1], 1.0 + e[2]
        n = np.abs(rng.normal(size=4)) * 10.0 ** rng.uniform(-2, 0) + 1e-3
        full = tail(m00, m01, m11, e, n)
        fast = tail(m00, m01, m11, e, n, jump=True)
        assert full[1] and fast[1]
        n_equal += full[0] == fast[0]
        assert abs(full[0] - fast[0]) <= 1
        if full[0] == fast[0]:
            worst = max(worst, float(np.abs(full[2] - fast[2]).max()))
    assert n_equal >= 1998            # a stop within 1e-9 relative of the tolerance may move by one step
    assert worst < 1e-9               # accumulated factor T = product of the P_j


def test_off_diagonal_defect_halves_exactly():
    """With cd^2 = m11 the ratio r = cb / cd obeys r' = r + (m01 / m11 - r) / 2."""
    m00, m01, m11 = 1.3, 0.21, 0.8
    ca, cb, cd = 1.1, 0.05, np.sqrt(m11)
    for _ in range(5):
        ia, idd = 1.0 / ca, 1.0 / cd
        ib = -cb * ia * idd
        r01, r11 = ia * m01 + ib * m11, idd * m11
        pb, pd = 0.5 * r01 * idd, 0.5 * (r11 * idd - 1.0)
        r_old = cb / cd
        cb = ca * pb + cb * (1.0 + pd)
        cd *= 1.0 + pd
        assert abs(pd) < 1e-15
        assert abs(cb / cd - (r_old + 0.5 * (m01 / m11 - r_old))) < 1e-15
        r00 = ia * m00 + ib * m01
        ca *= 1.0 + 0.5 * (r00 * ia + r01 * ib - 1.0)
