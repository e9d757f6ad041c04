// CTA-cooperative mixed-radix Stockham FFT for arbitrary lengths.
//
// Used by both hot kernels: the fused multitaper FFT (fp32, replaces the reference's
// scipy.fft.fft call at transforms.py:1405) and the Wilson factorisation (fp64,
// replaces the fft/ifft pair at minimum_phase_decomposition.py:129-142).
//
// The butterflies are plain host/device templates so that tests/cpu/fft_host_test.cpp
// can compile this header with g++ and check the index arithmetic against numpy
// without a GPU.  Only sc_cta_fft() (the part that touches threadIdx/__syncthreads)
// is CUDA-only.
#pragma once
#include <math.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define SC_HD __host__ __device__ __forceinline__
#else
#define SC_HD inline
#endif

#define SC_FFT_MAX_STAGES 32

struct ScFftPlan {
    int n;
    int nstages;
    int radix[SC_FFT_MAX_STAGES];
};

// Factorise n into the radices the device code has register butterflies for
// (10, 8, 5, 4, 3, 2, 7, 11, 13); any other prime factor p becomes one O(p^2)
// "generic" stage.  Returns 0 on success.
static inline int sc_fft_make_plan(int n, ScFftPlan* p) {
    p->n = n;
    p->nstages = 0;
    if (n < 1) return -1;
    int m = n;
    const int pref[] = {10, 8, 5, 4, 3, 2, 7, 11, 13};
    for (int i = 0; i < 9; ++i) {
        const int r = pref[i];
        while (m % r == 0) {
            if (p->nstages >= SC_FFT_MAX_STAGES) return -1;
            p->radix[p->nstages++] = r;
            m /= r;
        }
    }
    for (int q = 17; m > 1; q += 2) {
        if ((long long)q * q > m) q = m;  // remaining m is prime
        while (m % q == 0) {
            if (p->nstages >= SC_FFT_MAX_STAGES) return -1;
            p->radix[p->nstages++] = q;
            m /= q;
        }
    }
    return 0;
}

template <typename R>
struct alignas(2 * sizeof(R)) cx {
    R x, y;
};

template <typename R> SC_HD cx<R> cmake(R a, R b) { cx<R> r; r.x = a; r.y = b; return r; }
template <typename R> SC_HD cx<R> cadd(cx<R> a, cx<R> b) { return cmake<R>(a.x + b.x, a.y + b.y); }
template <typename R> SC_HD cx<R> csub(cx<R> a, cx<R> b) { return cmake<R>(a.x - b.x, a.y - b.y); }
template <typename R> SC_HD cx<R> cmul(cx<R> a, cx<R> b) {
    return cmake<R>(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
template <typename R> SC_HD cx<R> cconj(cx<R> a) { return cmake<R>(a.x, -a.y); }
template <typename R> SC_HD cx<R> cscale(cx<R> a, R s) { return cmake<R>(a.x * s, a.y * s); }
// multiply by -i (forward) or +i (inverse)
template <typename R> SC_HD cx<R> cmul_mi(cx<R> a, bool inv) {
    return inv ? cmake<R>(-a.y, a.x) : cmake<R>(a.y, -a.x);
}

// ---------------------------------------------------------------------------
// register butterflies: v <- DFT_r(v) (forward: exp(-2 pi i/r), inverse: conj)
// ---------------------------------------------------------------------------
template <typename R> SC_HD void sc_dft2(cx<R>& a, cx<R>& b) {
    cx<R> t = csub(a, b);
    a = cadd(a, b);
    b = t;
}

template <typename R> SC_HD void sc_dft3(cx<R>* v, bool inv) {
    const R s = (R)0.86602540378443864676372317075294;  // sin(pi/3)
    cx<R> t1 = cadd(v[1], v[2]);
    cx<R> t2 = cmake<R>(v[0].x - (R)0.5 * t1.x, v[0].y - (R)0.5 * t1.y);
    cx<R> d = csub(v[1], v[2]);
    cx<R> t3 = cmul_mi(cscale(d, s), inv);  // -i*s*(v1-v2) forward
    v[0] = cadd(v[0], t1);
    v[1] = cadd(t2, t3);
    v[2] = csub(t2, t3);
}

template <typename R> SC_HD void sc_dft4(cx<R>* v, bool inv) {
    cx<R> a = cadd(v[0], v[2]), b = csub(v[0], v[2]);
    cx<R> c = cadd(v[1], v[3]), d = cmul_mi(csub(v[1], v[3]), inv);
    v[0] = cadd(a, c);
    v[1] = cadd(b, d);
    v[2] = csub(a, c);
    v[3] = csub(b, d);
}

template <typename R> SC_HD void sc_dft5(cx<R>* v, bool inv) {
    // 32 real operations (12 additions + 20 fused multiply-adds): the sine products are folded into the output
    // additions -- n1 = s1 (b1 + r b2), n2 = s1 (r b1 - b2) with r = s2/s1, and v1/v4 = m1 +- (-+i) n1 become one
    // FMA per component instead of a multiply and an add.
    const R c1 = (R)0.30901699437494742410229341718282;   // cos(2pi/5)
    const R c2 = (R)-0.80901699437494742410229341718282;  // cos(4pi/5)
    const R s1 = (R)0.95105651629515357211643933337938;   // sin(2pi/5)
    const R rr = (R)0.61803398874989484820458683436564;   // sin(4pi/5) / sin(2pi/5)
    const R sg = inv ? -s1 : s1;                          // forward: multiply by -i, inverse: by +i
    cx<R> a1 = cadd(v[1], v[4]), b1 = csub(v[1], v[4]);
    cx<R> a2 = cadd(v[2], v[3]), b2 = csub(v[2], v[3]);
    cx<R> m1 = cmake<R>(v[0].x + c1 * a1.x + c2 * a2.x, v[0].y + c1 * a1.y + c2 * a2.y);
    cx<R> m2 = cmake<R>(v[0].x + c2 * a1.x + c1 * a2.x, v[0].y + c2 * a1.y + c1 * a2.y);
    cx<R> t1 = cmake<R>(b1.x + rr * b2.x, b1.y + rr * b2.y);
    cx<R> t2 = cmake<R>(rr * b1.x - b2.x, rr * b1.y - b2.y);
    v[0] = cadd(v[0], cadd(a1, a2));
    // -i s1 t = (s1 t.y, -s1 t.x) forward; +i s1 t = (-s1 t.y, s1 t.x) inverse
    v[1] = cmake<R>(m1.x + sg * t1.y, m1.y - sg * t1.x);
    v[4] = cmake<R>(m1.x - sg * t1.y, m1.y + sg * t1.x);
    v[2] = cmake<R>(m2.x + sg * t2.y, m2.y - sg * t2.x);
    v[3] = cmake<R>(m2.x - sg * t2.y, m2.y + sg * t2.x);
}

template <typename R> SC_HD void sc_dft8(cx<R>* v, bool inv) {
    const R h = (R)0.70710678118654752440084436210485;
    // decimation in frequency: 8 = 2 x 4
    cx<R> e[4], o[4];
#pragma unroll
    for (int a = 0; a < 4; ++a) {
        e[a] = cadd(v[a], v[a + 4]);
        o[a] = csub(v[a], v[a + 4]);
    }
    // o[a] *= W8^a
    {
        cx<R> t = o[1];  // W8^1 = h(1 - i) forward, h(1 + i) inverse
        o[1] = inv ? cmake<R>(h * (t.x - t.y), h * (t.x + t.y)) : cmake<R>(h * (t.x + t.y), h * (t.y - t.x));
        o[2] = cmul_mi(o[2], inv);
        t = o[3];  // W8^3 = h(-1 - i) forward, h(-1 + i) inverse
        o[3] = inv ? cmake<R>(h * (-t.x - t.y), h * (t.x - t.y)) : cmake<R>(h * (t.y - t.x), h * (-t.x - t.y));
    }
    sc_dft4(e, inv);
    sc_dft4(o, inv);
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        v[2 * q] = e[q];
        v[2 * q + 1] = o[q];
    }
}

template <typename R> SC_HD void sc_dft10(cx<R>* v, bool inv) {
    // prime-factor (Good-Thomas) form of 10 = 2 x 5: with n = (5 n1 + 2 n2) mod 10 and k = (5 k1 + 6 k2) mod 10 the
    // kernel W10^(n k) splits into W2^(n1 k1) W5^(n2 k2), so there are NO twiddle products between the radix-2 and
    // the radix-5 parts (the decimation-in-frequency form needs four complex multiplications by W10^a).
    cx<R> e[5], o[5];
    e[0] = cadd(v[0], v[5]); o[0] = csub(v[0], v[5]);
    e[1] = cadd(v[2], v[7]); o[1] = csub(v[2], v[7]);
    e[2] = cadd(v[4], v[9]); o[2] = csub(v[4], v[9]);
    e[3] = cadd(v[6], v[1]); o[3] = csub(v[6], v[1]);
    e[4] = cadd(v[8], v[3]); o[4] = csub(v[8], v[3]);
    sc_dft5(e, inv);
    sc_dft5(o, inv);
    v[0] = e[0]; v[6] = e[1]; v[2] = e[2]; v[8] = e[3]; v[4] = e[4];
    v[5] = o[0]; v[1] = o[1]; v[7] = o[2]; v[3] = o[3]; v[9] = o[4];
}

// O(r^2) register DFT for the odd primes 7, 11, 13; roots taken from the twiddle table.
template <typename R, int RADIX>
SC_HD void sc_dft_prime(cx<R>* v, bool inv, const cx<R>* tw, int n) {
    cx<R> w[RADIX];
    const int st = n / RADIX;
#pragma unroll
    for (int t = 0; t < RADIX; ++t) {
        w[t] = tw[t * st];
        if (inv) w[t].y = -w[t].y;
    }
    cx<R> out[RADIX];
#pragma unroll
    for (int q = 0; q < RADIX; ++q) {
        cx<R> acc = v[0];
#pragma unroll
        for (int t = 1; t < RADIX; ++t) acc = cadd(acc, cmul(v[t], w[(q * t) % RADIX]));
        out[q] = acc;
    }
#pragma unroll
    for (int q = 0; q < RADIX; ++q) v[q] = out[q];
}

template <typename R, int RADIX>
SC_HD void sc_dft(cx<R>* v, bool inv, const cx<R>* tw, int n) {
    if (RADIX == 2) sc_dft2(v[0], v[1]);
    else if (RADIX == 3) sc_dft3(v, inv);
    else if (RADIX == 4) sc_dft4(v, inv);
    else if (RADIX == 5) sc_dft5(v, inv);
    else if (RADIX == 8) sc_dft8(v, inv);
    else if (RADIX == 10) sc_dft10(v, inv);
    else sc_dft_prime<R, RADIX>(v, inv, tw, n);
}

// One Stockham butterfly j (0 <= j < n/RADIX) of the stage whose already-transformed
// sub-length is Ls: reads src[j + t*n/RADIX], writes dst[(j-k)*RADIX + k + q*Ls], k = j % Ls.
// tw[q] = exp(-2 pi i q/n), q in [0, n).
template <typename R, int RADIX>
SC_HD void sc_fft_item_k(const cx<R>* src, cx<R>* dst, int j, int k, int n, int m, int Ls, int tws, const cx<R>* tw,
                         bool inv) {
    cx<R> v[RADIX];
#pragma unroll
    for (int t = 0; t < RADIX; ++t) v[t] = src[j + t * m];
    if (k != 0) {
#pragma unroll
        for (int t = 1; t < RADIX; ++t) {
            cx<R> w = tw[t * k * tws];  // t*k*tws < RADIX*Ls*tws = n
            if (inv) w.y = -w.y;
            v[t] = cmul(v[t], w);
        }
    }
    sc_dft<R, RADIX>(v, inv, tw, n);
    const int ob = (j - k) * RADIX + k;
#pragma unroll
    for (int q = 0; q < RADIX; ++q) dst[ob + q * Ls] = v[q];
}

template <typename R, int RADIX>
SC_HD void sc_fft_item(const cx<R>* src, cx<R>* dst, int j, int n, int Ls, const cx<R>* tw, bool inv) {
    const int m = n / RADIX;
    sc_fft_item_k<R, RADIX>(src, dst, j, j % Ls, n, m, Ls, m / Ls, tw, inv);
}

// Runtime-radix fallback (large prime factors): item = (j, q) pair, 0 <= item < n.
template <typename R>
SC_HD void sc_fft_item_generic(const cx<R>* src, cx<R>* dst, int item, int radix, int n, int Ls,
                               const cx<R>* tw, bool inv) {
    const int m = n / radix;
    const int j = item % m;
    const int q = item / m;
    const int k = j % Ls;
    const int tws = m / Ls;
    const long long base = (long long)k * tws + (long long)q * m;  // < n
    cx<R> acc = src[j];
    for (int t = 1; t < radix; ++t) {
        cx<R> w = tw[(int)((base * t) % n)];
        if (inv) w.y = -w.y;
        acc = cadd(acc, cmul(src[j + t * m], w));
    }
    dst[(j - k) * radix + k + q * Ls] = acc;
}

// ---------------------------------------------------------------------------
// Compile-time plans: every stride, sub-length and twiddle offset is a constant, so loads and
// stores become [reg + immediate] and the modulo by Ls a multiply-shift.  Twiddles come from
// per-stage tables T_s[(t-1)*Ls + k] = W_N^(t*k*N/(Ls*r)) laid out with k fastest, which
// consecutive butterflies read without shared-memory bank conflicts (the flat table read
// tw[t*k*tws] is a strided gather).  Only the register radices 2,3,4,5,8,10 are allowed.
// ---------------------------------------------------------------------------
template <int N, int... RS> struct ScStaticPlan {
    static constexpr int n = N;
};

template <int LS, int R0, int... REST> struct ScTwCount {
    static constexpr int value = (LS > 1 ? (R0 - 1) * LS : 0) + ScTwCount<LS * R0, REST...>::value;
};
template <int LS, int R0> struct ScTwCount<LS, R0> {
    static constexpr int value = (LS > 1 ? (R0 - 1) * LS : 0);
};

// number of stage-twiddle entries of a plan
template <typename PLAN> struct ScStaticTw;
template <int N, int... RS> struct ScStaticTw<ScStaticPlan<N, RS...>> {
    static constexpr int count = ScTwCount<1, RS...>::value;
};

// Twiddle factors of one butterfly (stage with sub-length LS > 1, k = j % LS): v[t] *= W^(t*k).
// POW = false reads the RADIX-1 factors from the stage table; POW = true reads only W^k and forms
// the higher powers by multiplication (w[t] = w[t/2] * w[t - t/2], depth log2 t) -- RADIX-2 fewer
// shared-memory loads per butterfly for RADIX-2 complex multiplies and a few ulp of accuracy, for
// shared-memory-bound callers that do not need the last bits (the fp32 phase of the Wilson iteration).
template <typename R, int RADIX, int LS, bool POW>
SC_HD void sc_static_twiddle(cx<R>* v, const cx<R>* tws, int k, bool inv) {
    if (POW) {
        cx<R> w[RADIX];
        w[1] = tws[k];
        if (inv) w[1].y = -w[1].y;
#pragma unroll
        for (int t = 2; t < RADIX; ++t) w[t] = cmul(w[t / 2], w[t - t / 2]);
#pragma unroll
        for (int t = 1; t < RADIX; ++t) v[t] = cmul(v[t], w[t]);
    } else {
#pragma unroll
        for (int t = 1; t < RADIX; ++t) {
            cx<R> w = tws[(t - 1) * LS + k];
            if (inv) w.y = -w.y;
            v[t] = cmul(v[t], w);
        }
    }
}

template <typename R, int N, int LS, int RADIX, bool POW = false>
SC_HD void sc_static_item(const cx<R>* src, cx<R>* dst, int j, const cx<R>* tws, bool inv) {
    constexpr int M = N / RADIX;
    const int k = LS == 1 ? 0 : (LS == M ? j : j % LS);
    cx<R> v[RADIX];
#pragma unroll
    for (int t = 0; t < RADIX; ++t) v[t] = src[j + t * M];
    if (LS > 1) sc_static_twiddle<R, RADIX, LS, POW>(v, tws, k, inv);
    sc_dft<R, RADIX>(v, inv, (const cx<R>*)0, N);
    const int ob = (j - k) * RADIX + k;
#pragma unroll
    for (int q = 0; q < RADIX; ++q) dst[ob + q * LS] = v[q];
}

// forward DFT of RADIX points whose upper half (v[RADIX/2..]) is zero; only radix 10 has a shortcut
template <typename R, int RADIX>
SC_HD void sc_dft_lowhalf(cx<R>* v) {
    if (RADIX == 10) {
        // prime-factor form (see sc_dft10) with v[5..9] = 0: the radix-2 part degenerates to sign changes
        cx<R> e[5], o[5];
        e[0] = v[0]; o[0] = v[0];
        e[1] = v[2]; o[1] = v[2];
        e[2] = v[4]; o[2] = v[4];
        e[3] = v[1]; o[3] = cmake<R>(-v[1].x, -v[1].y);
        e[4] = v[3]; o[4] = cmake<R>(-v[3].x, -v[3].y);
        sc_dft5(e, false);
        sc_dft5(o, false);
        v[0] = e[0]; v[6] = e[1]; v[2] = e[2]; v[8] = e[3]; v[4] = e[4];
        v[5] = o[0]; v[1] = o[1]; v[7] = o[2]; v[3] = o[3]; v[9] = o[4];
    } else {
#pragma unroll
        for (int t = RADIX / 2; t < RADIX; ++t) v[t] = cmake<R>((R)0, (R)0);
        sc_dft<R, RADIX>(v, false, (const cx<R>*)0, 0);
    }
}

// Fused middle of inverse-FFT -> elementwise window -> forward-FFT for a plan whose first and last radix
// are equal: butterfly j of the LAST inverse stage produces the samples n = j + q*M (q < RADIX), exactly
// the inputs of butterfly j of the FIRST forward stage, so the window is applied in registers and one
// write + one read of the whole sequence (and a barrier) disappear.  win(bb, n, value) returns the
// windowed sample; samples n >= KCUT are known to be windowed to zero (KCUT = N keeps all).
template <typename R, int N, int RADIX, int KCUT, bool POW, typename WIN>
SC_HD void sc_static_mid_item(const cx<R>* src, cx<R>* dst, int bb, int j, const cx<R>* tws_last, WIN& win) {
    constexpr int M = N / RADIX;
    constexpr bool HALF = (KCUT % M == 0) && (KCUT / M == RADIX / 2) && (RADIX % 2 == 0);
    cx<R> v[RADIX];
#pragma unroll
    for (int t = 0; t < RADIX; ++t) v[t] = src[j + t * M];
    sc_static_twiddle<R, RADIX, M, POW>(v, tws_last, j, true);
    sc_dft<R, RADIX>(v, true, (const cx<R>*)0, N);
    if (HALF) {
#pragma unroll
        for (int q = 0; q < RADIX / 2; ++q) v[q] = win(bb, j + q * M, v[q]);
        sc_dft_lowhalf<R, RADIX>(v);
    } else {
#pragma unroll
        for (int q = 0; q < RADIX; ++q)
            v[q] = (j + q * M < KCUT) ? win(bb, j + q * M, v[q]) : cmake<R>((R)0, (R)0);
        sc_dft<R, RADIX>(v, false, (const cx<R>*)0, N);
    }
#pragma unroll
    for (int q = 0; q < RADIX; ++q) dst[j * RADIX + q] = v[q];
}

// fill the stage tables of one stage from the flat table tw[q] = exp(-2 pi i q/N)
template <typename R, int N, int LS, int RADIX>
SC_HD void sc_static_fill_stage(cx<R>* tws, const cx<R>* tw, int tid, int nthreads) {
    if (LS > 1) {
        constexpr int TWS = N / (LS * RADIX);
        for (int e = tid; e < (RADIX - 1) * LS; e += nthreads) {
            const int t = e / LS + 1, k = e % LS;
            tws[e] = tw[t * k * TWS];
        }
    }
}

template <typename R, int N, int LS, int R0, int... REST> struct ScStaticStages {
    static constexpr int TWN = (LS > 1 ? (R0 - 1) * LS : 0);
    // SYNC is called after every stage (device: __syncthreads, host tests: no-op)
    template <int NB, typename SYNC>
    static SC_HD cx<R>* run(cx<R>* src, cx<R>* dst, const cx<R>* tws, bool inv, int tid, int nthreads, SYNC sync) {
        constexpr int M = N / R0;
        for (int idx = tid; idx < NB * M; idx += nthreads) {
            const int bb = idx / M, j = idx % M;
            sc_static_item<R, N, LS, R0>(src + bb * N, dst + bb * N, j, tws, inv);
        }
        sync();
        return ScStaticStages<R, N, LS * R0, REST...>::template run<NB>(dst, src, tws + TWN, inv, tid, nthreads, sync);
    }
    static SC_HD void fill(cx<R>* tws, const cx<R>* tw, int tid, int nthreads) {
        sc_static_fill_stage<R, N, LS, R0>(tws, tw, tid, nthreads);
        ScStaticStages<R, N, LS * R0, REST...>::fill(tws + TWN, tw, tid, nthreads);
    }
};
template <typename R, int N, int LS, int R0> struct ScStaticStages<R, N, LS, R0> {
    static_assert(LS * R0 == N, "radices must multiply to N");
    template <int NB, typename SYNC>
    static SC_HD cx<R>* run(cx<R>* src, cx<R>* dst, const cx<R>* tws, bool inv, int tid, int nthreads, SYNC sync) {
        constexpr int M = N / R0;
        for (int idx = tid; idx < NB * M; idx += nthreads) {
            const int bb = idx / M, j = idx % M;
            sc_static_item<R, N, LS, R0>(src + bb * N, dst + bb * N, j, tws, inv);
        }
        sync();
        return dst;
    }
    static SC_HD void fill(cx<R>* tws, const cx<R>* tw, int tid, int nthreads) {
        sc_static_fill_stage<R, N, LS, R0>(tws, tw, tid, nthreads);
    }
};

template <typename R, typename PLAN> struct ScStaticFft;
template <typename R, int N, int... RS> struct ScStaticFft<R, ScStaticPlan<N, RS...>> {
    static constexpr int n = N;
    static constexpr int tw_count = ScTwCount<1, RS...>::value;
    // NB transforms of length N at a + b*N; returns the buffer holding the (unnormalised) result
    template <int NB, typename SYNC>
    static SC_HD cx<R>* run(cx<R>* a, cx<R>* b, const cx<R>* tws, bool inv, int tid, int nthreads, SYNC sync) {
        return ScStaticStages<R, N, 1, RS...>::template run<NB>(a, b, tws, inv, tid, nthreads, sync);
    }
    static SC_HD void fill(cx<R>* tws, const cx<R>* tw, int tid, int nthreads) {
        ScStaticStages<R, N, 1, RS...>::fill(tws, tw, tid, nthreads);
    }
};

// inverse FFT -> window (zero beyond KCUT) -> forward FFT of NB sequences at a + b*N, three-stage plans
// with equal outer radices; 5 shared-memory passes instead of 7 (see sc_static_mid_item).  Both transforms
// unnormalised; returns the buffer that holds the forward result (always b).
template <typename R, typename PLAN, int KCUT, bool POW = false> struct ScStaticConv {
    static constexpr bool supported = false;
    template <int NB, typename SYNC, typename WIN>  // never called: callers branch on `supported`
    static SC_HD cx<R>* run(cx<R>* a, cx<R>*, const cx<R>*, int, int, SYNC, WIN) { return a; }
};
template <typename R, int N, int R0, int R1, int KCUT, bool POW>
struct ScStaticConv<R, ScStaticPlan<N, R0, R1, R0>, KCUT, POW> {
    static constexpr bool supported = true;
    template <int LS, int RADIX, int NB>
    static SC_HD void stage(const cx<R>* src, cx<R>* dst, const cx<R>* tws, bool inv, int tid, int nthreads) {
        constexpr int M = N / RADIX;
        for (int idx = tid; idx < NB * M; idx += nthreads) {
            const int bb = idx / M, j = idx % M;
            sc_static_item<R, N, LS, RADIX, POW>(src + bb * N, dst + bb * N, j, tws, inv);
        }
    }
    template <int NB, typename SYNC, typename WIN>
    static SC_HD cx<R>* run(cx<R>* a, cx<R>* b, const cx<R>* tws, int tid, int nthreads, SYNC sync, WIN win) {
        const cx<R>* tw1 = tws;                  // stage 1 table (LS = R0)
        const cx<R>* tw2 = tws + (R1 - 1) * R0;  // stage 2 table (LS = R0*R1)
        stage<1, R0, NB>(a, b, tws, true, tid, nthreads);
        sync();
        stage<R0, R1, NB>(b, a, tw1, true, tid, nthreads);
        sync();
        constexpr int M = N / R0;
        for (int idx = tid; idx < NB * M; idx += nthreads) {
            const int bb = idx / M, j = idx % M;
            sc_static_mid_item<R, N, R0, KCUT, POW>(a + bb * N, b + bb * N, bb, j, tw2, win);
        }
        sync();
        stage<R0, R1, NB>(b, a, tw1, false, tid, nthreads);
        sync();
        stage<R0 * R1, R0, NB>(a, b, tw2, false, tid, nthreads);
        sync();
        return b;
    }
};

// ---- register-resident twiddles ---------------------------------------------------------------------------------
// When a CTA transforms NB sequences with NB * N / R0 <= nthreads, thread `tid` executes butterfly j = tid % (N/R0)
// of sequence tid / (N/R0) in EVERY stage of EVERY transform: the twiddle factors it needs -- W^(t k) with
// k = j % R0 in the middle stages and k = j in the outer stages (R0 == R1 plans) -- never change, so they are loaded
// once into registers (2 * (R0 - 1) complex values) instead of 9 shared-memory loads per butterfly and stage (31 %
// of the shared-memory wavefronts of the Wilson iteration, whose FFT passes are shared-memory bound).
template <typename R, int RADIX> struct ScTwRegs {
    cx<R> mid[RADIX - 1];    // stage with sub-length LS = R0:   W_N^(t * (j % R0) * N / (R0 * R1)), t = 1 .. R1-1
    cx<R> outer[RADIX - 1];  // stage with sub-length LS = N/R0: W_N^(t * j),                        t = 1 .. R0-1
};

template <typename R, int RADIX>
SC_HD void sc_apply_twiddle_regs(cx<R>* v, const cx<R>* w, bool inv) {
#pragma unroll
    for (int t = 1; t < RADIX; ++t) v[t] = cmul(v[t], inv ? cconj(w[t - 1]) : w[t - 1]);
}

template <typename R, int N, int LS, int RADIX>
SC_HD void sc_static_item_rt(const cx<R>* src, cx<R>* dst, int j, const cx<R>* w, bool inv) {
    constexpr int M = N / RADIX;
    const int k = LS == 1 ? 0 : (LS == M ? j : j % LS);
    cx<R> v[RADIX];
#pragma unroll
    for (int t = 0; t < RADIX; ++t) v[t] = src[j + t * M];
    if (LS > 1) sc_apply_twiddle_regs<R, RADIX>(v, w, inv);
    sc_dft<R, RADIX>(v, inv, (const cx<R>*)0, N);
    const int ob = (j - k) * RADIX + k;
#pragma unroll
    for (int q = 0; q < RADIX; ++q) dst[ob + q * LS] = v[q];
}

// ScStaticConv for R0 == R1 three-stage plans with register twiddles; requires NB * N / R0 <= nthreads.
template <typename R, int N, int R0, int KCUT> struct ScStaticConvRt {
    static constexpr int M = N / R0;
    // twiddles of thread `tid` from the flat table tw[q] = exp(-2 pi i q / N)
    static SC_HD void load(ScTwRegs<R, R0>& w, const cx<R>* tw, int tid) {
        const int j = tid % M;
#pragma unroll
        for (int t = 1; t < R0; ++t) {
            w.mid[t - 1] = tw[(t * (j % R0) * (N / (R0 * R0))) % N];
            w.outer[t - 1] = tw[(t * j) % N];
        }
    }
    template <int NB, typename SYNC, typename WIN>
    static SC_HD cx<R>* run(cx<R>* a, cx<R>* b, const ScTwRegs<R, R0>& w, int tid, SYNC sync, WIN win) {
        const bool act = tid < NB * M;
        const int bb = tid / M, j = tid - bb * M;
        cx<R>* sa = a + bb * N;
        cx<R>* sb = b + bb * N;
        if (act) sc_static_item_rt<R, N, 1, R0>(sa, sb, j, w.mid, true);
        sync();
        if (act) sc_static_item_rt<R, N, R0, R0>(sb, sa, j, w.mid, true);
        sync();
        if (act) {  // last inverse stage -> window -> first forward stage, in registers (see sc_static_mid_item)
            constexpr bool HALF = (KCUT % M == 0) && (KCUT / M == R0 / 2) && (R0 % 2 == 0);
            cx<R> v[R0];
#pragma unroll
            for (int t = 0; t < R0; ++t) v[t] = sa[j + t * M];
            sc_apply_twiddle_regs<R, R0>(v, w.outer, true);
            sc_dft<R, R0>(v, true, (const cx<R>*)0, N);
            if (HALF) {
#pragma unroll
                for (int q = 0; q < R0 / 2; ++q) v[q] = win(bb, j + q * M, v[q]);
                sc_dft_lowhalf<R, R0>(v);
            } else {
#pragma unroll
                for (int q = 0; q < R0; ++q) v[q] = (j + q * M < KCUT) ? win(bb, j + q * M, v[q]) : cmake<R>((R)0, (R)0);
                sc_dft<R, R0>(v, false, (const cx<R>*)0, N);
            }
#pragma unroll
            for (int q = 0; q < R0; ++q) sb[j * R0 + q] = v[q];
        }
        sync();
        if (act) sc_static_item_rt<R, N, R0, R0>(sb, sa, j, w.mid, false);
        sync();
        if (act) sc_static_item_rt<R, N, R0 * R0, R0>(sa, sb, j, w.outer, false);
        sync();
        return b;
    }
};

typedef ScStaticPlan<1000, 10, 10, 10> ScPlan1000;
typedef ScStaticPlan<120, 10, 4, 3> ScPlan120;

#if defined(__CUDACC__)
template <typename R, int RADIX>
__device__ __forceinline__ void sc_cta_fft_pass(const cx<R>* src, cx<R>* dst, int nbatch, int bstride, int n,
                                                int Ls, const cx<R>* tw, bool inv) {
    // Integer divisions are hoisted out of the per-butterfly path: the first stage has k = 0, the
    // last stage has k = j, and the batch index advances by carry instead of idx / m.
    const int m = n / RADIX;
    const int tws = m / Ls;
    const int total = nbatch * m;
    const int step = blockDim.x;
    const int step_b = step / m, step_j = step - step_b * m;
    int bb = threadIdx.x / m;
    int j = threadIdx.x - bb * m;
    for (int idx = threadIdx.x; idx < total; idx += step) {
        const int k = Ls == 1 ? 0 : (Ls == m ? j : j % Ls);
        sc_fft_item_k<R, RADIX>(src + (size_t)bb * bstride, dst + (size_t)bb * bstride, j, k, n, m, Ls, tws, tw, inv);
        j += step_j;
        bb += step_b;
        if (j >= m) {
            j -= m;
            ++bb;
        }
    }
}

// nbatch FFTs of length plan.n at a + b*bstride, ping-pong with buffer b (same layout).
// All threads of the CTA must call it; returns the buffer that holds the result.
// Unnormalised in both directions.
// LEAN = true routes the register-hungry prime radices (7, 11, 13) through the runtime-radix
// path so that kernels compiled under a tight register cap (fp64 Wilson) do not spill.
template <typename R, bool LEAN = false>
__device__ cx<R>* sc_cta_fft(cx<R>* a, cx<R>* b, int nbatch, int bstride, const ScFftPlan& plan,
                             const cx<R>* tw, bool inv) {
    const int n = plan.n;
    cx<R>* src = a;
    cx<R>* dst = b;
    int Ls = 1;
    for (int s = 0; s < plan.nstages; ++s) {
        const int r = plan.radix[s];
        switch ((LEAN && (r == 7 || r == 11 || r == 13)) ? -1 : r) {
            case 2: sc_cta_fft_pass<R, 2>(src, dst, nbatch, bstride, n, Ls, tw, inv); break;
            case 3: sc_cta_fft_pass<R, 3>(src, dst, nbatch, bstride, n, Ls, tw, inv); break;
            case 4: sc_cta_fft_pass<R, 4>(src, dst, nbatch, bstride, n, Ls, tw, inv); break;
            case 5: sc_cta_fft_pass<R, 5>(src, dst, nbatch, bstride, n, Ls, tw, inv); break;
            case 7: sc_cta_fft_pass<R, 7>(src, dst, nbatch, bstride, n, Ls, tw, inv); break;
            case 8: sc_cta_fft_pass<R, 8>(src, dst, nbatch, bstride, n, Ls, tw, inv); break;
            case 10: sc_cta_fft_pass<R, 10>(src, dst, nbatch, bstride, n, Ls, tw, inv); break;
            case 11: sc_cta_fft_pass<R, 11>(src, dst, nbatch, bstride, n, Ls, tw, inv); break;
            case 13: sc_cta_fft_pass<R, 13>(src, dst, nbatch, bstride, n, Ls, tw, inv); break;
            default: {
                const int total = nbatch * n;
                for (int idx = threadIdx.x; idx < total; idx += blockDim.x) {
                    const int bb = idx / n;
                    sc_fft_item_generic<R>(src + (size_t)bb * bstride, dst + (size_t)bb * bstride, idx - bb * n, r,
                                           n, Ls, tw, inv);
                }
            }
        }
        __syncthreads();
        Ls *= r;
        cx<R>* t = src;
        src = dst;
        dst = t;
    }
    return src;
}
#endif  // __CUDACC__
