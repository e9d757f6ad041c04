// Pairwise spectral Granger for REAL time series (conjugate-symmetric spectra) -- the
// from_multitaper hot path.  Same mathematics as wilson.cu (minimum_phase_decomposition.py:227-322,
// connectivity.py:2282-2340) but exploiting S(-f) = conj S(f):
//
//  * only the nfft/2+1 non-negative bins are ever evaluated; the 2x2 factor G(f) and the three
//    independent entries of S(f) live in REGISTERS (each thread owns FPT frequencies), so the
//    only shared-memory traffic is the FFT ping-pong buffer;
//  * the linear predictor B = G^-1 S G^-H + I is Hermitian with B(-f) = conj B(f), so its four lag sequences
//    are real: Z1 = B00 + i B11 and Z2 = B01 + i B10 are inverse-transformed by TWO complex FFTs into
//    c00 + i c11 and c01 + i c10, the plus operator is then elementwise on those packed sequences, and two
//    forward FFTs give P00 + i P11 and P01 + i P10, unpacked by Hermitian symmetry: 4 complex FFTs per
//    iteration instead of the reference's 8;
//  * for three-stage plans with equal outer radices (nfft = 1000 = 10*10*10) the last inverse stage, the
//    plus operator and the first forward stage run in registers (ScStaticConv), 5 shared-memory passes per
//    inverse+forward pair instead of 7;
//  * the Granger epilogue (transfer function, noise covariance, log ratio) is evaluated from the
//    registers and written straight into the (B, Fnn, S, S) output.
//
// Kernel variants (template flags of granger_herm_kernel; the launcher herm_launch picks one):
//   <.., LEAN=0, MIXED=0>           runtime precision mode, fp64 + fp32 buffers (88 KB at nfft = 1000): plain fp64 calls
//   <.., LEAN=0, MIXED=1>           mixed precision known at launch: fp32 buffers only (40 KB), explicit pair subsets
//   <.., LEAN=0, MIXED=1, GROUPED=1> the headline path: all pairs, S % 4 == 0 -- problems ordered by (window, row, four
//                                   aligned columns), one staging pass per group, 16-byte row stores (84 KB)
//   <.., LEAN=1>                    experiment (SC_GRANGER_LEAN=1): fp64 factor in shared memory, register twiddles
// Measured history and dead ends: profiles/r02_granger_experiments.txt.
#include <stdlib.h>

#include "wilson_common.cuh"

namespace {

using namespace scw;

constexpr int kMaxF32Iters = 12;     // fp32 phase never runs longer than this
#ifndef SC_GRANGER_GROUP
#define SC_GRANGER_GROUP 4
#endif
constexpr int kGroup = SC_GRANGER_GROUP;  // non-grouped order: consecutive pair indices one CTA handles back to back
#ifndef SC_TAIL_JUMP_PA
// closed-form late tail once the diagonal steps of the tail recursion are below these (float64 model of both forms:
// tests/test_granger_tail_model.py -- same stopping iterate in 4000 of 4000 random problems, accumulated factor equal
// to 5e-10 at 1e-6 / 1e-9 and to 1.5e-11 at 1e-7 / 1e-10; measured: 267.2 ms without the closed form, 263.3 ms at
// 1e-9 / 1e-12, 260.1 ms at 1e-7 / 1e-10)
#define SC_TAIL_JUMP_PA 1e-6
#define SC_TAIL_JUMP_PD 1e-9
#endif
constexpr double kTailRest = 32.0;   // closed-form tail once the non-constant part of the update is below
                                     // kTailRest * tol: those modes converge quadratically, so what is left of
                                     // them after the update is far below tol (measured deviation from the
                                     // reference's own final iterate: 2e-8 relative at 10 * tol)
// Hand over from the fp32 phase to the fp64 (defect-correction) phase once the non-constant part of the update
// (scaled units, |G'| ~ 1) is below kSwitch: the iteration converges quadratically in those modes, so one or two
// fp64-phase iterations then reach 1e-8.  Measured on config 4 (Granger stage, after the defect-correction form
// made the fp64-phase iterations cheaper): 1e-4 312 ms, 2e-4 303, 3e-4 300, 5e-4 303, 7e-4 304, 1.5e-3 306,
// 3e-3 305, 6e-3 312 -- a flat optimum around 3e-4 .. 7e-4.  Overridable at compile time for such sweeps.
#ifndef SC_GRANGER_KSWITCH
#define SC_GRANGER_KSWITCH 4e-4f
#endif
constexpr float kSwitch = SC_GRANGER_KSWITCH;

#ifndef SC_GRANGER_POW_TWIDDLES
#define SC_GRANGER_POW_TWIDDLES 0  // stage twiddles as powers of the first one (fewer shared loads, more
                                   // multiplies): 1 = fp32 phase, 2 = both.  Measured neutral on B200 -> off
#endif

struct CtaSync {
    __device__ __forceinline__ void operator()() const { __syncthreads(); }
};

// FFT policies: runtime plan (any length) or a compile-time plan (fft_device.cuh).
struct NoTwRegs {};  // placeholder of plans without register-resident twiddles

template <typename PLAN> struct PlanR0 { static constexpr int value = 0; };
template <int N, int R0, int... REST> struct PlanR0<ScStaticPlan<N, R0, REST...>> { static constexpr int value = R0; };

struct DynFft {
    static constexpr int kStaticN = 0;  // FFT length known at run time only
    static constexpr bool kRegTw = false;
    typedef NoTwRegs TwRegs;
    static __device__ __forceinline__ void load_regs(TwRegs&, const cx<float>*) {}
    template <typename WIN>
    static __device__ __forceinline__ cx<float>* conv2_rt(cx<float>* a, cx<float>*, const TwRegs&, WIN) { return a; }
    static constexpr bool kPrefetch = false;  // runtime-plan kernels are at the register cap already
    template <typename R> struct Fused { static constexpr bool value = false; };
    template <typename R, typename WIN>
    static __device__ __forceinline__ cx<R>* conv2(cx<R>* a, cx<R>*, const cx<R>*, WIN) { return a; }
    static __host__ __device__ int tw_entries(int n) { return n; }
    template <typename R>
    static __device__ __forceinline__ void fill(cx<R>* tws, const cx<R>* tw, int n) {
        for (int q = threadIdx.x; q < n; q += kThreads) tws[q] = tw[q];
    }
    template <typename R>
    static __device__ __forceinline__ cx<R>* run2(cx<R>* a, cx<R>* b, const ScFftPlan& plan, const cx<R>* tws,
                                                  bool inv) {
        return sc_cta_fft<R, true>(a, b, 2, plan.n, plan, tws, inv);
    }
};
template <typename PLAN, bool RT = ScStaticConv<float, PLAN, (PLAN::n + 1) / 2, false>::supported &&
                                   (PlanR0<PLAN>::value > 0) && (2 * (PLAN::n / (PlanR0<PLAN>::value > 0 ? PlanR0<PLAN>::value : 1)) <= kThreads)>
struct StatFftRegs {
    static constexpr bool kRegTw = false;
    typedef NoTwRegs TwRegs;
    static __device__ __forceinline__ void load_regs(TwRegs&, const cx<float>*) {}
    template <typename WIN>
    static __device__ __forceinline__ cx<float>* conv2_rt(cx<float>* a, cx<float>*, const TwRegs&, WIN) { return a; }
};
// three-stage plans with equal radices whose two packed sequences need at most one butterfly per thread and stage:
// the fp32 twiddles live in registers (fft_device.cuh: ScStaticConvRt)
template <typename PLAN> struct StatFftRegs<PLAN, true> {
    static constexpr bool kRegTw = true;
    static constexpr int R0 = PlanR0<PLAN>::value;
    typedef ScTwRegs<float, R0> TwRegs;
    typedef ScStaticConvRt<float, PLAN::n, R0, (PLAN::n + 1) / 2> Conv;
    static __device__ __forceinline__ void load_regs(TwRegs& w, const cx<float>* tw_flat) { Conv::load(w, tw_flat, threadIdx.x); }
    template <typename WIN>
    static __device__ __forceinline__ cx<float>* conv2_rt(cx<float>* a, cx<float>* b, const TwRegs& w, WIN win) {
        return Conv::template run<2>(a, b, w, threadIdx.x, CtaSync(), win);
    }
};

template <typename PLAN> struct StatFft : StatFftRegs<PLAN> {
    static constexpr int kStaticN = PLAN::n;  // compile-time length: bin guards, mirror indices and 1/N fold to constants
    static constexpr bool kPrefetch = true;  // fetch the next problem's spectrum under the epilogue
    static constexpr int kCut = (PLAN::n + 1) / 2;  // the plus operator keeps lags [0, kCut)
    template <typename R> struct Fused {
        static constexpr bool pow = SC_GRANGER_POW_TWIDDLES >= 2 || (SC_GRANGER_POW_TWIDDLES == 1 && sizeof(R) == 4);
        static constexpr bool value = ScStaticConv<R, PLAN, kCut, pow>::supported;
    };
    // inverse FFT -> window -> forward FFT of the two packed sequences (fft_device.cuh: ScStaticConv)
    template <typename R, typename WIN>
    static __device__ __forceinline__ cx<R>* conv2(cx<R>* a, cx<R>* b, const cx<R>* tws, WIN win) {
        return ScStaticConv<R, PLAN, kCut, Fused<R>::pow>::template run<2>(a, b, tws, threadIdx.x, kThreads, CtaSync(),
                                                                         win);
    }
    static __host__ __device__ int tw_entries(int) { return ScStaticTw<PLAN>::count; }
    template <typename R>
    static __device__ __forceinline__ void fill(cx<R>* tws, const cx<R>* tw, int) {
        ScStaticFft<R, PLAN>::fill(tws, tw, threadIdx.x, kThreads);
    }
    template <typename R>
    static __device__ __forceinline__ cx<R>* run2(cx<R>* a, cx<R>* b, const ScFftPlan&, const cx<R>* tws, bool inv) {
        return ScStaticFft<R, PLAN>::template run<2>(a, b, tws, inv, threadIdx.x, kThreads, CtaSync());
    }
};

template <typename R> struct RealOps;
template <> struct RealOps<double> {
    static __device__ __forceinline__ double rcp(double x) { return 1.0 / x; }
};
template <> struct RealOps<float> {
    // one MUFU.RCP (1 ulp) instead of the IEEE division sequence: the fp32 arithmetic only drives the iteration
    // (the fixed point and the stopping test are decided in fp64), and |det|^2 is O(1) on the row-scaled problem
    static __device__ __forceinline__ float rcp(float x) {
        float r;
        asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
        return r;
    }
};

// log(total) - log(total - part) of connectivity.py:1773-1779 (0 -> eps in the denominator, non-positive
// results -> NaN), rounded to the float32 output: the difference total - part is formed in fp64 exactly as the
// reference does, the logarithm of the RATIO is then taken in single precision -- through log1p when the
// ratio is close to one so that small Granger values keep their relative accuracy.
__device__ __forceinline__ float log_ratio(double total, double part) {
    double rest = total - part;
    if (rest == 0.0) rest = kEps64;
    // the differences are formed in fp64; the quotients are taken in fp32 (the output is float32 and a
    // correctly rounded fp32 division of two fp32-rounded operands is accurate to 2e-7 relative)
    // one MUFU reciprocal (1 ulp) instead of two IEEE divisions: 1e-7 relative on a float32 result
    float itotal;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(itotal) : "f"((float)total));
    const float ratio = (float)rest * itotal;
    float gc;
    if (ratio >= 0.5f) gc = -log1pf((float)(rest - total) * itotal);  // = -log1p(-(total - rest)/total)
    else gc = -logf(ratio);
    return gc > 0.f ? gc : __int_as_float(0x7fc00000);
}

// The plus operator (mpd.py:129-142) on the packed lag sequences z1 = c00 + i c11 (bb = 0) and
// z2 = c01 + i c10 (bb = 1): scale by 1/N, halve lag 0, zero the lag-0 lower triangle (c10[0]); lags >= kcut
// are zeroed by the caller.  Also records the lag-0 residual of G^-1 S G^-H - I.
// The 1/N of the inverse transform is NOT applied here (it would cost two multiplies per kept lag): the sequences
// stay scaled by N and the caller folds 1/N into the unpacking of the forward transforms.
template <typename R> struct PlusWindow {
    R inv_n;
    R* lag0;
    __device__ __forceinline__ cx<R> operator()(int bb, int k, cx<R> z) const {
        if (k == 0) {
            if (bb == 0) {
                lag0[0] = z.x * inv_n - (R)2;
                lag0[2] = z.y * inv_n - (R)2;
                return cmake<R>(z.x * (R)0.5, z.y * (R)0.5);
            }
            lag0[1] = z.x * inv_n;
            return cmake<R>(z.x * (R)0.5, (R)0);
        }
        return z;
    }
};

// One Wilson iteration on the half spectrum held in registers.  On return stat[0] = max |dG|^2 of
// this thread, stat[1] = max |dG - G_prev (P0 - I)|^2 where P0 is the lag-0 (constant) part of the
// causal factor (the update with the constant-matrix mode removed, see granger_herm_kernel),
// stat[2..5] = max over this thread's bins of |g00|^2, |g10|^2, |g01|^2, |g11|^2 of G_prev;
// lag0[0..2] = lag-0 residual e00, e01, e11 of G^-1 S G^-H - I (real for real time series).
template <typename R, int FPT, typename FFT, bool RT = false>
__device__ __forceinline__ void herm_iteration(cx<R> (&g00)[FPT], cx<R> (&g01)[FPT], cx<R> (&g10)[FPT],
                                               cx<R> (&g11)[FPT], const float (&s00)[FPT], const float (&s11)[FPT],
                                               const float2 (&s01)[FPT], R k00, R k11, R k01, cx<R>* ZA, cx<R>* ZB,
                                               const ScFftPlan& plan, const cx<R>* tw, int N, int fnn,
                                               R* lag0, R (&stat)[6], const typename FFT::TwRegs* twr = nullptr) {
    // ---- linear predictor (mpd.py:218-224), Hermitian: b00, b11 real, b10 = conj(b01) ----
#pragma unroll
    for (int q = 0; q < FPT; ++q) {
        const int f = threadIdx.x + q * kThreads;
        if (f < fnn) {
            const cx<R> det = csub(cmul(g00[q], g11[q]), cmul(g01[q], g10[q]));
            const R dn = RealOps<R>::rcp(det.x * det.x + det.y * det.y);
            const cx<R> idet = cmake<R>(det.x * dn, -det.y * dn);
            const cx<R> u0 = cmul(g11[q], idet), u1 = cmul(g01[q], idet);   // row 0 of G^-1 = (u0, -u1)
            const cx<R> v0 = cmul(g10[q], idet), v1 = cmul(g00[q], idet);   // row 1 of G^-1 = (-v0, v1)
            const R a = (R)s00[q] * k00, d = (R)s11[q] * k11;
            const cx<R> c = cmake<R>((R)s01[q].x * k01, (R)s01[q].y * k01);
            // M = Ginv S Ginv^H with rows r0 = (u0, -u1), r1 = (-v0, v1)
            // t = r S : t0 = r.x*a + r.y*conj(c), t1 = r.x*c + r.y*d
            const cx<R> cc = cconj(c);
            const cx<R> t00 = csub(cscale(u0, a), cmul(u1, cc));
            const cx<R> t01 = csub(cmul(u0, c), cscale(u1, d));
            const cx<R> t10 = csub(cmul(v1, cc), cscale(v0, a));
            const cx<R> t11 = csub(cscale(v1, d), cmul(v0, c));
            // M00 = t0 . conj(r0), M11 = t1 . conj(r1), M01 = t0 . conj(r1)
            const R b00 = (t00.x * u0.x + t00.y * u0.y) - (t01.x * u1.x + t01.y * u1.y) + (R)1;
            const R b11 = (t11.x * v1.x + t11.y * v1.y) - (t10.x * v0.x + t10.y * v0.y) + (R)1;
            // t0 . conj(r1) = -t00*conj(v0) + t01*conj(v1)
            const cx<R> b01 = cmake<R>(-(t00.x * v0.x + t00.y * v0.y) + (t01.x * v1.x + t01.y * v1.y),
                                       -(t00.y * v0.x - t00.x * v0.y) + (t01.y * v1.x - t01.x * v1.y));
            // Z1 = B00 + i B11, Z2 = B01 + i B10 with B10 = conj(B01) and B(-f) = conj(B(f)): the inverse
            // transforms are z1 = c00 + i c11 and z2 = c01 + i c10 (all four lag sequences are real)
            const int fm = f == 0 ? 0 : N - f;
            ZA[f] = cmake<R>(b00, b11);
            ZA[fm] = cmake<R>(b00, b11);
            // DC / Nyquist (f == fm): B01 is real there -- every operand of its imaginary part is an exact zero -- so
            // the two stores write the same value and no branch is needed
            ZA[N + f] = cmake<R>(b01.x + b01.y, b01.x + b01.y);
            ZA[N + fm] = cmake<R>(b01.x - b01.y, b01.x - b01.y);
        }
    }
    __syncthreads();
    // ---- plus operator (mpd.py:129-142) between the inverse and the forward transforms ----
    const PlusWindow<R> win = {(R)1 / (R)N, lag0};
    const cx<R>* Q;
    if constexpr (RT && sizeof(R) == 4) {
        Q = FFT::conv2_rt(ZA, ZB, *twr, win);
    } else if (FFT::template Fused<R>::value) {
        Q = FFT::template conv2<R>(ZA, ZB, tw, win);
    } else {
        cx<R>* c = FFT::template run2<R>(ZA, ZB, plan, tw, true);
        cx<R>* o = (c == ZA) ? ZB : ZA;
        const int kcut = (N + 1) / 2;
        for (int k = threadIdx.x; k < N; k += kThreads) {
            cx<R> y1 = cmake<R>((R)0, (R)0), y2 = y1;
            if (k < kcut) {
                y1 = win(0, k, c[k]);
                y2 = win(1, k, c[N + k]);
            }
            o[k] = y1;
            o[N + k] = y2;
        }
        __syncthreads();
        Q = FFT::template run2<R>(o, c, plan, tw, false);
    }
    // ---- G <- G P (mpd.py:305-307) ----
    // P = P0 + P' with the constant (lag-0) part P0 = I + (h00, h01; 0, h11): the non-constant part of the update,
    // rest = G P', is formed directly (its maximum steers the precision hand-over and the tail entry), and
    // G_new = G + dG with dG = G (P0 - I) + rest.  The transforms come back scaled by N (see PlusWindow): 0.5/N
    // is folded into the Hermitian unpacking.
    R err2 = (R)0, rest2 = (R)0, c00 = (R)0, c10 = (R)0, c01m = (R)0, c11m = (R)0;
    const R h00 = (R)0.5 * lag0[0], h01 = (R)0.5 * lag0[1], h11 = (R)0.5 * lag0[2];  // P0 - I
    const R h = (R)0.5 / (R)N;
    const R o00 = (R)1 + h00, o11 = (R)1 + h11;
#pragma unroll
    for (int q = 0; q < FPT; ++q) {
        const int f = threadIdx.x + q * kThreads;
        if (f < fnn) {
            const int fm = f == 0 ? 0 : N - f;
            const cx<R> a1 = Q[f], m1 = Q[fm], a2 = Q[N + f], m2 = Q[N + fm];
            // Y1 = P00 + i P11, Y2 = P01 + i P10 (spectra of real sequences): even / odd Hermitian parts, minus P0
            const cx<R> p00 = cmake<R>(h * (a1.x + m1.x) - o00, h * (a1.y - m1.y));
            const cx<R> p11 = cmake<R>(h * (a1.y + m1.y) - o11, h * (m1.x - a1.x));
            const cx<R> p01 = cmake<R>(h * (a2.x + m2.x) - h01, h * (a2.y - m2.y));
            const cx<R> p10 = cmake<R>(h * (a2.y + m2.y), h * (m2.x - a2.x));
            const cx<R> r00 = cadd(cmul(g00[q], p00), cmul(g01[q], p10));
            const cx<R> r01 = cadd(cmul(g00[q], p01), cmul(g01[q], p11));
            const cx<R> r10 = cadd(cmul(g10[q], p00), cmul(g11[q], p10));
            const cx<R> r11 = cadd(cmul(g10[q], p01), cmul(g11[q], p11));
            rest2 = fmax(rest2, fmax(fmax(r00.x * r00.x + r00.y * r00.y, r01.x * r01.x + r01.y * r01.y),
                                     fmax(r10.x * r10.x + r10.y * r10.y, r11.x * r11.x + r11.y * r11.y)));
            const cx<R> d00 = cmake<R>(r00.x + h00 * g00[q].x, r00.y + h00 * g00[q].y);
            const cx<R> d10 = cmake<R>(r10.x + h00 * g10[q].x, r10.y + h00 * g10[q].y);
            const cx<R> d01 = cmake<R>(r01.x + h01 * g00[q].x + h11 * g01[q].x, r01.y + h01 * g00[q].y + h11 * g01[q].y);
            const cx<R> d11 = cmake<R>(r11.x + h01 * g10[q].x + h11 * g11[q].x, r11.y + h01 * g10[q].y + h11 * g11[q].y);
            err2 = fmax(err2, fmax(fmax(d00.x * d00.x + d00.y * d00.y, d01.x * d01.x + d01.y * d01.y),
                                   fmax(d10.x * d10.x + d10.y * d10.y, d11.x * d11.x + d11.y * d11.y)));
            c00 = fmax(c00, g00[q].x * g00[q].x + g00[q].y * g00[q].y);
            c10 = fmax(c10, g10[q].x * g10[q].x + g10[q].y * g10[q].y);
            c01m = fmax(c01m, g01[q].x * g01[q].x + g01[q].y * g01[q].y);
            c11m = fmax(c11m, g11[q].x * g11[q].x + g11[q].y * g11[q].y);
            g00[q] = cadd(g00[q], d00); g01[q] = cadd(g01[q], d01);
            g10[q] = cadd(g10[q], d10); g11[q] = cadd(g11[q], d11);
        }
    }
    stat[0] = err2; stat[1] = rest2; stat[2] = c00; stat[3] = c10; stat[4] = c01m; stat[5] = c11m;
}

// Plus operator on the packed lag sequences of the DEFECT E = G^-1 S G^-H - I (B = 2I + E, [B]+ = I + [E]+):
// same window as PlusWindow, but lag 0 IS the residual (nothing to subtract).
struct PlusWindowDefect {
    float inv_n;
    float* lag0;
    __device__ __forceinline__ cx<float> operator()(int bb, int k, cx<float> z) const {
        if (k == 0) {
            if (bb == 0) {
                lag0[0] = z.x * inv_n;
                lag0[2] = z.y * inv_n;
                return cmake<float>(z.x * 0.5f, z.y * 0.5f);
            }
            lag0[1] = z.x * inv_n;
            return cmake<float>(z.x * 0.5f, 0.f);
        }
        return z;
    }
};

// Defect-correction form of one fp64-phase Wilson iteration (mixed-precision mode).  Near the fixed point
// B = G^-1 S G^-H + I = 2I + E with a small E, and the plus operator is linear with [2I]+ = I, so
// G P = G + G [E]+.  The defect is E = G^-1 (S - G G^H) G^-H: ONLY the residual R = S - G G^H suffers
// cancellation, so it alone is formed in fp64 (about 40 flops per bin); everything downstream of it -- the two
// congruence products with G^-1, the causal projection (four FFTs) and the product G [E]+ -- is LINEAR in R and
// runs in FP32 on the row-scaled problem (G' = D G, R' = D R D with D = diag(r0, r1), which leaves E unchanged
// and keeps every intermediate O(|E|) or O(1)): an fp32 error of ~3e-7 RELATIVE TO E (|E| <~ 1e-3 when fp64
// takes over, 1e-6 at the last iteration) moves G by ~1e-11 absolute, three orders below the 1e-8 stopping
// tolerance.  G itself is accumulated in fp64 (G += D^-1 dG').  Same outputs as herm_iteration<double>
// (stat[], lag0[] as float); the maxima are taken per row in scaled units and unscaled once per thread.
// Where the fp64 factor G of the defect iterations lives: in registers (arrays indexed by the thread's bin slot q), or
// in shared memory ([4][fnn] complex doubles, bin-major: consecutive threads touch consecutive 16-byte slots, each
// thread only ever touches its own bins).  The shared-memory variant lets the kernel run under an 85-register cap
// (3 CTAs per SM): the fp32 phase, where nearly all iterations happen, does not pay registers for fp64 state.
template <int FPT> struct GRegs {
    cd (&g00)[FPT]; cd (&g01)[FPT]; cd (&g10)[FPT]; cd (&g11)[FPT];
    __device__ __forceinline__ void load(int q, int, cd& a, cd& b, cd& c, cd& d) const { a = g00[q]; b = g01[q]; c = g10[q]; d = g11[q]; }
    __device__ __forceinline__ void store(int q, int, cd a, cd b, cd c, cd d) const { g00[q] = a; g01[q] = b; g10[q] = c; g11[q] = d; }
};
struct GSmem {
    cd* g;       // [4][stride]
    int stride;
    __device__ __forceinline__ void load(int, int f, cd& a, cd& b, cd& c, cd& d) const {
        a = g[f]; b = g[stride + f]; c = g[2 * stride + f]; d = g[3 * stride + f];
    }
    __device__ __forceinline__ void store(int, int f, cd a, cd b, cd c, cd d) const {
        g[f] = a; g[stride + f] = b; g[2 * stride + f] = c; g[3 * stride + f] = d;
    }
};

template <int FPT, typename FFT, typename GACC, bool RT = false>
__device__ __forceinline__ void herm_iteration_defect(const GACC gacc,
                                                      const float (&s00)[FPT], const float (&s11)[FPT],
                                                      const float2 (&s01)[FPT], float r0, float r1, cx<float>* ZA,
                                                      cx<float>* ZB, const ScFftPlan& plan, const cx<float>* tw, int N,
                                                      int fnn, float* lag0, double (&stat)[6],
                                                      const typename FFT::TwRegs* twr = nullptr) {
    const double dr0 = (double)r0, dr1 = (double)r1;  // powers of two: G' = D G is exact
#pragma unroll
    for (int q = 0; q < FPT; ++q) {
        const int f = threadIdx.x + q * kThreads;
        if (f < fnn) {
            cd G00, G01, G10, G11;
            gacc.load(q, f, G00, G01, G10, G11);
            // scaled factor G' = D G (exact) and the scaled residual R' = S' - G' G'^H in fp64 (s00, s11, s01 hold
            // S' = D S D, scaled in place by the caller)
            G00.x *= dr0; G00.y *= dr0; G01.x *= dr0; G01.y *= dr0;
            G10.x *= dr1; G10.y *= dr1; G11.x *= dr1; G11.y *= dr1;
            const double n00 = G00.x * G00.x + G00.y * G00.y + G01.x * G01.x + G01.y * G01.y;
            const double n11 = G10.x * G10.x + G10.y * G10.y + G11.x * G11.x + G11.y * G11.y;
            const double n01x = G00.x * G10.x + G00.y * G10.y + G01.x * G11.x + G01.y * G11.y;
            const double n01y = G00.y * G10.x - G00.x * G10.y + G01.y * G11.x - G01.x * G11.y;
            const float a = (float)((double)s00[q] - n00), d = (float)((double)s11[q] - n11);
            const cx<float> c = cmake<float>((float)((double)s01[q].x - n01x), (float)((double)s01[q].y - n01y));
            // its fp32 image and inverse
            const cx<float> f00 = cmake<float>((float)G00.x, (float)G00.y);
            const cx<float> f01 = cmake<float>((float)G01.x, (float)G01.y);
            const cx<float> f10 = cmake<float>((float)G10.x, (float)G10.y);
            const cx<float> f11 = cmake<float>((float)G11.x, (float)G11.y);
            const cx<float> det = csub(cmul(f00, f11), cmul(f01, f10));
            const float dn = RealOps<float>::rcp(det.x * det.x + det.y * det.y);
            const cx<float> idet = cmake<float>(det.x * dn, -det.y * dn);
            const cx<float> u0 = cmul(f11, idet), u1 = cmul(f01, idet);   // row 0 of G'^-1 = (u0, -u1)
            const cx<float> v0 = cmul(f10, idet), v1 = cmul(f00, idet);   // row 1 of G'^-1 = (-v0, v1)
            const cx<float> cc = cconj(c);
            const cx<float> t00 = csub(cscale(u0, a), cmul(u1, cc));
            const cx<float> t01 = csub(cmul(u0, c), cscale(u1, d));
            const cx<float> t10 = csub(cmul(v1, cc), cscale(v0, a));
            const cx<float> t11 = csub(cscale(v1, d), cmul(v0, c));
            const float e00 = (t00.x * u0.x + t00.y * u0.y) - (t01.x * u1.x + t01.y * u1.y);
            const float e11 = (t11.x * v1.x + t11.y * v1.y) - (t10.x * v0.x + t10.y * v0.y);
            const float e01x = -(t00.x * v0.x + t00.y * v0.y) + (t01.x * v1.x + t01.y * v1.y);
            const float e01y = -(t00.y * v0.x - t00.x * v0.y) + (t01.y * v1.x - t01.x * v1.y);
            const int fm = f == 0 ? 0 : N - f;
            ZA[f] = cmake<float>(e00, e11);
            ZA[fm] = cmake<float>(e00, e11);
            ZA[N + f] = cmake<float>(e01x + e01y, e01x + e01y);   // f == fm: e01y is an exact zero (see herm_iteration)
            ZA[N + fm] = cmake<float>(e01x - e01y, e01x - e01y);
        }
    }
    __syncthreads();
    const PlusWindowDefect win = {1.0f / (float)N, lag0};
    const cx<float>* Q;
    if constexpr (RT) {
        Q = FFT::conv2_rt(ZA, ZB, *twr, win);
    } else if (FFT::template Fused<float>::value) {
        Q = FFT::template conv2<float>(ZA, ZB, tw, win);
    } else {
        cx<float>* c = FFT::template run2<float>(ZA, ZB, plan, tw, true);
        cx<float>* o = (c == ZA) ? ZB : ZA;
        const int kcut = (N + 1) / 2;
        for (int k = threadIdx.x; k < N; k += kThreads) {
            cx<float> y1 = cmake<float>(0.f, 0.f), y2 = y1;
            if (k < kcut) {
                y1 = win(0, k, c[k]);
                y2 = win(1, k, c[N + k]);
            }
            o[k] = y1;
            o[N + k] = y2;
        }
        __syncthreads();
        Q = FFT::template run2<float>(o, c, plan, tw, false);
    }
    // row-wise maxima in scaled units: row 0 scales back by 1/r0, row 1 by 1/r1
    float err_r0 = 0.f, err_r1 = 0.f, rest_r0 = 0.f, rest_r1 = 0.f, c00 = 0.f, c10 = 0.f, c01m = 0.f, c11m = 0.f;
    const float h00 = 0.5f * lag0[0], h01 = 0.5f * lag0[1], h11 = 0.5f * lag0[2];  // P0 - I
    const double i0 = 1.0 / dr0, i1 = 1.0 / dr1;
    const float hn = 0.5f / (float)N;  // the transforms come back scaled by N (see PlusWindowDefect)
#pragma unroll
    for (int q = 0; q < FPT; ++q) {
        const int f = threadIdx.x + q * kThreads;
        if (f < fnn) {
            const int fm = f == 0 ? 0 : N - f;
            const cx<float> a1 = Q[f], m1 = Q[fm], a2 = Q[N + f], m2 = Q[N + fm];
            // [E]+ : Y1 = P00 + i P11, Y2 = P01 + i P10 (spectra of real sequences)
            const cx<float> p00 = cmake<float>(hn * (a1.x + m1.x), hn * (a1.y - m1.y));
            const cx<float> p11 = cmake<float>(hn * (a1.y + m1.y), hn * (m1.x - a1.x));
            const cx<float> p01 = cmake<float>(hn * (a2.x + m2.x), hn * (a2.y - m2.y));
            const cx<float> p10 = cmake<float>(hn * (a2.y + m2.y), hn * (m2.x - a2.x));
            cd G00, G01, G10, G11;
            gacc.load(q, f, G00, G01, G10, G11);
            const cx<float> f00 = cmake<float>((float)(G00.x * dr0), (float)(G00.y * dr0));
            const cx<float> f01 = cmake<float>((float)(G01.x * dr0), (float)(G01.y * dr0));
            const cx<float> f10 = cmake<float>((float)(G10.x * dr1), (float)(G10.y * dr1));
            const cx<float> f11 = cmake<float>((float)(G11.x * dr1), (float)(G11.y * dr1));
            // dG' = G' [E]+
            cx<float> d00 = cadd(cmul(f00, p00), cmul(f01, p10));
            cx<float> d01 = cadd(cmul(f00, p01), cmul(f01, p11));
            cx<float> d10 = cadd(cmul(f10, p00), cmul(f11, p10));
            cx<float> d11 = cadd(cmul(f10, p01), cmul(f11, p11));
            err_r0 = fmaxf(err_r0, fmaxf(d00.x * d00.x + d00.y * d00.y, d01.x * d01.x + d01.y * d01.y));
            err_r1 = fmaxf(err_r1, fmaxf(d10.x * d10.x + d10.y * d10.y, d11.x * d11.x + d11.y * d11.y));
            c00 = fmaxf(c00, f00.x * f00.x + f00.y * f00.y);
            c10 = fmaxf(c10, f10.x * f10.x + f10.y * f10.y);
            c01m = fmaxf(c01m, f01.x * f01.x + f01.y * f01.y);
            c11m = fmaxf(c11m, f11.x * f11.x + f11.y * f11.y);
            // G += D^-1 dG' (fp64 accumulation)
            G00.x += (double)d00.x * i0; G00.y += (double)d00.y * i0;
            G01.x += (double)d01.x * i0; G01.y += (double)d01.y * i0;
            G10.x += (double)d10.x * i1; G10.y += (double)d10.y * i1;
            G11.x += (double)d11.x * i1; G11.y += (double)d11.y * i1;
            gacc.store(q, f, G00, G01, G10, G11);
            // the update with its constant-matrix part G (P0 - I) removed
            d00.x -= h00 * f00.x; d00.y -= h00 * f00.y;
            d10.x -= h00 * f10.x; d10.y -= h00 * f10.y;
            d01.x -= h01 * f00.x + h11 * f01.x; d01.y -= h01 * f00.y + h11 * f01.y;
            d11.x -= h01 * f10.x + h11 * f11.x; d11.y -= h01 * f10.y + h11 * f11.y;
            rest_r0 = fmaxf(rest_r0, fmaxf(d00.x * d00.x + d00.y * d00.y, d01.x * d01.x + d01.y * d01.y));
            rest_r1 = fmaxf(rest_r1, fmaxf(d10.x * d10.x + d10.y * d10.y, d11.x * d11.x + d11.y * d11.y));
        }
    }
    const double w0 = i0 * i0, w1 = i1 * i1;
    stat[0] = fmax((double)err_r0 * w0, (double)err_r1 * w1);
    stat[1] = fmax((double)rest_r0 * w0, (double)rest_r1 * w1);
    stat[2] = (double)c00 * w0; stat[3] = (double)c10 * w1; stat[4] = (double)c01m * w0; stat[5] = (double)c11m * w1;
}

// LEAN = the mixed-precision mode (every FFT in fp32): fp32 ping-pong buffers only, the fp64 factor of the defect
// iterations in shared memory instead of registers, and -- for plans that support it (nfft = 1000) -- the FFT
// twiddles of each thread's fixed butterfly in the registers that frees (no twiddle tables in shared memory at all:
// the iteration's FFT passes are bound by shared-memory wavefronts, 31 % of which were twiddle loads).
// Measured dead end, kept out: an 85-register cap for 3 CTAs per SM (fp64 factor in shared memory, tables kept)
// leaves the iterations at the same speed -- they are throughput, not latency bound -- while the 11 KB of L1 that
// three 72 KB CTAs leave slow the strided cross-spectral gathers down by 2x (profiles/r02_granger_experiments.txt).
// MIXED = the launcher knows the call is in mixed-precision mode: like LEAN only the fp32 ping-pong buffers and tables
// are allocated (40 KB instead of 88 KB at nfft = 1000) and the plain fp64 iteration is compiled out, but the fp64
// factor of the defect iterations stays in registers.
//
// GROUPED (all pairs, S % 4 == 0, MIXED): a CTA works through GROUPS (window b, row i, four columns j0..j0+3 with
// j0 % 4 == 0, members j > i) instead of single pairs.  The strided gathers of a problem -- 501 bins x 5 values, every
// one in a different matrix, i.e. one 32-byte sector per 4 or 8 useful bytes, with a measured L1 hit rate of 2 % --
// are replaced by ONE staging pass per group: per bin the 32-byte sector S[i][j0..j0+3] (two 16-byte loads), the
// 16-byte power[j0..j0+3], the members' S[j][j] and S[i][i], power[i]: 9 loads instead of 20 and 8 sectors instead of
// 20 per bin and group.  Every thread stages exactly the bins it later consumes, so no barrier is needed.  The
// row outputs out[i][j0..j0+3] are collected in shared memory and written as one 16-byte store per bin.
template <int FPT, typename FFT, bool LEAN, bool MIXED = false, bool GROUPED = false>
__global__ void __launch_bounds__(kThreads, 512 / kThreads) granger_herm_kernel(const W2Params p) {
    static_assert(!GROUPED || (MIXED && !LEAN), "the grouped problem order exists for the mixed-precision kernel only");
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ double red4[2 * 4 * kWarps];
    int phase_s = 0;
    __shared__ double tail_sh[3];
    __shared__ int tail_it[2];
    __shared__ unsigned redf[2 * kWarps];
    __shared__ unsigned long long redd[2 * 6 * kWarps];
    int phase_f = 0, phase_d = 0;
    const int N = FFT::kStaticN > 0 ? FFT::kStaticN : p.nfft;
    const int fnn = N / 2 + 1;
    __shared__ double lag0_sh[3];
    __shared__ float lag0f_sh[3];
    // mixed-precision mode runs EVERY FFT in fp32 (fp32 phase + defect-correction iterations): the fp64 ping-pong
    // buffers and twiddle tables are then not allocated at all (40 KB instead of 88 KB of shared memory at
    // nfft = 1000, the rest stays L1 for the cross-spectral gathers)
    constexpr bool lean = LEAN || MIXED;  // fp32 buffers only
    cd* ZA = reinterpret_cast<cd*>(smem_raw);
    cd* ZB = ZA + 2 * (size_t)N;
    cd* tws = ZB + 2 * (size_t)N;
    cx<float>* twsf = reinterpret_cast<cx<float>*>(tws + FFT::tw_entries(N));
    cx<float>* ZAf = reinterpret_cast<cx<float>*>(ZA);  // the fp32 phase reuses the fp64 buffers
    cx<float>* ZBf = ZAf + 2 * (size_t)N;
    constexpr bool kRT = LEAN && FFT::kRegTw;  // fp32 twiddles in registers: no shared-memory tables
    if (lean) twsf = ZBf + 2 * (size_t)N;
    else FFT::template fill<double>(tws, p.tw, N);
    // LEAN: fp64 factor [4][fnn] behind the fp32 buffers (and tables, if any), 16-byte aligned
    const int gstride = fnn;
    cd* G64 = reinterpret_cast<cd*>(
        smem_raw + ((((size_t)4 * N + (kRT ? 0 : FFT::tw_entries(N))) * sizeof(cx<float>) + 15) & ~(size_t)15));
    const GSmem gsm = {G64, gstride};
    // GROUPED: staging area behind the fp32 buffers and tables (G64 is not used by this variant)
    float4* stS = reinterpret_cast<float4*>(G64);   // [fnn][2]  S[i][j0..j0+3] (4 complex)
    float4* stPj = stS + 2 * (size_t)fnn;           // [fnn]     power[j0..j0+3]
    float4* stDj = stPj + fnn;                      // [fnn]     Re S[j][j], j = j0..j0+3
    float4* stOut = stDj + fnn;                     // [fnn]     out[i][j0..j0+3]
    float* stPi = reinterpret_cast<float*>(stOut + fnn);  // [fnn] power[i]
    float* stDi = stPi + fnn;                       // [fnn]     Re S[i][i]
    typename FFT::TwRegs twr;
    if constexpr (kRT) FFT::load_regs(twr, p.tw32);
    else if (p.tw32) FFT::template fill<float>(twsf, p.tw32, N);
    __syncthreads();
    const long long npairs = p.n_pairs;
    const long long nprob = p.B * npairs;
    // executed-work counters of this CTA (uniform over the CTA; thread 0 publishes them at the end):
    // fp32-phase iterations, fp64-phase iterations, closed-form tail steps, problems
    unsigned long long cnt_f32 = 0, cnt_f64 = 0, cnt_tail = 0, cnt_prob = 0;
    const float fnan = __int_as_float(0x7fc00000);
    const double qnan = __longlong_as_double(0x7ff8000000000000LL);

    // the three independent entries of S(f), f = 0..nfft/2, of problem `prob` (pair pk of window b)
    float s00[FPT], s11[FPT];
    float2 s01[FPT];
    auto pair_of = [&](long long prob, long long& b, long long& pk, int& pi, int& pj) {
        b = prob / npairs;
        pk = prob % npairs;
        if (p.pairs) {
            pi = p.pairs[2 * pk];
            pj = p.pairs[2 * pk + 1];
        } else {
            decode_pair(pk, p.S, pi, pj);
        }
    };
    auto load_s = [&](long long b, int pi, int pj) {
#pragma unroll
        for (int q = 0; q < FPT; ++q) {
            const int f = threadIdx.x + q * kThreads;
            s00[q] = 0.f; s11[q] = 0.f; s01[q] = make_float2(0.f, 0.f);
            if (f < fnn) {
                const float2* m = reinterpret_cast<const float2*>(p.csm) + ((size_t)b * p.F + f) * p.S * p.S;
                s00[q] = __ldg(&m[(size_t)pi * p.S + pi]).x;
                s11[q] = __ldg(&m[(size_t)pj * p.S + pj]).x;
                s01[q] = __ldg(&m[(size_t)pi * p.S + pj]);
            }
        }
    };
    long long b = 0, pk = 0, nb_ = 0, npk = 0;
    int pi = 0, pj = 0, npi = 0, npj = 0;
    // Problem order: a CTA works through GROUPS of kGroup consecutive pair indices -- (i, j), (i, j+1), ... of one
    // window.  Their cross-spectral entries S_ij sit in the same 32-byte sectors (4 pairs each), S_ii is the same
    // element, and their power / output columns share sectors too, so after the first problem of a group the strided
    // gathers below (501 bins x 3 entries, every one in a different matrix) mostly hit the L1 the kernel leaves free.
    const long long first = (long long)blockIdx.x * kGroup;
    auto next_prob = [&](long long pr) {
        return ((pr + 1) % kGroup != 0) ? pr + 1 : (pr / kGroup + gridDim.x) * kGroup;
    };
    // ---- GROUPED problem order: groups (b, i, jb) with jb = j0 / 4 in [(i + 1) / 4, S / 4) ----
    const int S4 = (int)(p.S / 4);
    // groups of the rows before row i: sum_{r < i} (S4 - (r + 1) / 4) = i S4 - (2 q (q - 1) + q (r + 1)), i = 4 q + r
    auto groups_before = [&](int i) -> long long {
        const long long q4 = i >> 2, r4 = i & 3;
        return (long long)i * S4 - (2 * q4 * (q4 - 1) + q4 * (r4 + 1));
    };
    const long long groups_per_window = GROUPED ? groups_before((int)p.S - 1) : 0;
    const long long ngroups = groups_per_window * p.B;
    int j0 = 0, jlo = 0;
    auto begin_group = [&](long long g) {
        b = g / groups_per_window;
        const long long gw = g - b * groups_per_window;
        int lo = 0, hi = (int)p.S - 2;          // largest row i with groups_before(i) <= gw
        while (lo < hi) {
            const int mid = (lo + hi + 1) >> 1;
            if (groups_before(mid) <= gw) lo = mid;
            else hi = mid - 1;
        }
        pi = lo;
        j0 = 4 * ((pi + 1) / 4 + (int)(gw - groups_before(pi)));
        jlo = j0 > pi ? j0 : pi + 1;
        // stage: every thread loads the bins it consumes itself (no barrier)
#pragma unroll
        for (int q = 0; q < FPT; ++q) {
            const int f = threadIdx.x + q * kThreads;
            if (f < fnn) {
                const float2* m = reinterpret_cast<const float2*>(p.csm) + ((size_t)b * p.F + f) * p.S * p.S;
                const float4* row = reinterpret_cast<const float4*>(m + (size_t)pi * p.S + j0);
                stS[2 * f] = __ldg(row);
                stS[2 * f + 1] = __ldg(row + 1);
                const float* pw = p.power + ((size_t)b * p.F + f) * p.S;
                stPj[f] = __ldg(reinterpret_cast<const float4*>(pw + j0));
                stPi[f] = __ldg(pw + pi);
                stDi[f] = __ldg(&m[(size_t)pi * p.S + pi]).x;
                float4 dj = make_float4(0.f, 0.f, 0.f, 0.f);
                if (jlo <= j0) dj.x = __ldg(&m[(size_t)j0 * p.S + j0]).x;
                if (jlo <= j0 + 1) dj.y = __ldg(&m[(size_t)(j0 + 1) * p.S + j0 + 1]).x;
                if (jlo <= j0 + 2) dj.z = __ldg(&m[(size_t)(j0 + 2) * p.S + j0 + 2]).x;
                dj.w = __ldg(&m[(size_t)(j0 + 3) * p.S + j0 + 3]).x;
                stDj[f] = dj;
            }
        }
    };
    // write the collected row outputs out[i][jlo..j0+3] of the finished group
    auto flush_group = [&]() {
        float* out = reinterpret_cast<float*>(p.out);
#pragma unroll
        for (int q = 0; q < FPT; ++q) {
            const int f = threadIdx.x + q * kThreads;
            if (f < fnn) {
                float* dst = out + (((size_t)b * fnn + f) * p.S + pi) * p.S + j0;
                const float4 v = stOut[f];
                if (jlo == j0) {
                    *reinterpret_cast<float4*>(dst) = v;
                } else {
                    if (jlo <= j0 + 1) dst[1] = v.y;
                    if (jlo <= j0 + 2) dst[2] = v.z;
                    dst[3] = v.w;
                }
            }
        }
    };
    long long grp = blockIdx.x, prob = first;
    bool have;
    if constexpr (GROUPED) {
        have = grp < ngroups;
        if (have) {
            begin_group(grp);
            pj = jlo;
        }
    } else {
        have = first < nprob;
        if (have) {
            pair_of(first, b, pk, pi, pj);
            load_s(b, pi, pj);
        }
    }
    while (have) {
        if constexpr (GROUPED) {
            pk = (long long)pi * (2 * p.S - pi - 1) / 2 + (pj - pi - 1);  // index of (i, j) in combinations(range(S), 2)
            const int c = pj - j0;
#pragma unroll
            for (int q = 0; q < FPT; ++q) {
                const int f = threadIdx.x + q * kThreads;
                s00[q] = 0.f; s11[q] = 0.f; s01[q] = make_float2(0.f, 0.f);
                if (f < fnn) {
                    s00[q] = stDi[f];
                    s11[q] = reinterpret_cast<const float*>(stDj + f)[c];
                    s01[q] = reinterpret_cast<const float2*>(stS + 2 * f)[c];
                }
            }
        } else if (!FFT::kPrefetch && prob != first) {
            pair_of(prob, b, pk, pi, pj);
            load_s(b, pi, pj);
        }
        double a[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll
        for (int q = 0; q < FPT; ++q) {
            const int f = threadIdx.x + q * kThreads;
            if (f < fnn) {
                const double w = (f == 0 || 2 * f == N) ? 1.0 : 2.0;  // bins f and nfft-f
                a[0] += w * s00[q]; a[1] += w * s01[q].x; a[2] += w * s11[q];
            }
        }
        block_sum4(a, red4, phase_s);
        // ---- Cholesky of the real lag-0 covariance, G0 = L^T (mpd.py:75-77) ----
        const double inv_nd = 1.0 / (double)N;
        const double a00 = a[0] * inv_nd, a10 = a[1] * inv_nd, a11 = a[2] * inv_nd;
        // reciprocal square roots (one MUFU + Newton each) instead of sqrt, division, sqrt
        const double x00 = rsqrt(a00);
        const double l00 = a00 * x00, l10 = a10 * x00, d11 = a11 - l10 * l10, l11 = d11 * rsqrt(d11);
        int flag = 0, it_done = 0;
        if (!(a00 > 0.0) || !(d11 > 0.0) || !isfinite(l00) || !isfinite(l11)) flag = SC_FLAG_NOT_SPD;
        constexpr int GN = LEAN ? 1 : FPT;  // LEAN keeps the fp64 factor in shared memory (G64)
        cd g00[GN], g01[GN], g10[GN], g11[GN];
#pragma unroll
        for (int q = 0; q < GN; ++q) {
            g00[q] = cmake<double>(l00, 0.0); g01[q] = cmake<double>(l10, 0.0);
            g10[q] = cmake<double>(0.0, 0.0); g11[q] = cmake<double>(l11, 0.0);
        }
        const GRegs<GN> greg = {g00, g01, g10, g11};
        // this thread's bin slot q (frequency f) of the fp64 factor, wherever it lives
        auto gload = [&](int q, int f, cd& a_, cd& b_, cd& c_, cd& d_) {
            if constexpr (LEAN) gsm.load(q, f, a_, b_, c_, d_);
            else greg.load(q, f, a_, b_, c_, d_);
        };
        auto gstore = [&](int q, int f, cd a_, cd b_, cd c_, cd d_) {
            if constexpr (LEAN) gsm.store(q, f, a_, b_, c_, d_);
            else greg.store(q, f, a_, b_, c_, d_);
        };
        if (!flag) {
            bool converged = false;
            int it0 = 0;
            // row scales D = diag(a00, a11)^-1/2 of the fp32 arithmetic (fp32 phase and defect iterations)
            // POWERS OF TWO (so that S' and G' = D G are exact): r = 2^-floor(e/2) for a = m 2^e, m in [0.5, 1)
            int e0, e1;
            frexp(a00, &e0);
            frexp(a11, &e1);
            const float r0 = ldexpf(1.f, max(-60, min(60, -(e0 >> 1)))), r1 = ldexpf(1.f, max(-60, min(60, -(e1 >> 1))));
            if (MIXED || (p.tw32 && p.mixed)) {
                // S' = D S D in place (exact); the next problem's prefetch overwrites the registers anyway
#pragma unroll
                for (int q = 0; q < FPT; ++q) {
                    s00[q] *= r0 * r0; s11[q] *= r1 * r1; s01[q].x *= r0 * r1; s01[q].y *= r0 * r1;
                }
                // ---- fp32 phase: the same iteration on the row-scaled problem S' = D S D, G' = D G with
                // D = diag(a00, a11)^-1/2 (the iteration is equivariant under a left diagonal scaling, so
                // this only keeps every intermediate O(1) in single precision).  It stops as soon as the
                // update falls below kSwitch (relative), after which fp64 takes over: the fixed point and
                // the stopping test are decided entirely in fp64.
                cx<float> f00[FPT], f01[FPT], f10[FPT], f11[FPT];
#pragma unroll
                for (int q = 0; q < FPT; ++q) {
                    f00[q] = cmake<float>((float)l00 * r0, 0.f); f01[q] = cmake<float>((float)l10 * r0, 0.f);
                    f10[q] = cmake<float>(0.f, 0.f);             f11[q] = cmake<float>((float)l11 * r1, 0.f);
                }
                const int cap = p.max_iter < kMaxF32Iters ? p.max_iter : kMaxF32Iters;
                for (; it0 < cap; ++it0) {
                    float stf[6];
                    herm_iteration<float, FPT, FFT, kRT>(f00, f01, f10, f11, s00, s11, s01, 1.f, 1.f, 1.f, ZAf,
                                                         ZBf, p.plan, twsf, N, fnn, lag0f_sh, stf, &twr);
                    // update minus its constant-matrix (tail) part; the barrier inside also fences the buffers
                    const float errf2 = block_max_nonneg(stf[1], redf, phase_f);  // squared: compared with kSwitch^2
                    if (errf2 < kSwitch * kSwitch) {
                        ++it0;
                        break;
                    }
                }
                const double u0 = (double)(1.f / r0), u1 = (double)(1.f / r1);  // powers of two: exact
#pragma unroll
                for (int q = 0; q < FPT; ++q) {
                    const int f = threadIdx.x + q * kThreads;
                    if (f < fnn || !LEAN)
                        gstore(q, f, cmake<double>(f00[q].x * u0, f00[q].y * u0), cmake<double>(f01[q].x * u0, f01[q].y * u0),
                               cmake<double>(f10[q].x * u1, f10[q].y * u1), cmake<double>(f11[q].x * u1, f11[q].y * u1));
                }
                it_done = it0;
                cnt_f32 += it0;
            }
            for (int it = it0; it < p.max_iter && !converged; ++it) {
                double st[6];
                const bool defect = LEAN || MIXED || (p.tw32 && p.mixed);
                if constexpr (LEAN) {
                    herm_iteration_defect<FPT, FFT, GSmem, kRT>(gsm, s00, s11, s01, r0, r1, ZAf, ZBf, p.plan, twsf, N, fnn,
                                                                lag0f_sh, st, &twr);
                } else {
                    if (MIXED || defect)
                        herm_iteration_defect<FPT, FFT>(greg, s00, s11, s01, r0, r1, ZAf, ZBf, p.plan, twsf, N, fnn,
                                                        lag0f_sh, st);
                    else if constexpr (!MIXED)
                        herm_iteration<double, FPT, FFT>(g00, g01, g10, g11, s00, s11, s01, 1.0, 1.0, 1.0, ZA, ZB, p.plan, tws,
                                                         N, fnn, lag0_sh, st);
                }
                block_maxn_nonneg<6>(st, redd, phase_d);  // also fences ZA/ZB reuse
                const double tol2 = p.tol * p.tol;  // the maxima are squared magnitudes: compare squares, no sqrt
                it_done = it + 1;
                ++cnt_f64;
                converged = st[0] < tol2;
                if (!converged && p.tail && st[1] < kTailRest * kTailRest * tol2) {
                    // Tail in closed form.  The reference halves every lag-0 coefficient of the causal factor
                    // and THEN zeroes its lower triangle (mpd.py:132-138), so the lag-0 off-diagonal residual
                    // is only half-corrected per iteration: once every other mode has converged (the update
                    // minus its constant-matrix part is below tol), all further iterations multiply G by
                    // constant upper-triangular 2x2 matrices P_j = I + upper_half(E_j), with
                    // E_j = C_j^-1 (I + e0) C_j^-T - I and C_{j+1} = C_j P_j.  That 2x2 recursion is run here
                    // exactly, up to the iterate at which the reference stops (first max|dG| < tol).
                    // The scalar recursion is a chain of dependent fp64 divisions: one warp runs it and
                    // broadcasts the accumulated factor T (the branch is uniform over the CTA).
                    if (threadIdx.x < 32) {
                        const double e00 = defect ? (double)lag0f_sh[0] : lag0_sh[0];
                        const double e01 = defect ? (double)lag0f_sh[1] : lag0_sh[1];
                        const double e11 = defect ? (double)lag0f_sh[2] : lag0_sh[2];
                        const double m00 = 1.0 + e00, m01 = e01, m11 = 1.0 + e11;  // I + e0 (symmetric)
                        double ca = 1.0 + 0.5 * e00, cb = 0.5 * e01, cd_ = 1.0 + 0.5 * e11;  // C = P0 (already applied)
                        double ta = 1.0, tb = 0.0, td = 1.0;                                 // T = product of later P_j
                        const double n00 = sqrt(st[2]), n10 = sqrt(st[3]), n01 = sqrt(st[4]), n11 = sqrt(st[5]);
                        int itt = it_done, conv = 0;
                        while (itt < p.max_iter) {
                            // E = C^-1 M C^-T - I for upper-triangular C = [[ca, cb], [0, cd_]]
                            const double ia = 1.0 / ca, id = 1.0 / cd_, ib = -cb * ia * id;
                            const double r00 = ia * m00 + ib * m01, r01 = ia * m01 + ib * m11;  // row 0 of C^-1 M
                            const double r11 = id * m11;                                          // row 1: (id*m01, id*m11)
                            const double E00 = r00 * ia + r01 * ib - 1.0, E01 = r01 * id, E11 = r11 * id - 1.0;
                            const double pa = 0.5 * E00, pb = 0.5 * E01, pd = 0.5 * E11;  // P_j - I
                            // max |G_j (P_j - I)| with the column maxima of G at tail entry
                            const double d0 = fmax(n00, n10) * fabs(pa);
                            const double d1 = fmax(n00 * fabs(pb) + n01 * fabs(pd), n10 * fabs(pb) + n11 * fabs(pd));
                            // T <- T P_j, C <- C P_j
                            tb = ta * pb + tb * (1.0 + pd); ta *= 1.0 + pa; td *= 1.0 + pd;
                            cb = ca * pb + cb * (1.0 + pd); ca *= 1.0 + pa; cd_ *= 1.0 + pd;
                            ++itt;
                            if (fmax(d0, d1) < p.tol) {
                                conv = 1;
                                break;
                            }
#ifndef SC_GRANGER_NO_TAIL_JUMP
                            // Late tail in closed form.  The recursion decouples: cd' = (cd + m11 / cd) / 2 and
                            // ca' = (ca + q / ca) / 2 are Newton square roots (quadratic), while r = cb / cd obeys
                            // r' = r + (m01 / m11 - r) / 2 once cd^2 = m11, i.e. the off-diagonal defect HALVES exactly
                            // and q = q* + m11 (r - m01 / m11)^2 drags ca along one step behind: pa' = -(3/2) pb^2.
                            // So as soon as the diagonal steps are small, every further step is pb <- pb / 2,
                            // pa <- -6 pb^2, pd = 0: a handful of multiplies instead of two divisions and thirty
                            // dependent operations, stopping at the same iterate (the test values halve exactly like
                            // the recursion's own).
                            if (fabs(pa) < SC_TAIL_JUMP_PA && fabs(pd) < SC_TAIL_JUMP_PD) {
                                double pbm = pb, pam;
                                const double nmax = fmax(n00, n10);
                                while (itt < p.max_iter) {
                                    pbm *= 0.5;
                                    pam = -6.0 * pbm * pbm;  // = -(3/2) pb_prev^2: ca follows sqrt(q(r)) one step behind
                                    tb = fma(ta, pbm, tb);
                                    ta = fma(ta, pam, ta);
                                    ++itt;
                                    if (nmax * fmax(fabs(pam), fabs(pbm)) < p.tol) {
                                        conv = 1;
                                        break;
                                    }
                                }
                                break;
                            }
#endif
                        }
                        if (threadIdx.x == 0) {
                            tail_sh[0] = ta; tail_sh[1] = tb; tail_sh[2] = td;
                            tail_it[0] = itt; tail_it[1] = conv;
                        }
                    }
                    __syncthreads();
                    const double ta = tail_sh[0], tb = tail_sh[1], td = tail_sh[2];
                    cnt_tail += tail_it[0] - it_done;
                    it_done = tail_it[0];
                    converged = tail_it[1] != 0;
#pragma unroll
                    for (int q = 0; q < FPT; ++q) {
                        const int f = threadIdx.x + q * kThreads;
                        if (f >= fnn && LEAN) continue;
                        cd A, B, C, D;
                        gload(q, f, A, B, C, D);
                        B.x = A.x * tb + B.x * td; B.y = A.y * tb + B.y * td;
                        D.x = C.x * tb + D.x * td; D.y = C.y * tb + D.y * td;
                        A.x *= ta; A.y *= ta; C.x *= ta; C.y *= ta;
                        gstore(q, f, A, B, C, D);
                    }
                    break;
                }
            }
            if (!converged) flag |= SC_FLAG_NOT_CONVERGED;
        }
        ++cnt_prob;
        if (threadIdx.x == 0) {
            if (p.iters) p.iters[pk * p.B + b] = it_done;
            if (p.flags) p.flags[pk * p.B + b] = flag;
        }
        // the spectrum registers are free now: fetch the next problem's S under the epilogue
        const int cpi = pi, cpj = pj;
        const long long cb = b;
        if (!GROUPED && FFT::kPrefetch && next_prob(prob) < nprob) {
            pair_of(next_prob(prob), nb_, npk, npi, npj);
            load_s(nb_, npi, npj);
        }
        float* out = reinterpret_cast<float*>(p.out);
        if (flag & SC_FLAG_NOT_SPD) {
            for (int f = threadIdx.x; f < fnn; f += kThreads) {
                float* m = out + ((size_t)cb * fnn + f) * p.S * p.S;
                if constexpr (GROUPED) reinterpret_cast<float*>(stOut + f)[cpj - j0] = fnan;
                else m[(size_t)cpi * p.S + cpj] = fnan;
                m[(size_t)cpj * p.S + cpi] = fnan;
            }
        } else {
        // ---- Granger epilogue (connectivity.py:1705-1709, 1739-1748, 1847-1848, 1773-1779) ----
        float pw_i[FPT], pw_j[FPT];
        double h[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll
        for (int q = 0; q < FPT; ++q) {
            const int f = threadIdx.x + q * kThreads;
            pw_i[q] = 0.f; pw_j[q] = 0.f;
            if (f < fnn) {
                if constexpr (GROUPED) {
                    pw_i[q] = stPi[f];
                    pw_j[q] = reinterpret_cast<const float*>(stPj + f)[cpj - j0];
                } else {
                    pw_i[q] = __ldg(&p.power[((size_t)cb * p.F + f) * p.S + cpi]);
                    pw_j[q] = __ldg(&p.power[((size_t)cb * p.F + f) * p.S + cpj]);
                }
                const double w = (f == 0 || 2 * f == N) ? 1.0 : 2.0;
                cd A, B, C, D;
                gload(q, f, A, B, C, D);
                h[0] += w * A.x; h[1] += w * B.x; h[2] += w * C.x; h[3] += w * D.x;
            }
        }
        block_sum4(h, red4, phase_s);
        const double h00 = h[0] * inv_nd, h01 = h[1] * inv_nd, h10 = h[2] * inv_nd, h11 = h[3] * inv_nd;
        const double lam = kTikhonov * (h00 * h00 + h01 * h01 + h10 * h10 + h11 * h11) * 0.25;
        const double m00 = h00 + lam, m11 = h11 + lam;
        const double mdet = m00 * m11 - h01 * h10;
        const double imdet = 1.0 / mdet;
        const double v00 = m11 * imdet, v01 = -h01 * imdet, v10 = -h10 * imdet, v11 = m00 * imdet;
        const double c00 = h00 * h00 + h01 * h01, c01 = h00 * h10 + h01 * h11, c11 = h10 * h10 + h11 * h11;
        // r01 = c11 - c01^2 / c00 and r10 = c00 - c01^2 / c11 share the numerator det(Sigma): one division
        const double sdet = c00 * c11 - c01 * c01, icc = 1.0 / (c00 * c11);
        const double r01 = sdet * c11 * icc;
        const double r10 = sdet * c00 * icc;
#pragma unroll
        for (int q = 0; q < FPT; ++q) {
            const int f = threadIdx.x + q * kThreads;
            if (f < fnn) {
                cd A, B, C, D;
                gload(q, f, A, B, C, D);
                const cd t01 = cmake<double>(A.x * v01 + B.x * v11, A.y * v01 + B.y * v11);
                const cd t10 = cmake<double>(C.x * v00 + D.x * v10, C.y * v00 + D.y * v10);
                const float gc01 = log_ratio((double)pw_i[q], r01 * (t01.x * t01.x + t01.y * t01.y));
                const float gc10 = log_ratio((double)pw_j[q], r10 * (t10.x * t10.x + t10.y * t10.y));
                float* m = out + ((size_t)cb * fnn + f) * p.S * p.S;
                if constexpr (GROUPED) reinterpret_cast<float*>(stOut + f)[cpj - j0] = gc01;
                else m[(size_t)cpi * p.S + cpj] = gc01;
                m[(size_t)cpj * p.S + cpi] = gc10;
            }
        }
        }  // !NOT_SPD
        // ---- next problem ----
        if constexpr (GROUPED) {
            if (++pj > j0 + 3) {
                flush_group();
                grp += gridDim.x;
                have = grp < ngroups;
                if (have) {
                    begin_group(grp);
                    pj = jlo;
                }
            }
        } else {
            prob = next_prob(prob);
            b = nb_; pk = npk; pi = npi; pj = npj;
            have = prob < nprob;
        }
    }
    if (p.exec_counters && threadIdx.x == 0) {
        atomicAdd(p.exec_counters + 0, cnt_f32);
        atomicAdd(p.exec_counters + 1, cnt_f64);
        atomicAdd(p.exec_counters + 2, cnt_tail);
        atomicAdd(p.exec_counters + 3, cnt_prob);
    }
}

size_t herm_smem(int nfft) { return (size_t)5 * nfft * sizeof(cd) + (size_t)nfft * 8; }  // ZA, ZB, twiddles (f64 + f32)

template <int FPT, typename FFT, bool LEAN, bool MIXED = false, bool GROUPED = false>
int herm_launch_as(W2Params& p, cudaStream_t st) {
    const int fnn = p.nfft / 2 + 1;
    const size_t f32 = (((size_t)4 * p.nfft + (LEAN && FFT::kRegTw ? 0 : FFT::tw_entries(p.nfft))) * sizeof(cx<float>) + 15) & ~(size_t)15;
    const size_t smem = GROUPED ? f32 + (size_t)fnn * 88  // staging: 5 float4 + 2 float per bin
                        : MIXED ? f32 : LEAN ? f32 + (size_t)4 * fnn * sizeof(cd)
                             : (size_t)(4 * p.nfft + FFT::tw_entries(p.nfft)) * sizeof(cd) +
                                   (size_t)FFT::tw_entries(p.nfft) * sizeof(cx<float>);
    if (smem > 48 * 1024)
        SC_CUDA_OK(cudaFuncSetAttribute(granger_herm_kernel<FPT, FFT, LEAN, MIXED, GROUPED>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        (int)smem));
    if (const char* e = getenv("SC_GRANGER_CARVEOUT"))  // experiment hook: shared-memory carveout in percent
        SC_CUDA_OK(cudaFuncSetAttribute(granger_herm_kernel<FPT, FFT, LEAN, MIXED, GROUPED>, cudaFuncAttributePreferredSharedMemoryCarveout,
                                        atoi(e)));
    int per_sm = 1;
    SC_CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, granger_herm_kernel<FPT, FFT, LEAN, MIXED, GROUPED>, kThreads, smem));
    if (per_sm < 1) per_sm = 1;
    const long long nprob = p.B * p.n_pairs;
    long long grid = (long long)sc_num_sms() * per_sm;
    const long long ngroups = GROUPED ? nprob / 4 + 1 : (nprob + kGroup - 1) / kGroup;
    if (grid > ngroups) grid = ngroups;
    granger_herm_kernel<FPT, FFT, LEAN, MIXED, GROUPED><<<(unsigned)grid, kThreads, smem, st>>>(p);
    SC_LAUNCH_OK();
    return SC_OK;
}

template <int FPT, typename FFT>
int herm_launch(W2Params& p, cudaStream_t st) {
    // mixed-precision mode (needs the fp32 twiddles) with a plan whose twiddles fit in registers -> the lean kernel
    // (fp64 factor in shared memory, no twiddle tables); everything else -> the register kernel
    // Measured (profiles/r02_granger_experiments.txt): the lean kernel runs config 4 in 305.5 ms, the register kernel
    // in 302.6 ms -- the iteration is bound by the lock-step alternation of load / butterfly / store phases between
    // barriers, not by shared-memory wavefronts, registers or the fp64 pipe -- so the simpler register kernel stays
    // the default and the lean one is kept behind SC_GRANGER_LEAN=1 for the record.
    const char* lean = getenv("SC_GRANGER_LEAN");
    if (FFT::kRegTw && p.tw32 && p.mixed && lean && lean[0] == '1') return herm_launch_as<FPT, FFT, true>(p, st);
    if (p.tw32 && p.mixed) {
        // all pairs of a signal count that keeps the 16-byte row segments aligned -> grouped problem order
        const char* ng = getenv("SC_GRANGER_NO_GROUPS");
        const bool aligned = ((reinterpret_cast<uintptr_t>(p.csm) | reinterpret_cast<uintptr_t>(p.power) |
                               reinterpret_cast<uintptr_t>(p.out)) & 15) == 0;   // 16-byte row segments
        if (!p.pairs && p.S % 4 == 0 && p.S >= 8 && p.n_pairs == p.S * (p.S - 1) / 2 && aligned && !(ng && ng[0] == '1'))
            return herm_launch_as<FPT, FFT, false, true, true>(p, st);
        return herm_launch_as<FPT, FFT, false, true>(p, st);
    }
    return herm_launch_as<FPT, FFT, false>(p, st);
}

}  // namespace

int sc_granger_herm_supported(int nfft) {
    const int fnn = nfft / 2 + 1;
    return fnn <= 4 * scw::kThreads && herm_smem(nfft) + 4096 <= (size_t)sc_max_smem_optin();
}

int sc_granger_herm_launch(scw::W2Params& p, void* stream) {
    if (sc_fft_make_plan(p.nfft, &p.plan)) {
        sc_set_error("granger: cannot factorise nfft=%d", p.nfft);
        return SC_ERR_UNSUPPORTED;
    }
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    if (p.nfft == 1000) return herm_launch<(501 + scw::kThreads - 1) / scw::kThreads, StatFft<ScPlan1000>>(p, st);
    if (p.nfft == 120) return herm_launch<1, StatFft<ScPlan120>>(p, st);
    const int fpt = (p.nfft / 2 + 1 + scw::kThreads - 1) / scw::kThreads;
    switch (fpt) {
        case 1: return herm_launch<1, DynFft>(p, st);
        case 2: return herm_launch<2, DynFft>(p, st);
        case 3: return herm_launch<3, DynFft>(p, st);
        default: return herm_launch<4, DynFft>(p, st);
    }
}
