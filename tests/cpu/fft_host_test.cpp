// Host build of the device FFT butterflies (tests only): runs every Stockham stage
// sequentially so numpy can check the index arithmetic without a GPU.
#include <vector>
#include <cmath>
#include "../../spectral_connectivity_b200/csrc/fft_device.cuh"

template <typename R, int RADIX>
static void pass(const cx<R>* src, cx<R>* dst, int n, int Ls, const cx<R>* tw, bool inv) {
    for (int j = 0; j < n / RADIX; ++j) sc_fft_item<R, RADIX>(src, dst, j, n, Ls, tw, inv);
}

template <typename R>
static int run(R* data, int n, int inverse, int force_generic) {
    ScFftPlan plan;
    if (sc_fft_make_plan(n, &plan)) return -1;
    std::vector<cx<R>> a(n), b(n), tw(n);
    for (int i = 0; i < n; ++i) {
        a[i].x = data[2 * i];
        a[i].y = data[2 * i + 1];
        tw[i].x = (R)cos(-2.0 * M_PI * i / n);
        tw[i].y = (R)sin(-2.0 * M_PI * i / n);
    }
    cx<R>* src = a.data();
    cx<R>* dst = b.data();
    int Ls = 1;
    const bool inv = inverse != 0;
    for (int s = 0; s < plan.nstages; ++s) {
        const int r = plan.radix[s];
        int rr = force_generic ? -1 : r;
        switch (rr) {
            case 2: pass<R, 2>(src, dst, n, Ls, tw.data(), inv); break;
            case 3: pass<R, 3>(src, dst, n, Ls, tw.data(), inv); break;
            case 4: pass<R, 4>(src, dst, n, Ls, tw.data(), inv); break;
            case 5: pass<R, 5>(src, dst, n, Ls, tw.data(), inv); break;
            case 7: pass<R, 7>(src, dst, n, Ls, tw.data(), inv); break;
            case 8: pass<R, 8>(src, dst, n, Ls, tw.data(), inv); break;
            case 10: pass<R, 10>(src, dst, n, Ls, tw.data(), inv); break;
            case 11: pass<R, 11>(src, dst, n, Ls, tw.data(), inv); break;
            case 13: pass<R, 13>(src, dst, n, Ls, tw.data(), inv); break;
            default:
                for (int it = 0; it < n; ++it) sc_fft_item_generic<R>(src, dst, it, r, n, Ls, tw.data(), inv);
        }
        Ls *= r;
        std::swap(src, dst);
    }
    for (int i = 0; i < n; ++i) {
        data[2 * i] = src[i].x;
        data[2 * i + 1] = src[i].y;
    }
    return plan.nstages;
}

extern "C" int fft_host_f64(double* data, int n, int inverse, int force_generic) {
    return run<double>(data, n, inverse, force_generic);
}
extern "C" int fft_host_f32(float* data, int n, int inverse, int force_generic) {
    return run<float>(data, n, inverse, force_generic);
}
template <typename PLAN>
static int run_static(double* data, int inverse) {
    typedef ScStaticFft<double, PLAN> F;
    const int n = F::n;
    std::vector<cx<double>> a(n), b(n), tw(n), tws(F::tw_count + 1);
    for (int i = 0; i < n; ++i) {
        a[i].x = data[2 * i];
        a[i].y = data[2 * i + 1];
        tw[i].x = cos(-2.0 * M_PI * i / n);
        tw[i].y = sin(-2.0 * M_PI * i / n);
    }
    F::fill(tws.data(), tw.data(), 0, 1);
    cx<double>* res = F::template run<1>(a.data(), b.data(), tws.data(), inverse != 0, 0, 1, [] {});
    for (int i = 0; i < n; ++i) {
        data[2 * i] = res[i].x;
        data[2 * i + 1] = res[i].y;
    }
    return F::tw_count;
}

extern "C" int fft_static_f64(double* data, int n, int inverse) {
    if (n == 1000) return run_static<ScPlan1000>(data, inverse);
    if (n == 120) return run_static<ScPlan120>(data, inverse);
    return -1;
}

// inverse FFT -> window (w[n] for n < 500, else 0; batch 1 additionally drops the imaginary part at n = 0)
// -> forward FFT of 2 sequences of length 1000 through the fused three-stage path
template <typename R> struct TestWin {
    const R* w;
    SC_HD cx<R> operator()(int bb, int n, cx<R> v) const {
        cx<R> r = cmake<R>(v.x * w[n], v.y * w[n]);
        if (bb == 1 && n == 0) r.y = (R)0;
        return r;
    }
};
template <typename R, bool POW>
static int run_conv(R* data, const R* w) {
    const int n = 1000;
    typedef ScStaticFft<R, ScPlan1000> F;
    typedef ScStaticConv<R, ScPlan1000, 500, POW> C;
    static_assert(C::supported, "plan 1000 supports the fused path");
    static_assert(!ScStaticConv<R, ScPlan120, 60, POW>::supported, "plan 120 does not");
    std::vector<cx<R>> a(2 * n), b(2 * n), tw(n), tws(F::tw_count + 1);
    for (int i = 0; i < 2 * n; ++i) {
        a[i].x = data[2 * i];
        a[i].y = data[2 * i + 1];
    }
    for (int i = 0; i < n; ++i) {
        tw[i].x = (R)cos(-2.0 * M_PI * i / n);
        tw[i].y = (R)sin(-2.0 * M_PI * i / n);
    }
    F::fill(tws.data(), tw.data(), 0, 1);
    TestWin<R> win{w};
    cx<R>* res = C::template run<2>(a.data(), b.data(), tws.data(), 0, 1, [] {}, win);
    for (int i = 0; i < 2 * n; ++i) {
        data[2 * i] = res[i].x;
        data[2 * i + 1] = res[i].y;
    }
    return 0;
}
extern "C" int fft_conv_f64(double* data, const double* w, int pow_twiddles) {
    return pow_twiddles ? run_conv<double, true>(data, w) : run_conv<double, false>(data, w);
}
extern "C" int fft_conv_f32(float* data, const float* w, int pow_twiddles) {
    return pow_twiddles ? run_conv<float, true>(data, w) : run_conv<float, false>(data, w);
}

extern "C" int fft_plan(int n, int* radices) {
    ScFftPlan plan;
    if (sc_fft_make_plan(n, &plan)) return -1;
    for (int i = 0; i < plan.nstages; ++i) radices[i] = plan.radix[i];
    return plan.nstages;
}
