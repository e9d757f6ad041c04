// Wilson spectral factorisation of 2x2 cross-spectral matrices and the fused pairwise
// spectral-Granger epilogue (general two-sided complex path, fp64).
//
// Replaces minimum_phase_decomposition.py:227-322 (minimum_phase_decomposition) with its
// helpers _get_initial_conditions (:48-93), _get_linear_predictor (:184-224),
// _get_causal_signal (:96-142), _check_convergence (:145-181); and, in Granger mode,
// connectivity.py:2282-2340 (_estimate_spectral_granger_prediction) with
// _estimate_transfer_function (:1712-1748), _estimate_noise_covariance (:1679-1709),
// _remove_instantaneous_causality (:1825-1848), _estimate_predictive_power (:1751-1779).
//
// The reference runs a Python loop over S(S-1)/2 pairs, each iteration issuing two batched
// LAPACK solves, an ifft, an fft and a batched matmul, with a host sync per iteration.  Here one
// CTA owns one (pair, window) problem end to end: the factor G(f), the linear predictor and the
// FFT ping-pong buffers stay in shared memory (or an L2-resident workspace for long FFTs), the
// 2x2 solves are closed form, convergence is decided on the device, and the Granger log-ratio is
// written straight into the (B, Fnn, S, S) output.
#include "wilson_common.cuh"

int sc_granger_herm_supported(int nfft);
int sc_granger_herm_launch(scw::W2Params& p, void* stream);

namespace {

using namespace scw;

// MODE 0: csm c128 [B][nfft][2][2] -> G c128.  MODE 1: Granger over pairs of a c64 [B][F][S][S] CSM.
template <int MODE>
__global__ void __launch_bounds__(kThreads) wilson2_kernel(const W2Params p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ double red[4 * kWarps];
    const int nfft = p.nfft;
    const size_t nf = (size_t)nfft;
    unsigned char* wsb = p.ws + (size_t)blockIdx.x * (p.use_smem ? 4 : 16) * nf * sizeof(cd);
    cd* Sm = reinterpret_cast<cd*>(wsb);  // [4][nfft] in the workspace (L2)
    cd *G, *Wa, *Wb;
    const cd* tw;
    if (p.use_smem) {
        G = reinterpret_cast<cd*>(smem_raw);
        Wa = G + 4 * nf;
        Wb = Wa + 4 * nf;
        cd* tws = Wb + 4 * nf;
        for (int q = threadIdx.x; q < nfft; q += kThreads) tws[q] = p.tw[q];
        tw = tws;
    } else {
        G = Sm + 4 * nf;
        Wa = G + 4 * nf;
        Wb = Wa + 4 * nf;
        tw = p.tw;
    }
    const long long npairs = MODE == 1 ? p.n_pairs : 1;
    const long long nprob = p.B * npairs;
    const int fnn = nfft / 2 + 1;
    const double qnan = __longlong_as_double(0x7ff8000000000000LL);

    for (long long prob = blockIdx.x; prob < nprob; prob += gridDim.x) {
        const long long b = prob / npairs;
        const long long pk = prob % npairs;
        int pi = 0, pj = 1;
        if (MODE == 1) {
            if (p.pairs) {
                pi = p.pairs[2 * pk];
                pj = p.pairs[2 * pk + 1];
            } else {
                decode_pair(pk, p.S, pi, pj);
            }
        }
        __syncthreads();  // previous problem's readers of Sm/G are done
        // ---- 1. load S(f), accumulate the lag-0 covariance -------------------
        double a[3] = {0.0, 0.0, 0.0};
        for (int f = threadIdx.x; f < nfft; f += kThreads) {
            cd s00, s01, s10, s11;
            if (MODE == 0) {
                const cd* src = reinterpret_cast<const cd*>(p.csm) + ((size_t)b * nf + f) * 4;
                s00 = src[0]; s01 = src[1]; s10 = src[2]; s11 = src[3];
            } else {
                int ff = f;
                bool mirror = false;
                if (p.herm && f > nfft / 2) {
                    ff = nfft - f;
                    mirror = true;
                }
                const float2* m = reinterpret_cast<const float2*>(p.csm) + ((size_t)b * p.F + ff) * p.S * p.S;
                const float2 v00 = m[(size_t)pi * p.S + pi], v01 = m[(size_t)pi * p.S + pj];
                const float2 v10 = m[(size_t)pj * p.S + pi], v11 = m[(size_t)pj * p.S + pj];
                const double sg = mirror ? -1.0 : 1.0;
                s00 = cmake<double>(v00.x, sg * v00.y);
                s01 = cmake<double>(v01.x, sg * v01.y);
                s10 = cmake<double>(v10.x, sg * v10.y);
                s11 = cmake<double>(v11.x, sg * v11.y);
            }
            Sm[f] = s00; Sm[nf + f] = s01; Sm[2 * nf + f] = s10; Sm[3 * nf + f] = s11;
            a[0] += s00.x; a[1] += s10.x; a[2] += s11.x;
        }
        block_sum<3>(a, red);
        // ---- 2. Cholesky of the real lag-0 matrix, G0 = L^T (mpd.py:75-77) ---
        const double a00 = a[0] / nfft, a10 = a[1] / nfft, a11 = a[2] / nfft;
        const double l00 = sqrt(a00);
        const double l10 = a10 / l00;
        const double d11 = a11 - l10 * l10;
        const double l11 = sqrt(d11);
        int flag = 0;
        int it_done = 0;
        if (!(a00 > 0.0) || !(d11 > 0.0) || !isfinite(l00) || !isfinite(l11)) flag = SC_FLAG_NOT_SPD;
        if (!flag) {
            for (int f = threadIdx.x; f < nfft; f += kThreads) {
                G[f] = cmake<double>(l00, 0.0);
                G[nf + f] = cmake<double>(l10, 0.0);
                G[2 * nf + f] = cmake<double>(0.0, 0.0);
                G[3 * nf + f] = cmake<double>(l11, 0.0);
            }
            __syncthreads();
            // ---- 3. Wilson iterations -----------------------------------------
            bool converged = false;
            const double inv_n = 1.0 / nfft;
            const int kcut = (nfft + 1) / 2;
            for (int it = 0; it < p.max_iter && !converged; ++it) {
                // linear predictor B = G^-1 S G^-H + I (mpd.py:218-224)
                for (int f = threadIdx.x; f < nfft; f += kThreads) {
                    const cd g00 = G[f], g01 = G[nf + f], g10 = G[2 * nf + f], g11 = G[3 * nf + f];
                    const cd idet = cdiv1(csub(cmul(g00, g11), cmul(g01, g10)));
                    const cd i00 = cmul(g11, idet), i01 = cneg(cmul(g01, idet));
                    const cd i10 = cneg(cmul(g10, idet)), i11 = cmul(g00, idet);
                    const cd s00 = Sm[f], s01 = Sm[nf + f], s10 = Sm[2 * nf + f], s11 = Sm[3 * nf + f];
                    const cd y00 = cadd(cmul(i00, s00), cmul(i01, s10));
                    const cd y01 = cadd(cmul(i00, s01), cmul(i01, s11));
                    const cd y10 = cadd(cmul(i10, s00), cmul(i11, s10));
                    const cd y11 = cadd(cmul(i10, s01), cmul(i11, s11));
                    cd b00 = cadd(cmulc(i00, y00), cmulc(i01, y01));
                    const cd b01 = cadd(cmulc(i00, y10), cmulc(i01, y11));
                    const cd b10 = cadd(cmulc(i10, y00), cmulc(i11, y01));
                    cd b11 = cadd(cmulc(i10, y10), cmulc(i11, y11));
                    b00.x += 1.0;
                    b11.x += 1.0;
                    Wa[f] = b00; Wa[nf + f] = b01; Wa[2 * nf + f] = b10; Wa[3 * nf + f] = b11;
                }
                __syncthreads();
                // plus operator (mpd.py:129-142)
                cd* c = sc_cta_fft<double>(Wa, Wb, 4, nfft, p.plan, tw, true);
                cd* other = (c == Wa) ? Wb : Wa;
                for (int idx = threadIdx.x; idx < 4 * nfft; idx += kThreads) {
                    const int e = idx / nfft, k = idx - e * nfft;
                    cd v = c[idx];
                    double sc = inv_n;
                    if (k == 0) sc = (e == 2) ? 0.0 : 0.5 * inv_n;
                    if (k >= kcut) sc = 0.0;
                    c[idx] = sc == 0.0 ? cmake<double>(0.0, 0.0) : cscale(v, sc);
                }
                __syncthreads();
                const cd* P = sc_cta_fft<double>(c, other, 4, nfft, p.plan, tw, false);
                // G <- G P, convergence on max |dG| (mpd.py:305-315)
                double err2 = 0.0;
                for (int f = threadIdx.x; f < nfft; f += kThreads) {
                    const cd g00 = G[f], g01 = G[nf + f], g10 = G[2 * nf + f], g11 = G[3 * nf + f];
                    const cd p00 = P[f], p01 = P[nf + f], p10 = P[2 * nf + f], p11 = P[3 * nf + f];
                    const cd n00 = cadd(cmul(g00, p00), cmul(g01, p10));
                    const cd n01 = cadd(cmul(g00, p01), cmul(g01, p11));
                    const cd n10 = cadd(cmul(g10, p00), cmul(g11, p10));
                    const cd n11 = cadd(cmul(g10, p01), cmul(g11, p11));
                    cd d;
                    d = csub(n00, g00); err2 = fmax(err2, d.x * d.x + d.y * d.y);
                    d = csub(n01, g01); err2 = fmax(err2, d.x * d.x + d.y * d.y);
                    d = csub(n10, g10); err2 = fmax(err2, d.x * d.x + d.y * d.y);
                    d = csub(n11, g11); err2 = fmax(err2, d.x * d.x + d.y * d.y);
                    G[f] = n00; G[nf + f] = n01; G[2 * nf + f] = n10; G[3 * nf + f] = n11;
                }
                const double err = sqrt(block_max(err2, red));
                it_done = it + 1;
                // NaN error never converges (NaN < tol is false), like the reference
                converged = err < p.tol;
                __syncthreads();
            }
            if (!converged) flag |= SC_FLAG_NOT_CONVERGED;
        }
        if (threadIdx.x == 0) {
            if (p.iters) p.iters[(MODE == 1 ? pk * p.B + b : b)] = it_done;
            if (p.flags) p.flags[(MODE == 1 ? pk * p.B + b : b)] = flag;
        }
        // ---- 4. epilogue ------------------------------------------------------
        if (MODE == 0) {
            cd* dst = reinterpret_cast<cd*>(p.out) + (size_t)b * nf * 4;
            for (int idx = threadIdx.x; idx < 4 * nfft; idx += kThreads) {
                const int f = idx >> 2, e = idx & 3;
                dst[idx] = (flag & SC_FLAG_NOT_SPD) ? cmake<double>(qnan, qnan) : G[(size_t)e * nf + f];
            }
        } else {
            float* out = reinterpret_cast<float*>(p.out);
            const float fnan = __int_as_float(0x7fc00000);
            if (flag & SC_FLAG_NOT_SPD) {
                for (int f = threadIdx.x; f < fnn; f += kThreads) {
                    float* m = out + ((size_t)b * fnn + f) * p.S * p.S;
                    m[(size_t)pi * p.S + pj] = fnan;
                    m[(size_t)pj * p.S + pi] = fnan;
                }
                continue;
            }
            // H0 = Re ifft(G)[lag 0] (connectivity.py:1739-1740)
            double h[4] = {0.0, 0.0, 0.0, 0.0};
            for (int f = threadIdx.x; f < nfft; f += kThreads) {
                h[0] += G[f].x; h[1] += G[nf + f].x; h[2] += G[2 * nf + f].x; h[3] += G[3 * nf + f].x;
            }
            block_sum<4>(h, red);
            const double h00 = h[0] / nfft, h01 = h[1] / nfft, h10 = h[2] / nfft, h11 = h[3] / nfft;
            // Tikhonov-regularised inverse (:1742-1747); lambda from this problem's H0 only
            const double lam = kTikhonov * (h00 * h00 + h01 * h01 + h10 * h10 + h11 * h11) * 0.25;
            const double m00 = h00 + lam, m11 = h11 + lam;
            const double mdet = m00 * m11 - h01 * h10;
            const double v00 = m11 / mdet, v01 = -h01 / mdet, v10 = -h10 / mdet, v11 = m00 / mdet;
            // noise covariance H0 H0^T (:1705-1709) and rotated covariance (:1847-1848)
            const double c00 = h00 * h00 + h01 * h01, c01 = h00 * h10 + h01 * h11, c11 = h10 * h10 + h11 * h11;
            const double r01 = c11 - c01 * c01 / c00;  // R[0][1]
            const double r10 = c00 - c01 * c01 / c11;  // R[1][0]
            for (int f = threadIdx.x; f < fnn; f += kThreads) {
                const cd g00 = G[f], g01 = G[nf + f], g10 = G[2 * nf + f], g11 = G[3 * nf + f];
                // H = G Minv: only the off-diagonal entries are needed
                const cd t01 = cmake<double>(g00.x * v01 + g01.x * v11, g00.y * v01 + g01.y * v11);
                const cd t10 = cmake<double>(g10.x * v00 + g11.x * v10, g10.y * v00 + g11.y * v10);
                const double pw_i = p.power[((size_t)b * p.F + f) * p.S + pi];
                const double pw_j = p.power[((size_t)b * p.F + f) * p.S + pj];
                double in01 = pw_i - r01 * (t01.x * t01.x + t01.y * t01.y);
                double in10 = pw_j - r10 * (t10.x * t10.x + t10.y * t10.y);
                if (in01 == 0.0) in01 = kEps64;  // connectivity.py:1776
                if (in10 == 0.0) in10 = kEps64;
                double gc01 = log(pw_i) - log(in01);
                double gc10 = log(pw_j) - log(in10);
                if (gc01 <= 0.0) gc01 = qnan;  // :1778
                if (gc10 <= 0.0) gc10 = qnan;
                float* m = out + ((size_t)b * fnn + f) * p.S * p.S;
                m[(size_t)pi * p.S + pj] = (float)gc01;
                m[(size_t)pj * p.S + pi] = (float)gc10;
            }
        }
    }
}

struct W2Config {
    int use_smem;
    size_t smem;
    int ctas;
    size_t ws_per_cta;
};

W2Config w2_config(int nfft) {
    W2Config c;
    const size_t need = (size_t)13 * nfft * sizeof(cd);  // G, Wa, Wb (4 each) + twiddles
    const size_t cap = (size_t)sc_max_smem_optin() - 2048;
    c.use_smem = need <= cap;
    c.smem = c.use_smem ? need : 0;
    const int per_sm = c.use_smem ? (int)(cap / need > 4 ? 4 : cap / need) : 2;
    c.ctas = sc_num_sms() * (per_sm < 1 ? 1 : per_sm);
    c.ws_per_cta = (size_t)(c.use_smem ? 4 : 16) * nfft * sizeof(cd);
    return c;
}

template <int MODE>
int w2_launch(W2Params& p, long long nprob, void* workspace, int64_t workspace_bytes, cudaStream_t st) {
    if (sc_fft_make_plan(p.nfft, &p.plan)) {
        sc_set_error("wilson: cannot factorise nfft=%d", p.nfft);
        return SC_ERR_UNSUPPORTED;
    }
    const W2Config c = w2_config(p.nfft);
    long long grid = nprob < c.ctas ? nprob : c.ctas;
    const int64_t need = (int64_t)c.ctas * (int64_t)c.ws_per_cta;
    if (!workspace || workspace_bytes < (int64_t)grid * (int64_t)c.ws_per_cta) {
        sc_set_error("wilson: workspace of %lld bytes required (sc_wilson_workspace_bytes), got %lld", (long long)need,
                     (long long)workspace_bytes);
        return SC_ERR_WORKSPACE;
    }
    p.ws = reinterpret_cast<unsigned char*>(workspace);
    p.use_smem = c.use_smem;
    if (c.smem > 48 * 1024)
        SC_CUDA_OK(cudaFuncSetAttribute(wilson2_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)c.smem));
    wilson2_kernel<MODE><<<(unsigned)grid, kThreads, c.smem, st>>>(p);
    SC_LAUNCH_OK();
    return SC_OK;
}

}  // namespace

extern "C" int64_t sc_wilson_workspace_bytes(int nfft) {
    if (nfft < 1) return 0;
    const W2Config c = w2_config(nfft);
    return (int64_t)c.ctas * (int64_t)c.ws_per_cta;
}

extern "C" int sc_wilson2(const void* csm_c128, int64_t B, int nfft, double tolerance, int max_iterations,
                          const void* twiddle, void* out_g_c128, int* out_iters, int* out_flags, void* workspace,
                          int64_t workspace_bytes, void* stream) {
    SC_CHECK_ARG(csm_c128 && twiddle && out_g_c128, "sc_wilson2: null pointer");
    SC_CHECK_ARG(B > 0 && nfft > 0 && max_iterations >= 0, "sc_wilson2: bad size");
    W2Params p = {};
    p.csm = csm_c128; p.power = nullptr; p.B = B; p.F = nfft; p.nfft = nfft; p.herm = 0; p.S = 2;
    p.pairs = nullptr; p.n_pairs = 1; p.tol = tolerance; p.max_iter = max_iterations;
    p.tw = reinterpret_cast<const cd*>(twiddle); p.out = out_g_c128; p.iters = out_iters; p.flags = out_flags;
    return w2_launch<0>(p, B, workspace, workspace_bytes, reinterpret_cast<cudaStream_t>(stream));
}

extern "C" int sc_granger_pairwise(const void* csm_c64, const float* power, int64_t B, int F, int nfft,
                                   int hermitian_half, int64_t S, const int* pairs, int64_t n_pairs, double tolerance,
                                   int max_iterations, int tail_extrapolation, int mixed_precision,
                                   const void* twiddle_c128, const void* twiddle_c64, float* out_gc, int* out_iters,
                                   int* out_flags, uint64_t* out_exec_counters, void* workspace,
                                   int64_t workspace_bytes, void* stream) {
    SC_CHECK_ARG(csm_c64 && power && twiddle_c128 && out_gc, "sc_granger_pairwise: null pointer");
    SC_CHECK_ARG(B > 0 && nfft > 0 && S >= 2 && max_iterations >= 0, "sc_granger_pairwise: bad size");
    SC_CHECK_ARG(hermitian_half ? F == nfft / 2 + 1 : F == nfft,
                 "sc_granger_pairwise: F=%d inconsistent with nfft=%d (hermitian_half=%d)", F, nfft, hermitian_half);
    if (!pairs) n_pairs = S * (S - 1) / 2;
    SC_CHECK_ARG(n_pairs > 0, "sc_granger_pairwise: no pairs");
    W2Params p = {};
    p.csm = csm_c64; p.power = power; p.B = B; p.F = F; p.nfft = nfft; p.herm = hermitian_half; p.S = S;
    p.pairs = pairs; p.n_pairs = n_pairs; p.tol = tolerance; p.max_iter = max_iterations;
    p.tw = reinterpret_cast<const cd*>(twiddle_c128); p.out = out_gc; p.iters = out_iters; p.flags = out_flags;
    p.tail = tail_extrapolation ? 1 : 0;
    p.mixed = (mixed_precision && twiddle_c64) ? 1 : 0;
    p.tw32 = reinterpret_cast<const cx<float>*>(twiddle_c64);
    p.exec_counters = reinterpret_cast<unsigned long long*>(out_exec_counters);
    if (hermitian_half && sc_granger_herm_supported(nfft)) return sc_granger_herm_launch(p, stream);
    return w2_launch<1>(p, B * n_pairs, workspace, workspace_bytes, reinterpret_cast<cudaStream_t>(stream));
}
