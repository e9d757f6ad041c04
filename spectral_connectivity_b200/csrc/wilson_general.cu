// Wilson spectral factorisation for S x S cross-spectral matrices (2 <= S <= 32) and the MVAR
// quantities derived from the factor.
//
// Replaces minimum_phase_decomposition.py:227-322 for general S (the pairwise 2x2 case has its own
// fused kernels in granger_herm.cu / wilson.cu) and connectivity.py:567-588, 1679-1748 (transfer function,
// noise covariance, MVAR Fourier coefficients) plus the normalisations of :1237-1426 (DTF, directed
// coherence, PDC, gPDC, dDTF).
//
// The reference runs, per iteration, two batched LAPACK solves, an ifft/fft pair along the frequency axis of
// every matrix entry and a batched matmul, then synchronises with the host to test convergence.  Here one
// iteration is three stream-ordered kernels over a batch of windows, with per-window convergence state kept
// on the device (no host synchronisation; converged windows are skipped):
//   wg_linpred   one warp per (window, frequency): G^-1 by Gauss-Jordan with partial pivoting in shared
//                memory, B = G^-1 S G^-H + I
//   wg_plus      one CTA per (window, matrix entry): inverse FFT along frequency, causal projection
//                (lag 0 halved, strictly lower triangle zeroed, lags >= (N+1)/2 dropped), forward FFT.
//                For real time series (hermitian_half) only entries i <= j are transformed -- c_ji[k] =
//                c_ij[-k] -- and the two causal sequences p_ij, p_ji share one forward FFT
//   wg_update    one warp per (window, frequency): G <- G P, max |dG| per window via atomicMax
//   wg_check     per window: iteration count, convergence flag (frozen at the first iterate < tol, :310-315)
#include "wilson_common.cuh"

namespace {

using namespace scw;

constexpr int kMaxS = 32;

// ---- warp-level helpers ---------------------------------------------------------------------
// Gauss-Jordan with partial pivoting on aug = [A | I] (S x 2S, row-major) -> [I | A^-1]; fac: S scratch.
__device__ bool warp_gj_inverse(cd* aug, cd* fac, int S, int lane) {
    const int ld = 2 * S;
    bool ok = true;
    for (int c = 0; c < S; ++c) {
        double bestv = -1.0;
        int best = c;
        for (int r = c + lane; r < S; r += 32) {
            const cd v = aug[r * ld + c];
            const double m = v.x * v.x + v.y * v.y;
            if (m > bestv) {
                bestv = m;
                best = r;
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const double ov = __shfl_xor_sync(0xffffffffu, bestv, o);
            const int ob = __shfl_xor_sync(0xffffffffu, best, o);
            if (ov > bestv || (ov == bestv && ob < best)) {
                bestv = ov;
                best = ob;
            }
        }
        if (!(bestv > 0.0)) ok = false;
        if (best != c)
            for (int col = lane; col < ld; col += 32) {
                const cd t = aug[c * ld + col];
                aug[c * ld + col] = aug[best * ld + col];
                aug[best * ld + col] = t;
            }
        __syncwarp();
        const cd pinv = cdiv1(aug[c * ld + c]);
        for (int r = lane; r < S; r += 32) fac[r] = aug[r * ld + c];
        __syncwarp();
        for (int col = lane; col < ld; col += 32) aug[c * ld + col] = cmul(aug[c * ld + col], pinv);
        __syncwarp();
        for (int idx = lane; idx < S * ld; idx += 32) {
            const int r = idx / ld, col = idx - r * ld;
            if (r != c) aug[idx] = csub(aug[idx], cmul(fac[r], aug[c * ld + col]));
        }
        __syncwarp();
    }
    return ok;
}

struct WgParams {
    const cd* csm;  // [B][F][S][S]
    cd* g;          // [B][F][S][S]
    cd* bp;         // [B][F][S][S] linear predictor B, overwritten by the causal factor P
    double* err;    // [B]
    int* state;     // [B] 0 active, 1 converged, 2 not SPD
    int* iters;     // [B]
    long long B;
    int F, nfft, herm, S;
    double tol;
    const cd* tw;
    ScFftPlan plan;
};

__device__ __forceinline__ double bin_weight(int f, int nfft, int herm) {
    if (!herm) return 1.0;
    return (f == 0 || 2 * f == nfft) ? 1.0 : 2.0;  // bins f and nfft-f of the full circle
}

// lag-0 covariance -> Cholesky -> G0 = L^T for every frequency (mpd.py:48-93, 290-295)
__global__ void wg_init_kernel(const WgParams p) {
    extern __shared__ double sm_a[];  // S*S
    const long long w = blockIdx.x;
    const int S = p.S, SS = S * S;
    const cd* src = p.csm + (size_t)w * p.F * SS;
    for (int e = threadIdx.x; e < SS; e += blockDim.x) {
        double acc = 0.0;
        for (int f = 0; f < p.F; ++f) acc += bin_weight(f, p.nfft, p.herm) * src[(size_t)f * SS + e].x;
        sm_a[e] = acc / p.nfft;
    }
    __syncthreads();
    __shared__ int bad;
    if (threadIdx.x == 0) {
        bad = 0;
        // in-place lower Cholesky of the (symmetric part read from the lower triangle) real matrix
        for (int j = 0; j < S && !bad; ++j) {
            double d = sm_a[j * S + j];
            for (int k = 0; k < j; ++k) d -= sm_a[j * S + k] * sm_a[j * S + k];
            if (!(d > 0.0) || !isfinite(d)) {
                bad = 1;
                break;
            }
            const double l = sqrt(d);
            sm_a[j * S + j] = l;
            for (int i = j + 1; i < S; ++i) {
                double v = sm_a[i * S + j];
                for (int k = 0; k < j; ++k) v -= sm_a[i * S + k] * sm_a[j * S + k];
                sm_a[i * S + j] = v / l;
            }
        }
        p.state[w] = bad ? 2 : 0;
        p.iters[w] = 0;
        p.err[w] = 0.0;
    }
    __syncthreads();
    cd* g = p.g + (size_t)w * p.F * SS;
    const double qnan = __longlong_as_double(0x7ff8000000000000LL);
    for (size_t idx = threadIdx.x; idx < (size_t)p.F * SS; idx += blockDim.x) {
        const int e = (int)(idx % SS);
        const int i = e / S, j = e % S;
        double v = (j >= i) ? sm_a[j * S + i] : 0.0;  // (L^T)[i][j] = L[j][i]
        if (bad) v = qnan;
        g[idx] = cmake<double>(v, bad ? qnan : 0.0);
    }
}

// B = G^-1 S G^-H + I per (window, frequency)  (mpd.py:218-224)
__global__ void wg_linpred_kernel(const WgParams p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int S = p.S, SS = S * S;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int wpb = blockDim.x >> 5;
    const size_t per_warp = (size_t)(2 * SS + S + 2 * SS);
    cd* aug = reinterpret_cast<cd*>(smem_raw) + (size_t)warp * per_warp;
    cd* fac = aug + 2 * SS;
    cd* sm = fac + S;
    cd* tm = sm + SS;
    const long long item = (long long)blockIdx.x * wpb + warp;
    if (item >= p.B * p.F) return;
    const long long w = item / p.F;
    if (p.state[w] != 0) return;
    const cd* g = p.g + (size_t)item * SS;
    const cd* s = p.csm + (size_t)item * SS;
    for (int e = lane; e < SS; e += 32) {
        const int i = e / S, j = e % S;
        aug[i * 2 * S + j] = g[e];
        aug[i * 2 * S + S + j] = cmake<double>(i == j ? 1.0 : 0.0, 0.0);
        sm[e] = s[e];
    }
    __syncwarp();
    warp_gj_inverse(aug, fac, S, lane);
    // T = Ginv S
    for (int e = lane; e < SS; e += 32) {
        const int i = e / S, j = e % S;
        cd acc = cmake<double>(0.0, 0.0);
        for (int k = 0; k < S; ++k) acc = cadd(acc, cmul(aug[i * 2 * S + S + k], sm[k * S + j]));
        tm[e] = acc;
    }
    __syncwarp();
    // B = T Ginv^H + I
    cd* out = p.bp + (size_t)item * SS;
    for (int e = lane; e < SS; e += 32) {
        const int i = e / S, j = e % S;
        cd acc = cmake<double>(i == j ? 1.0 : 0.0, 0.0);
        for (int k = 0; k < S; ++k) acc = cadd(acc, cmulc(tm[i * S + k], aug[j * 2 * S + S + k]));
        out[e] = acc;
    }
}

// causal projection of one matrix entry along the frequency axis (mpd.py:96-142), in place B -> P
__global__ void __launch_bounds__(kThreads) wg_plus_kernel(const WgParams p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int N = p.nfft, S = p.S, SS = S * S;
    cd* ZA = reinterpret_cast<cd*>(smem_raw);
    cd* ZB = ZA + N;
    cd* tws = ZB + N;
    const long long w = blockIdx.y;
    if (p.state[w] != 0) return;
    int i, j;
    if (p.herm) {  // blockIdx.x enumerates i <= j
        int k = blockIdx.x;
        i = 0;
        while (k >= S - i) {
            k -= S - i;
            ++i;
        }
        j = i + k;
    } else {
        i = blockIdx.x / S;
        j = blockIdx.x % S;
    }
    for (int q = threadIdx.x; q < N; q += kThreads) tws[q] = p.tw[q];
    cd* base = p.bp + (size_t)w * p.F * SS;
    if (p.herm) {
        for (int f = threadIdx.x; f < p.F; f += kThreads) {
            const cd v = base[(size_t)f * SS + i * S + j];
            ZA[f] = v;
            if (f != 0 && 2 * f != N) ZA[N - f] = cconj(v);
        }
    } else {
        for (int f = threadIdx.x; f < N; f += kThreads) ZA[f] = base[(size_t)f * SS + i * S + j];
    }
    __syncthreads();
    cd* c = sc_cta_fft<double, true>(ZA, ZB, 1, N, p.plan, tws, true);
    cd* o = (c == ZA) ? ZB : ZA;
    const double inv_n = 1.0 / N;
    const int kcut = (N + 1) / 2;
    for (int k = threadIdx.x; k < N; k += kThreads) {
        cd y = cmake<double>(0.0, 0.0);
        if (k < kcut) {
            const double wgt = k == 0 ? 0.5 * inv_n : inv_n;
            if (p.herm) {
                // real sequences: c_ij[k] and c_ji[k] = c_ij[-k]; lag 0 of the strictly lower entry is zeroed
                const double cij = c[k].x;
                const double cji = (i == j) ? 0.0 : (k == 0 ? 0.0 : c[N - k].x);
                y = cmake<double>(cij * wgt, cji * wgt);
            } else {
                const bool lower0 = (k == 0 && i > j);
                y = lower0 ? cmake<double>(0.0, 0.0) : cscale(c[k], wgt);
            }
        }
        o[k] = y;
    }
    __syncthreads();
    const cd* Q = sc_cta_fft<double, true>(o, c, 1, N, p.plan, tws, false);
    if (p.herm) {
        for (int f = threadIdx.x; f < p.F; f += kThreads) {
            const cd a = Q[f], m = Q[f == 0 ? 0 : N - f];
            const cd pij = cmake<double>(0.5 * (a.x + m.x), 0.5 * (a.y - m.y));
            base[(size_t)f * SS + i * S + j] = pij;
            if (i != j) base[(size_t)f * SS + j * S + i] = cmake<double>(0.5 * (a.y + m.y), 0.5 * (m.x - a.x));
        }
    } else {
        for (int f = threadIdx.x; f < N; f += kThreads) base[(size_t)f * SS + i * S + j] = Q[f];
    }
}

// G <- G P, max |dG| per window (mpd.py:305-315)
__global__ void wg_update_kernel(const WgParams p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int S = p.S, SS = S * S;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int wpb = blockDim.x >> 5;
    cd* gm = reinterpret_cast<cd*>(smem_raw) + (size_t)warp * 2 * SS;
    cd* pm = gm + SS;
    const long long item = (long long)blockIdx.x * wpb + warp;
    if (item >= p.B * p.F) return;
    const long long w = item / p.F;
    if (p.state[w] != 0) return;
    cd* g = p.g + (size_t)item * SS;
    const cd* pp = p.bp + (size_t)item * SS;
    for (int e = lane; e < SS; e += 32) {
        gm[e] = g[e];
        pm[e] = pp[e];
    }
    __syncwarp();
    double err2 = 0.0;
    for (int e = lane; e < SS; e += 32) {
        const int i = e / S, j = e % S;
        cd acc = cmake<double>(0.0, 0.0);
        for (int k = 0; k < S; ++k) acc = cadd(acc, cmul(gm[i * S + k], pm[k * S + j]));
        const cd d = csub(acc, gm[e]);
        err2 = fmax(err2, d.x * d.x + d.y * d.y);
        g[e] = acc;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) err2 = fmax(err2, __shfl_xor_sync(0xffffffffu, err2, o));
    if (lane == 0) {
        // non-negative doubles order like their bit patterns; a NaN error maps to a huge pattern and never converges
        atomicMax(reinterpret_cast<unsigned long long*>(p.err + w), (unsigned long long)__double_as_longlong(sqrt(err2)));
    }
}

__global__ void wg_check_kernel(const WgParams p) {
    const long long w = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (w >= p.B || p.state[w] != 0) return;
    p.iters[w] += 1;
    if (p.err[w] < p.tol) p.state[w] = 1;
    p.err[w] = 0.0;
}

__global__ void wg_finish_kernel(const int* state, int* flags, long long B) {
    const long long w = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (w >= B) return;
    flags[w] = state[w] == 2 ? SC_FLAG_NOT_SPD : (state[w] == 0 ? SC_FLAG_NOT_CONVERGED : 0);
}

// ---- MVAR quantities ------------------------------------------------------------------------------
// H0 = Re ifft(G)[lag 0] per window (connectivity.py:1705, 1739-1740); grid (entry chunks, windows)
__global__ void mvar_lag0_kernel(const cd* g, long long B, int F, int nfft, int herm, int S, double* h0) {
    const long long w = blockIdx.y;
    const int SS = S * S;
    const cd* src = g + (size_t)w * F * SS;
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < SS; e += gridDim.x * blockDim.x) {
        double acc = 0.0;
        for (int f = 0; f < F; ++f) acc += bin_weight(f, nfft, herm) * src[(size_t)f * SS + e].x;
        h0[(size_t)w * SS + e] = acc / nfft;
    }
}

// per window: Minv = (H0 + lam I)^-1, Sigma = H0 H0^T; per kept frequency: H = G Minv  (:1705-1709, :1739-1748)
__global__ void mvar_transfer_kernel(const cd* g, const double* h0, double lam_host, const double* lam_dev, int F, int nfo,
                                     int S, cd* h_out, double* sigma) {
    const double lam = lam_dev ? lam_dev[0] : lam_host;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int SS = S * S;
    cd* aug = reinterpret_cast<cd*>(smem_raw);  // S x 2S
    cd* fac = aug + 2 * SS;
    const long long w = blockIdx.x;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const double* h = h0 + (size_t)w * SS;
    for (int e = threadIdx.x; e < SS; e += blockDim.x) {
        const int i = e / S, j = e % S;
        aug[i * 2 * S + j] = cmake<double>(h[e] + (i == j ? lam : 0.0), 0.0);
        aug[i * 2 * S + S + j] = cmake<double>(i == j ? 1.0 : 0.0, 0.0);
        if (sigma) {
            double acc = 0.0;
            for (int k = 0; k < S; ++k) acc += h[i * S + k] * h[j * S + k];
            sigma[(size_t)w * SS + e] = acc;
        }
    }
    __syncthreads();
    if (warp == 0) warp_gj_inverse(aug, fac, S, lane);
    __syncthreads();
    const cd* src = g + (size_t)w * F * SS;
    cd* dst = h_out + (size_t)w * nfo * SS;
    for (size_t idx = threadIdx.x; idx < (size_t)nfo * SS; idx += blockDim.x) {
        const int f = (int)(idx / SS), e = (int)(idx % SS);
        const int i = e / S, j = e % S;
        cd acc = cmake<double>(0.0, 0.0);
        for (int k = 0; k < S; ++k) acc = cadd(acc, cscale(src[(size_t)f * SS + i * S + k], aug[k * 2 * S + S + j].x));
        dst[idx] = acc;
    }
}

// A = (H + lam I)^-1 per (window, frequency)  (connectivity.py:580-588)
__global__ void mvar_inverse_kernel(const cd* h, double lam_host, const double* lam_dev, long long BF, int S, cd* a_out) {
    const double lam = lam_dev ? lam_dev[0] : lam_host;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int SS = S * S;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
    cd* aug = reinterpret_cast<cd*>(smem_raw) + (size_t)warp * (2 * SS + S);
    cd* fac = aug + 2 * SS;
    const long long item = (long long)blockIdx.x * wpb + warp;
    if (item >= BF) return;
    const cd* src = h + (size_t)item * SS;
    for (int e = lane; e < SS; e += 32) {
        const int i = e / S, j = e % S;
        cd v = src[e];
        if (i == j) v.x += lam;
        aug[i * 2 * S + j] = v;
        aug[i * 2 * S + S + j] = cmake<double>(i == j ? 1.0 : 0.0, 0.0);
    }
    __syncwarp();
    warp_gj_inverse(aug, fac, S, lane);
    cd* dst = a_out + (size_t)item * SS;
    for (int e = lane; e < SS; e += 32) dst[e] = aug[(e / S) * 2 * S + S + (e % S)];
}

// sum over (frequency, source) of |H_ij|^2 per (window, target i): the dDTF denominator (:1418-1420)
__global__ void mvar_inflow_all_kernel(const cd* h, int F, int S, double* out) {
    const long long w = blockIdx.x;
    const int SS = S * S;
    for (int i = threadIdx.x; i < S; i += blockDim.x) {
        double acc = 0.0;
        for (int f = 0; f < F; ++f)
            for (int j = 0; j < S; ++j) {
                const cd v = h[((size_t)w * F + f) * SS + i * S + j];
                acc += v.x * v.x + v.y * v.y;
            }
        out[(size_t)w * S + i] = acc;
    }
}

// DTF / DC / PDC / gPDC / dDTF normalisations, one CTA per (window, frequency)
__global__ void mvar_measure_kernel(int measure, const cd* h, const cd* a, const double* sigma, const double* inflow_all,
                                    int F, int S, float* out) {
    extern __shared__ double sm_d[];  // |H|^2 [SS], |A|^2 [SS], rows/cols sums
    const int SS = S * S;
    double* h2 = sm_d;
    double* a2 = h2 + SS;
    double* rs = a2 + SS;   // per-row (inflow) sums
    double* cs = rs + S;    // per-column (outflow) sums
    const long long wf = blockIdx.x;
    const long long w = wf / F;
    for (int e = threadIdx.x; e < SS; e += blockDim.x) {
        const cd hv = h ? h[(size_t)wf * SS + e] : cmake<double>(0.0, 0.0);
        const cd av = a ? a[(size_t)wf * SS + e] : cmake<double>(0.0, 0.0);
        h2[e] = hv.x * hv.x + hv.y * hv.y;
        a2[e] = av.x * av.x + av.y * av.y;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < S; i += blockDim.x) {
        const double nv_i = sigma ? sigma[(size_t)w * SS + i * S + i] : 1.0;
        double r = 0.0, c = 0.0;
        for (int k = 0; k < S; ++k) {
            // inflow of target i: sum over sources k (noise variance indexed by the ROW, :1904-1925)
            r += (measure == 1 ? nv_i : 1.0) * h2[i * S + k];
            // outflow of source i: sum over targets k
            const double nv_k = sigma ? sigma[(size_t)w * SS + k * S + k] : 1.0;
            c += a2[k * S + i] / (measure == 3 ? nv_k : 1.0);
        }
        rs[i] = r;
        cs[i] = c;
    }
    __syncthreads();
    for (int e = threadIdx.x; e < SS; e += blockDim.x) {
        const int i = e / S, j = e % S;
        const double nv_i = sigma ? sigma[(size_t)w * SS + i * S + i] : 1.0;
        double v;
        switch (measure) {
            case 0: v = h2[e] / rs[i]; break;                                    // DTF  (:1264-1266)
            case 1: v = sqrt(nv_i) * h2[e] / sqrt(rs[i]); break;                 // DC   (:1290-1296)
            case 2: v = a2[e] / cs[j]; break;                                    // PDC  (:1322-1343)
            case 3: v = a2[e] / nv_i / cs[j]; break;                             // gPDC (:1372-1380)
            default: v = sqrt(h2[e] / inflow_all[(size_t)w * S + i]) * sqrt(a2[e] / cs[j]); break;  // dDTF (:1418-1426)
        }
        out[(size_t)wf * SS + e] = (float)v;
    }
}

#include "wilson_blocked.cuh"

constexpr int kMaxBigS = 1024;

int check_s(int S, const char* what) {
    if (S < 1 || S > kMaxBigS) {
        sc_set_error("%s: S=%d outside [1, %d]", what, S, kMaxBigS);
        return SC_ERR_UNSUPPORTED;
    }
    return SC_OK;
}

inline unsigned grid_for(long long n, int threads) {
    const long long g = (n + threads - 1) / threads;
    const long long cap = (long long)sc_num_sms() * 32;
    return (unsigned)(g < cap ? (g < 1 ? 1 : g) : cap);
}

// Wilson iteration for S > kMaxS (see wilson_blocked.cuh).  Workspace: 3 matrix buffers + inversion scratch.
int wilson_blocked(WgParams p, int max_iterations, cd* out_g, int* out_iters, int* out_flags, unsigned char* ws,
                   cudaStream_t st) {
    const int S = p.S, F = p.F;
    const long long B = p.B, count = B * F;
    const size_t mat = (size_t)count * S * S;
    cd* w0 = reinterpret_cast<cd*>(ws);
    cd* w1 = w0 + mat;
    cd* w2 = w1 + mat;
    unsigned char* tail = reinterpret_cast<unsigned char*>(w2 + mat);
    p.err = reinterpret_cast<double*>(tail);
    p.state = reinterpret_cast<int*>(tail + (size_t)B * 8);
    p.iters = reinterpret_cast<int*>(tail + (size_t)B * 16);
    double* lag0 = reinterpret_cast<double*>(tail + (((size_t)B * 32 + 255) / 256) * 256);
    unsigned char* scratch = reinterpret_cast<unsigned char*>(lag0) + (((size_t)B * S * S * 8 + 255) / 256) * 256;
    size_t plus_smem = (size_t)3 * p.nfft * sizeof(cd);
    const bool tile_plus = zb_plus_smem(p.nfft) <= (size_t)sc_max_smem_optin() - 2048;  // 16 sequences fit
    if (tile_plus) {
        plus_smem = zb_plus_smem(p.nfft);
        if (plus_smem > 48 * 1024)
            SC_CUDA_OK(cudaFuncSetAttribute(wg_plus_tile_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)plus_smem));
    } else if (plus_smem > 48 * 1024) {
        SC_CUDA_OK(cudaFuncSetAttribute(wg_plus_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)plus_smem));
    }
    const int n_t = (S + kPT - 1) / kPT;
    const int n_tiles = p.herm ? n_t * (n_t + 1) / 2 : n_t * n_t;
    // lag-0 covariance -> Cholesky -> G0 = L^T for every frequency
    mvar_lag0_kernel<<<dim3((unsigned)((S * S + 255) / 256), (unsigned)B), 256, 0, st>>>(p.csm, B, F, p.nfft, p.herm, S, lag0);
    zb_cholesky_kernel<<<(unsigned)B, 1024, (size_t)S * sizeof(double), st>>>(lag0, S, p.state, p.iters, p.err);
    cd* gcur = out_g;
    cd* free_a = w0;
    cd* free_b = w1;
    cd* bp = w2;
    zb_init_g_kernel<<<dim3(grid_for((long long)F * S * S, 256), (unsigned)B), 256, 0, st>>>(lag0, p.state, F, S, gcur);
    SC_LAUNCH_OK();
    const long long SS = (long long)S * S;
    const int n_entries = p.herm ? S * (S + 1) / 2 : S * S;
    for (int it = 0; it < max_iterations; ++it) {
        // Ginv: copy G, invert in place, unscramble into free_b
        SC_CUDA_OK(cudaMemcpyAsync(free_a, gcur, mat * sizeof(cd), cudaMemcpyDeviceToDevice, st));
        if (int rc = zb_inverse(free_a, free_b, count, S, p.state, F, nullptr, scratch, st)) return rc;
        ZGemmParams g;
        g.n_inner = F; g.S = S; g.state = p.state; g.passthrough = 0; g.err = nullptr; g.upper_only = 0;
        g.a_so = g.b_so = g.c_so = (long long)F * SS;
        g.a_si = g.b_si = g.c_si = SS;
        // T = Ginv S -> free_a
        g.A = free_b; g.Bm = p.csm; g.C = free_a; g.conj_b = 0; g.add_identity = 0;
        zb_gemm(g, count, st);
        // B = T Ginv^H + I -> bp
        // (B is Hermitian; for real series the projection reads its upper triangle only)
        g.A = free_a; g.Bm = free_b; g.C = bp; g.conj_b = 1; g.add_identity = 1; g.upper_only = p.herm;
        zb_gemm(g, count, st);
        g.upper_only = 0;
        p.bp = bp;
        if (tile_plus) wg_plus_tile_kernel<<<dim3(n_tiles, (unsigned)B), kThreads, plus_smem, st>>>(p);
        else wg_plus_kernel<<<dim3(n_entries, (unsigned)B), kThreads, plus_smem, st>>>(p);
        // G <- G P -> free_a, max |dG| per window; converged windows are copied through
        g.A = gcur; g.Bm = bp; g.C = free_a; g.conj_b = 0; g.add_identity = 0; g.passthrough = 1; g.err = p.err;
        zb_gemm(g, count, st);
        wg_check_kernel<<<(unsigned)((B + 127) / 128), 128, 0, st>>>(p);
        SC_LAUNCH_OK();
        cd* t = gcur;
        gcur = free_a;
        free_a = t;
    }
    if (gcur != out_g) SC_CUDA_OK(cudaMemcpyAsync(out_g, gcur, mat * sizeof(cd), cudaMemcpyDeviceToDevice, st));
    if (out_flags) wg_finish_kernel<<<(unsigned)((B + 127) / 128), 128, 0, st>>>(p.state, out_flags, B);
    if (out_iters) SC_CUDA_OK(cudaMemcpyAsync(out_iters, p.iters, (size_t)B * sizeof(int), cudaMemcpyDeviceToDevice, st));
    SC_LAUNCH_OK();
    return SC_OK;
}

}  // namespace

extern "C" int64_t sc_wilson_general_workspace_bytes(int64_t B, int F, int S) {
    if (B < 1 || F < 1 || S < 1) return 0;
    const int64_t mat = (int64_t)B * F * S * S * (int64_t)sizeof(cd);
    if (S <= kMaxS) return mat + (int64_t)B * 32;
    return 3 * mat + (((int64_t)B * 32 + 255) / 256) * 256 + (((int64_t)B * S * S * 8 + 255) / 256) * 256 +
           zb_inverse_scratch_bytes(B * F, S);
}

extern "C" int64_t sc_mvar_workspace_bytes(int64_t B, int F, int S) {
    if (B < 1 || F < 1 || S <= kMaxS) return 0;
    return (int64_t)B * F * S * S * (int64_t)sizeof(cd) + zb_inverse_scratch_bytes(B * F, S) + 256;
}

extern "C" int sc_wilson(const void* csm_c128, int64_t B, int F, int nfft, int hermitian_half, int S, double tolerance,
                         int max_iterations, const void* twiddle_c128, void* out_g_c128, int* out_iters, int* out_flags,
                         void* workspace, int64_t workspace_bytes, void* stream) {
    SC_CHECK_ARG(csm_c128 && twiddle_c128 && out_g_c128 && workspace, "sc_wilson: null pointer");
    SC_CHECK_ARG(B > 0 && nfft > 0 && max_iterations >= 0, "sc_wilson: bad size");
    SC_CHECK_ARG(hermitian_half ? F == nfft / 2 + 1 : F == nfft, "sc_wilson: F=%d inconsistent with nfft=%d", F, nfft);
    if (int rc = check_s(S, "sc_wilson")) return rc;
    const int64_t need = sc_wilson_general_workspace_bytes(B, F, S);
    if (workspace_bytes < need) {
        sc_set_error("sc_wilson: workspace of %lld bytes required, got %lld", (long long)need, (long long)workspace_bytes);
        return SC_ERR_WORKSPACE;
    }
    WgParams p;
    p.csm = reinterpret_cast<const cd*>(csm_c128);
    p.g = reinterpret_cast<cd*>(out_g_c128);
    p.bp = reinterpret_cast<cd*>(workspace);
    unsigned char* tail = reinterpret_cast<unsigned char*>(workspace) + (size_t)B * F * S * S * sizeof(cd);
    p.err = reinterpret_cast<double*>(tail);
    p.state = reinterpret_cast<int*>(tail + (size_t)B * 8);
    p.iters = reinterpret_cast<int*>(tail + (size_t)B * 16);
    p.B = B; p.F = F; p.nfft = nfft; p.herm = hermitian_half ? 1 : 0; p.S = S; p.tol = tolerance;
    p.tw = reinterpret_cast<const cd*>(twiddle_c128);
    if (sc_fft_make_plan(nfft, &p.plan)) {
        sc_set_error("sc_wilson: cannot factorise nfft=%d", nfft);
        return SC_ERR_UNSUPPORTED;
    }
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const int SS = S * S;
    const size_t plus_smem = (size_t)3 * nfft * sizeof(cd);
    if (plus_smem > (size_t)sc_max_smem_optin() - 1024) {
        sc_set_error("sc_wilson: nfft=%d needs %zu bytes of shared memory for the causal projection", nfft, plus_smem);
        return SC_ERR_UNSUPPORTED;
    }
    SC_CHECK_ARG(B <= 65535, "sc_wilson: at most 65535 windows per call");
    if (S > kMaxS)
        return wilson_blocked(p, max_iterations, reinterpret_cast<cd*>(out_g_c128), out_iters, out_flags,
                              reinterpret_cast<unsigned char*>(workspace), st);
    const size_t lp_per_warp = (size_t)(4 * SS + S) * sizeof(cd);
    int lp_wpb = (int)(((size_t)sc_max_smem_optin() - 2048) / lp_per_warp);
    lp_wpb = lp_wpb > 8 ? 8 : (lp_wpb < 1 ? 1 : lp_wpb);
    const size_t up_per_warp = (size_t)2 * SS * sizeof(cd);
    const int up_wpb = 8;
    if (plus_smem > 48 * 1024)
        SC_CUDA_OK(cudaFuncSetAttribute(wg_plus_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)plus_smem));
    if (lp_per_warp * lp_wpb > 48 * 1024)
        SC_CUDA_OK(cudaFuncSetAttribute(wg_linpred_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        (int)(lp_per_warp * lp_wpb)));
    if (up_per_warp * up_wpb > 48 * 1024)
        SC_CUDA_OK(cudaFuncSetAttribute(wg_update_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        (int)(up_per_warp * up_wpb)));
    const long long items = B * F;
    const unsigned lp_grid = (unsigned)((items + lp_wpb - 1) / lp_wpb);
    const unsigned up_grid = (unsigned)((items + up_wpb - 1) / up_wpb);
    const int n_entries = hermitian_half ? S * (S + 1) / 2 : SS;
    wg_init_kernel<<<(unsigned)B, 256, SS * sizeof(double), st>>>(p);
    SC_LAUNCH_OK();
    for (int it = 0; it < max_iterations; ++it) {
        wg_linpred_kernel<<<lp_grid, lp_wpb * 32, lp_per_warp * lp_wpb, st>>>(p);
        wg_plus_kernel<<<dim3(n_entries, (unsigned)B), kThreads, plus_smem, st>>>(p);
        wg_update_kernel<<<up_grid, up_wpb * 32, up_per_warp * up_wpb, st>>>(p);
        wg_check_kernel<<<(unsigned)((B + 127) / 128), 128, 0, st>>>(p);
        SC_LAUNCH_OK();
    }
    if (out_flags) wg_finish_kernel<<<(unsigned)((B + 127) / 128), 128, 0, st>>>(p.state, out_flags, B);
    if (out_iters) SC_CUDA_OK(cudaMemcpyAsync(out_iters, p.iters, (size_t)B * sizeof(int), cudaMemcpyDeviceToDevice, st));
    SC_LAUNCH_OK();
    return SC_OK;
}

extern "C" int sc_mvar_lag0(const void* g_c128, int64_t B, int F, int nfft, int hermitian_half, int S, double* out_h0,
                            void* stream) {
    SC_CHECK_ARG(g_c128 && out_h0 && B > 0 && F > 0, "sc_mvar_lag0: bad argument");
    if (int rc = check_s(S, "sc_mvar_lag0")) return rc;
    SC_CHECK_ARG(B <= 65535, "sc_mvar_lag0: at most 65535 windows per call");
    mvar_lag0_kernel<<<dim3((unsigned)((S * S + 255) / 256), (unsigned)B), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
        reinterpret_cast<const cd*>(g_c128), B, F, nfft, hermitian_half ? 1 : 0, S, out_h0);
    SC_LAUNCH_OK();
    return SC_OK;
}

extern "C" int sc_mvar_transfer(const void* g_c128, const double* h0, double lambda, const double* lambda_device,
                                int64_t B, int F, int n_freq_out, int S, void* out_h_c128, double* out_sigma,
                                void* workspace, int64_t workspace_bytes, void* stream) {
    SC_CHECK_ARG(g_c128 && h0 && out_h_c128 && B > 0 && F > 0 && n_freq_out > 0 && n_freq_out <= F,
                 "sc_mvar_transfer: bad argument");
    if (int rc = check_s(S, "sc_mvar_transfer")) return rc;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    if (S > kMaxS) {
        const int64_t need = sc_mvar_workspace_bytes(B, 2, S);  // (H0 + lambda I) and its inverse: 2 matrices per window
        if (!workspace || workspace_bytes < need) {
            sc_set_error("sc_mvar_transfer: workspace of %lld bytes required, got %lld", (long long)need,
                         (long long)workspace_bytes);
            return SC_ERR_WORKSPACE;
        }
        const long long SS = (long long)S * S;
        cd* work = reinterpret_cast<cd*>(workspace);
        cd* minv = work + (size_t)B * SS;
        unsigned char* scratch = reinterpret_cast<unsigned char*>(minv + (size_t)B * SS);
        zb_shift_copy_kernel<<<grid_for(B * SS, 256), 256, 0, st>>>(nullptr, h0, lambda, lambda_device, B, S, work);
        if (int rc = zb_inverse(work, minv, B, S, nullptr, 1, nullptr, scratch, st)) return rc;
        ZGemmParams g;
        g.A = reinterpret_cast<const cd*>(g_c128); g.Bm = minv; g.C = reinterpret_cast<cd*>(out_h_c128);
        g.a_so = (long long)F * SS; g.a_si = SS; g.b_so = SS; g.b_si = 0; g.c_so = (long long)n_freq_out * SS; g.c_si = SS;
        g.n_inner = n_freq_out; g.S = S; g.conj_b = 0; g.add_identity = 0; g.state = nullptr; g.passthrough = 0;
        g.err = nullptr; g.upper_only = 0;
        zb_gemm(g, B * n_freq_out, st);
        if (out_sigma) zb_sigma_kernel<<<grid_for(B * SS, 256), 256, 0, st>>>(h0, B, S, out_sigma);
        SC_LAUNCH_OK();
        return SC_OK;
    }
    const size_t smem = (size_t)(2 * S * S + S) * sizeof(cd);
    mvar_transfer_kernel<<<(unsigned)B, 256, smem, st>>>(reinterpret_cast<const cd*>(g_c128), h0, lambda, lambda_device, F,
                                                          n_freq_out, S, reinterpret_cast<cd*>(out_h_c128), out_sigma);
    SC_LAUNCH_OK();
    return SC_OK;
}

extern "C" int sc_mvar_inverse(const void* h_c128, double lambda, const double* lambda_device, int64_t BF, int S,
                               void* out_a_c128, void* workspace, int64_t workspace_bytes, void* stream) {
    SC_CHECK_ARG(h_c128 && out_a_c128 && BF > 0, "sc_mvar_inverse: bad argument");
    if (int rc = check_s(S, "sc_mvar_inverse")) return rc;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    if (S > kMaxS) {
        const int64_t need = sc_mvar_workspace_bytes(BF, 1, S);
        if (!workspace || workspace_bytes < need) {
            sc_set_error("sc_mvar_inverse: workspace of %lld bytes required, got %lld", (long long)need,
                         (long long)workspace_bytes);
            return SC_ERR_WORKSPACE;
        }
        cd* work = reinterpret_cast<cd*>(workspace);
        unsigned char* scratch = reinterpret_cast<unsigned char*>(work + (size_t)BF * S * S);
        zb_shift_copy_kernel<<<grid_for(BF * (long long)S * S, 256), 256, 0, st>>>(reinterpret_cast<const cd*>(h_c128), nullptr,
                                                                                  lambda, lambda_device, BF, S, work);
        return zb_inverse(work, reinterpret_cast<cd*>(out_a_c128), BF, S, nullptr, 1, nullptr, scratch, st);
    }
    const size_t per_warp = (size_t)(2 * S * S + S) * sizeof(cd);
    int wpb = (int)(((size_t)sc_max_smem_optin() - 2048) / per_warp);
    wpb = wpb > 8 ? 8 : (wpb < 1 ? 1 : wpb);
    if (per_warp * wpb > 48 * 1024)
        SC_CUDA_OK(cudaFuncSetAttribute(mvar_inverse_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(per_warp * wpb)));
    mvar_inverse_kernel<<<(unsigned)((BF + wpb - 1) / wpb), wpb * 32, per_warp * wpb, st>>>(
        reinterpret_cast<const cd*>(h_c128), lambda, lambda_device, BF, S, reinterpret_cast<cd*>(out_a_c128));
    SC_LAUNCH_OK();
    return SC_OK;
}

extern "C" int sc_mvar_measure(int measure, const void* h_c128, const void* a_c128, const double* sigma, int64_t B, int F,
                               int S, double* scratch, float* out, void* stream) {
    SC_CHECK_ARG(out && B > 0 && F > 0, "sc_mvar_measure: bad argument");
    SC_CHECK_ARG(measure >= 0 && measure <= 4, "sc_mvar_measure: unknown measure %d", measure);
    SC_CHECK_ARG((measure == 0 || measure == 1 || measure == 4) ? h_c128 != nullptr : true, "sc_mvar_measure: needs H");
    SC_CHECK_ARG((measure >= 2) ? a_c128 != nullptr : true, "sc_mvar_measure: needs A");
    SC_CHECK_ARG((measure == 1 || measure == 3) ? sigma != nullptr : true, "sc_mvar_measure: needs the noise covariance");
    SC_CHECK_ARG(measure != 4 || scratch, "sc_mvar_measure: dDTF needs a scratch of B*S doubles");
    if (int rc = check_s(S, "sc_mvar_measure")) return rc;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    if (measure == 4) {
        mvar_inflow_all_kernel<<<(unsigned)B, 64, 0, st>>>(reinterpret_cast<const cd*>(h_c128), F, S, scratch);
        SC_LAUNCH_OK();
    }
    if (S > kMaxS) {
        SC_CHECK_ARG(scratch, "sc_mvar_measure: S > %d needs a scratch of B*S + 2*B*F*S doubles", kMaxS);
        double* rs = scratch + (size_t)B * S;
        double* cs = rs + (size_t)B * F * S;
        const cd* h = reinterpret_cast<const cd*>(h_c128);
        const cd* a = reinterpret_cast<const cd*>(a_c128);
        const bool need_h = measure == 0 || measure == 1 || measure == 4, need_a = measure >= 2;
        zb_mvar_sums_kernel<<<(unsigned)(B * F), 256, 0, st>>>(measure, need_h ? h : nullptr, need_a ? a : nullptr, sigma, F, S,
                                                              rs, cs);
        zb_mvar_measure_kernel<<<grid_for(B * F * (long long)S * S, 256), 256, 0, st>>>(
            measure, need_h ? h : nullptr, need_a ? a : nullptr, sigma, scratch, rs, cs, B * F, F, S, out);
        SC_LAUNCH_OK();
        return SC_OK;
    }
    const size_t smem = (size_t)(2 * S * S + 2 * S) * sizeof(double);
    mvar_measure_kernel<<<(unsigned)(B * F), 128, smem, st>>>(measure, reinterpret_cast<const cd*>(h_c128),
                                                               reinterpret_cast<const cd*>(a_c128), sigma, scratch, F, S, out);
    SC_LAUNCH_OK();
    return SC_OK;
}

// ---- global coherence for S > 64 (svd_measures.cu keeps S <= 64 in shared memory) --------------------------
// Largest eigenvalue / eigenvector of each Hermitian PSD cross-spectral matrix by repeated squaring of the
// trace-normalised matrix (same scheme as svd_measures.cu: tr(P^2) reaches 1 when P has rank one), with the
// squarings done by the tiled c128 GEMM; a per-matrix state stops the work of converged matrices.
namespace {

constexpr int kGcSquarings = 32;

__global__ void __launch_bounds__(256) gc_init_kernel(const float2* csm, int S, cd* P, int* state) {
    __shared__ double red[8];
    const long long b = blockIdx.x;
    const size_t nn = (size_t)S * S;
    const float2* m = csm + b * nn;
    double t = 0.0;
    for (int i = threadIdx.x; i < S; i += 256) t += m[(size_t)i * S + i].x;
    double v[1] = {t};
    block_sum<1>(v, red);
    const double sc = v[0] > 0.0 ? 1.0 / v[0] : 0.0;
    for (size_t e = threadIdx.x; e < nn; e += 256) {  // Hermitian part (real diagonal), trace 1
        const int i = (int)(e / S), j = (int)(e % S);
        const float2 a = m[e], t = m[(size_t)j * S + i];
        P[b * nn + e] = cmake<double>(0.5 * ((double)a.x + t.x) * sc, 0.5 * ((double)a.y - t.y) * sc);
    }
    if (threadIdx.x == 0) state[b] = 0;
}

// Q <- Q / tr(Q); converged when 1 - Re tr(Q) < 1e-13 (tr(Q) = tr(P^2) with tr(P) = 1).  The division uses the
// COMPLEX trace: the input diagonal carries imaginary rounding noise (~1e-10 relative from the fp32 CSM), i.e. a
// component (1 + i b) * projector, and squaring doubles b every step while a real normalisation cannot remove
// it (b^2 then exceeds the convergence threshold and the iteration runs away: measured before this fix).
__global__ void __launch_bounds__(256) gc_norm_kernel(cd* Q, int S, int* state) {
    __shared__ double red[2 * 8];
    const long long b = blockIdx.x;
    if (state[b] != 0) return;
    const size_t nn = (size_t)S * S;
    cd* q = Q + b * nn;
    double tx = 0.0, ty = 0.0;
    for (int i = threadIdx.x; i < S; i += 256) {
        tx += q[(size_t)i * S + i].x;
        ty += q[(size_t)i * S + i].y;
    }
    double v[2] = {tx, ty};
    block_sum<2>(v, red);
    const double n2 = v[0] * v[0] + v[1] * v[1];
    const bool ok = n2 > 0.0 && isfinite(n2);
    const cd inv = ok ? cmake<double>(v[0] / n2, -v[1] / n2) : cmake<double>(0.0, 0.0);
    for (size_t e = threadIdx.x; e < nn; e += 256) q[e] = cmul(q[e], inv);
    if (threadIdx.x == 0 && (!ok || 1.0 - v[0] < 1e-13)) state[b] = 1;
}

// eigenvector = the column of the (nearly rank-one) power with the largest diagonal entry; eigenvalue = its
// Rayleigh quotient with the ORIGINAL matrix
__global__ void __launch_bounds__(256) gc_extract_kernel(const float2* csm, const cd* P, int S, float* value,
                                                          float2* vector) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ double red[2 * 8];
    __shared__ int best_sh;
    cd* vec = reinterpret_cast<cd*>(smem_raw);  // [S]
    const long long b = blockIdx.x;
    const size_t nn = (size_t)S * S;
    const cd* p = P + b * nn;
    const float2* a = csm + b * nn;
    if (threadIdx.x == 0) {
        int best = 0;
        double bv = -1.0;
        for (int i = 0; i < S; ++i) {
            const double d = p[(size_t)i * S + i].x;
            if (d > bv) {
                bv = d;
                best = i;
            }
        }
        best_sh = best;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < S; i += 256) vec[i] = p[(size_t)i * S + best_sh];
    __syncthreads();
    double num = 0.0, den = 0.0;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int i = warp; i < S; i += 8) {  // one warp per row of A: coalesced
        cd acc = cmake<double>(0.0, 0.0);
        for (int k = lane; k < S; k += 32) {
            const float2 av = a[(size_t)i * S + k];
            acc = cadd(acc, cmul(cmake<double>(av.x, av.y), vec[k]));
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            acc.x += __shfl_xor_sync(0xffffffffu, acc.x, o);
            acc.y += __shfl_xor_sync(0xffffffffu, acc.y, o);
        }
        if (lane == 0) {
            num += vec[i].x * acc.x + vec[i].y * acc.y;
            den += vec[i].x * vec[i].x + vec[i].y * vec[i].y;
        }
    }
    double v[2] = {num, den};
    block_sum<2>(v, red);
    const double lam = v[1] > 0.0 ? v[0] / v[1] : 0.0;
    const double inv = v[1] > 0.0 ? 1.0 / sqrt(v[1]) : 0.0;
    if (threadIdx.x == 0) value[b] = (float)lam;
    for (int i = threadIdx.x; i < S; i += 256)
        vector[b * S + i] = make_float2((float)(vec[i].x * inv), (float)(vec[i].y * inv));
}

}  // namespace

int64_t sc_global_coherence_blocked_workspace(int64_t BF, int S) {
    return 2 * BF * (int64_t)S * S * (int64_t)sizeof(cd) + BF * 4 + 256;
}

int sc_global_coherence_blocked(const void* csm_c64, int64_t BF, int S, float* out_value, void* out_vector_c64,
                                void* workspace, int64_t workspace_bytes, void* stream) {
    if (int rc = check_s(S, "sc_global_coherence")) return rc;
    const int64_t need = sc_global_coherence_blocked_workspace(BF, S);
    if (!workspace || workspace_bytes < need) {
        sc_set_error("sc_global_coherence: workspace of %lld bytes required, got %lld", (long long)need,
                     (long long)workspace_bytes);
        return SC_ERR_WORKSPACE;
    }
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const size_t mat = (size_t)BF * S * S;
    cd* P = reinterpret_cast<cd*>(workspace);
    cd* Q = P + mat;
    int* state = reinterpret_cast<int*>(Q + mat);
    const float2* csm = reinterpret_cast<const float2*>(csm_c64);
    gc_init_kernel<<<(unsigned)BF, 256, 0, st>>>(csm, S, P, state);
    ZGemmParams g;
    g.a_so = g.b_so = g.c_so = (long long)S * S;
    g.a_si = g.b_si = g.c_si = 0;
    g.n_inner = 1; g.S = S; g.conj_b = 0; g.add_identity = 0; g.state = state; g.passthrough = 1; g.err = nullptr;
    g.upper_only = 0;
    for (int it = 0; it < kGcSquarings; ++it) {
        g.A = P; g.Bm = P; g.C = Q;
        zb_gemm(g, BF, st);
        gc_norm_kernel<<<(unsigned)BF, 256, 0, st>>>(Q, S, state);
        cd* t = P;
        P = Q;
        Q = t;
    }
    gc_extract_kernel<<<(unsigned)BF, 256, (size_t)S * sizeof(cd), st>>>(csm, P, S, out_value,
                                                                        reinterpret_cast<float2*>(out_vector_c64));
    SC_LAUNCH_OK();
    return SC_OK;
}
