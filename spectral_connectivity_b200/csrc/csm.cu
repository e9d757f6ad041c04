// Cross-spectral matrix + expectation, and the pairwise-measure epilogues (SIMT fp32 path).
//
// Replaces connectivity.py:447-526 (_cross_spectral_matrix + _expectation_cross_spectral_matrix):
// the reference materialises the un-averaged (W,T,K,F,S,S) tensor with a k=1 batched matmul
// (:1799-1822) and then reduces it with xp.mean (:67-75).  Here each CTA owns one 64x64 tile of
// one (batch, frequency) matrix and contracts over the R observations directly from the planar
// coefficient layout [B][F][2][R][S]; only upper-triangular tiles are computed and mirrored.
// The per-observation non-linearities of PLV (:899-903) and the PLI family (:970-980,
// :1010-1028, :1090-1127) are applied in registers before accumulation.
//
// The tensor-core (tcgen05) variant of SC_CSM_CROSS lives in csm_tc.cu; this file is the
// path for small S and for the non-linear modes, which are not GEMMs.
#include "sc_common.cuh"

namespace {

constexpr int TILE = 64;
constexpr int RC = 16;
constexpr int kThreads = 256;

template <int MODE>
__global__ void __launch_bounds__(kThreads) csm_kernel(const float* __restrict__ xp, long long BF, long long R,
                                                       long long S, float scale, float* __restrict__ out,
                                                       int ntile) {
    __shared__ __align__(16) float As[2][RC][TILE];
    __shared__ __align__(16) float Bs[2][RC][TILE];
    const long long npair = (long long)ntile * (ntile + 1) / 2;
    const long long bf = blockIdx.x / npair;
    long long pidx = blockIdx.x % npair;
    // decode upper-triangular tile pair (ti <= tj)
    int ti = 0;
    while (pidx >= ntile - ti) {
        pidx -= ntile - ti;
        ++ti;
    }
    const int tj = ti + (int)pidx;
    const long long i0 = (long long)ti * TILE, j0 = (long long)tj * TILE;
    const long long plane = R * S;
    const float* base = xp + bf * 2 * plane;

    const int tx = threadIdx.x % 16, ty = threadIdx.x / 16;
    constexpr int NACC = MODE == SC_CSM_PLI ? 4 : 2;
    float acc[NACC][4][4];
#pragma unroll
    for (int a = 0; a < NACC; ++a)
#pragma unroll
        for (int u = 0; u < 4; ++u)
#pragma unroll
            for (int v = 0; v < 4; ++v) acc[a][u][v] = 0.f;

    const int lc = threadIdx.x % TILE;  // column within tile for loads
    const int lr = threadIdx.x / TILE;  // 0..3
    for (long long r0 = 0; r0 < R; r0 += RC) {
#pragma unroll
        for (int q = 0; q < RC / 4; ++q) {
            const int rr = lr + 4 * q;
            const long long r = r0 + rr;
            const bool rok = r < R;
#pragma unroll
            for (int c = 0; c < 2; ++c) {
                const float* src = base + c * plane + r * S;
                As[c][rr][lc] = (rok && i0 + lc < S) ? __ldg(src + i0 + lc) : 0.f;
                Bs[c][rr][lc] = (rok && j0 + lc < S) ? __ldg(src + j0 + lc) : 0.f;
            }
        }
        __syncthreads();
        const int rcv = (int)((R - r0) < RC ? (R - r0) : RC);
        for (int rr = 0; rr < rcv; ++rr) {
            const float4 are = *reinterpret_cast<const float4*>(&As[0][rr][ty * 4]);
            const float4 aim = *reinterpret_cast<const float4*>(&As[1][rr][ty * 4]);
            const float4 bre = *reinterpret_cast<const float4*>(&Bs[0][rr][tx * 4]);
            const float4 bim = *reinterpret_cast<const float4*>(&Bs[1][rr][tx * 4]);
            const float ar[4] = {are.x, are.y, are.z, are.w}, ai[4] = {aim.x, aim.y, aim.z, aim.w};
            const float br[4] = {bre.x, bre.y, bre.z, bre.w}, bi[4] = {bim.x, bim.y, bim.z, bim.w};
#pragma unroll
            for (int u = 0; u < 4; ++u)
#pragma unroll
                for (int v = 0; v < 4; ++v) {
                    if (MODE == SC_CSM_CROSS) {
                        acc[0][u][v] = fmaf(ar[u], br[v], fmaf(ai[u], bi[v], acc[0][u][v]));
                        acc[1][u][v] = fmaf(ai[u], br[v], fmaf(-ar[u], bi[v], acc[1][u][v]));
                    } else if (MODE == SC_CSM_PLV) {
                        const float pr = fmaf(ar[u], br[v], ai[u] * bi[v]);
                        const float pi = fmaf(ai[u], br[v], -ar[u] * bi[v]);
                        const float inv = 1.0f / sqrtf(fmaf(pr, pr, pi * pi));
                        acc[0][u][v] += pr * inv;  // 0/0 -> NaN like numpy x/abs(x)
                        acc[1][u][v] += pi * inv;
                    } else {
                        float im = fmaf(ai[u], br[v], -ar[u] * bi[v]);
                        if (i0 + ty * 4 + u == j0 + tx * 4 + v) im = 0.f;
                        acc[0][u][v] += (im > 0.f) ? 1.f : ((im < 0.f) ? -1.f : im);  // sign, NaN propagates
                        acc[1][u][v] += fabsf(im);
                        acc[2][u][v] = fmaf(im, im, acc[2][u][v]);
                        acc[3][u][v] += im;
                    }
                }
        }
        __syncthreads();
    }

    const long long mat = bf * S * S;
    const long long pstride = BF * S * S;
#pragma unroll
    for (int u = 0; u < 4; ++u) {
        const long long i = i0 + ty * 4 + u;
        if (i >= S) continue;
#pragma unroll
        for (int v = 0; v < 4; ++v) {
            const long long j = j0 + tx * 4 + v;
            if (j >= S) continue;
            if (MODE == SC_CSM_PLI) {
                const float sg[4] = {-1.f, 1.f, 1.f, -1.f};  // mirror signs: sign(Im), |Im|, Im^2, Im
#pragma unroll
                for (int a = 0; a < 4; ++a) {
                    const float val = acc[a][u][v] * scale;
                    out[a * pstride + mat + i * S + j] = val;
                    if (ti != tj) out[a * pstride + mat + j * S + i] = sg[a] * val;
                }
            } else {
                float2* o = reinterpret_cast<float2*>(out);
                const float2 val = make_float2(acc[0][u][v] * scale, acc[1][u][v] * scale);
                o[mat + i * S + j] = val;
                if (ti != tj) o[mat + j * S + i] = make_float2(val.x, -val.y);
            }
        }
    }
}

__global__ void power_kernel(const float* __restrict__ xp, long long BF, long long R, long long S, float scale,
                             float* __restrict__ out) {
    const long long total = BF * S;
    const long long plane = R * S;
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * blockDim.x) {
        const long long bf = idx / S, s = idx % S;
        const float* re = xp + bf * 2 * plane + s;
        const float* im = re + plane;
        float a0 = 0.f, a1 = 0.f;
        long long r = 0;
        for (; r + 1 < R; r += 2) {
            const float x0 = __ldg(re + r * S), y0 = __ldg(im + r * S);
            const float x1 = __ldg(re + (r + 1) * S), y1 = __ldg(im + (r + 1) * S);
            a0 = fmaf(x0, x0, fmaf(y0, y0, a0));
            a1 = fmaf(x1, x1, fmaf(y1, y1, a1));
        }
        if (r < R) {
            const float x0 = __ldg(re + r * S), y0 = __ldg(im + r * S);
            a0 = fmaf(x0, x0, fmaf(y0, y0, a0));
        }
        out[idx] = (a0 + a1) * scale;
    }
}

constexpr double kEps64 = 2.220446049250313e-16;

__global__ void epilogue_kernel(int measure, const float* __restrict__ in0, const float* __restrict__ in1,
                                long long BF, long long S, double nobs, float* __restrict__ out) {
    const long long total = BF * S * S;
    const float qnan = __int_as_float(0x7fc00000);
    // one matrix row (bf, i) per block iteration, threads over j: no per-element 64-bit divisions
    for (long long row = blockIdx.x; row < BF * S; row += gridDim.x) {
      const long long bf = row / S;
      const long long i = row - bf * S;
      for (long long j = threadIdx.x; j < S; j += blockDim.x) {
        const long long idx = row * S + j;
        switch (measure) {
            case SC_M_COHERENCY:
            case SC_M_COHERENCE_MAG:
            case SC_M_COHERENCE_PHASE:
            case SC_M_IMAG_COHERENCE: {
                // fp32 throughout: inputs are fp32 and the normalisation is well conditioned (1e-7 relative);
                // sqrt(p_i) * sqrt(p_j) cannot under/overflow where p_i * p_j could
                const float2 c = reinterpret_cast<const float2*>(in0)[idx];
                float norm = sqrtf(in1[bf * S + i]) * sqrtf(in1[bf * S + j]);
                norm = norm < (float)kEps64 ? (float)kEps64 : norm;  // connectivity.py:649-652; NaN propagates like np.maximum
                const float inv = 1.0f / norm;
                const float re = c.x * inv, im = c.y * inv;
                if (measure == SC_M_COHERENCY) {
                    reinterpret_cast<float2*>(out)[idx] = i == j ? make_float2(qnan, qnan) : make_float2(re, im);
                } else if (measure == SC_M_COHERENCE_MAG) {
                    float m = fmaf(re, re, im * im);
                    m = m < 0.f ? 0.f : (m > 1.f ? 1.f : m);  // np.clip semantics: NaN stays NaN
                    out[idx] = i == j ? qnan : m;
                } else if (measure == SC_M_COHERENCE_PHASE) {
                    out[idx] = i == j ? qnan : atan2f(im, re);
                } else {
                    const float a = fabsf(im);
                    out[idx] = a > 1.f ? 1.f : a;
                }
                break;
            }
            case SC_M_PLV: {
                const float2 c = reinterpret_cast<const float2*>(in0)[idx];
                out[idx] = (float)sqrt((double)c.x * c.x + (double)c.y * c.y);
                break;
            }
            case SC_M_PPC: {
                const float2 c = reinterpret_cast<const float2*>(in0)[idx];
                const double sx = c.x * nobs, sy = c.y * nobs;
                out[idx] = (float)((sx * sx + sy * sy - nobs) / (nobs * nobs - nobs));
                break;
            }
            case SC_M_PLI: out[idx] = in0[idx]; break;
            case SC_M_DPLI2: {
                const double p = in0[idx];
                out[idx] = (float)((nobs * p * p - 1.0) / (nobs - 1.0));
                break;
            }
            case SC_M_WPLI: {
                double w = in0[total + idx];
                if (w < kEps64) w = 1.0;  // connectivity.py:1027
                out[idx] = (float)((double)in0[3 * total + idx] / w);
                break;
            }
            case SC_M_DWPLI2: {
                const double s_abs = (double)in0[total + idx] * nobs;
                const double s_sq = (double)in0[2 * total + idx] * nobs;
                const double s_im = (double)in0[3 * total + idx] * nobs;
                const double w = s_abs * s_abs - s_sq;
                out[idx] = w == 0.0 ? qnan : (float)((s_im * s_im - s_sq) / w);  // connectivity.py:1125-1127
                break;
            }
            default: break;
        }
      }
    }
}

// Coherence family (coherency / |coherency|^2 / phase / imaginary coherence) for S % 4 == 0: HBM-bound, 8 B in +
// 4 (or 8) B out per pair-frequency.  One thread owns 4 consecutive columns of one matrix row (two 16-byte
// CSM loads, one 16-byte store) and ROWS rows are in flight per thread to cover the DRAM latency.
template <int ROWS>
__global__ void __launch_bounds__(256) coherence_epilogue_vec_kernel(int measure, const float4* __restrict__ csm,
                                                                      const float* __restrict__ power, long long n_rows,
                                                                      int S, float* __restrict__ out) {
    const float qnan = __int_as_float(0x7fc00000);
    const int q_per_row = S >> 2;                       // 4-column groups per row
    const int rows_per_cta = 256 / q_per_row > 0 ? 256 / q_per_row : 1;
    const int active = rows_per_cta * q_per_row;        // threads with work (all 256 when S | 1024)
    if ((int)threadIdx.x >= active && q_per_row <= 256) return;
    const int rl = threadIdx.x / q_per_row;             // row within the CTA's group
    for (long long base = (long long)blockIdx.x * rows_per_cta * ROWS; base < n_rows;
         base += (long long)gridDim.x * rows_per_cta * ROWS) {
        for (int q0 = threadIdx.x % q_per_row; q0 < q_per_row; q0 += 256) {  // S > 1024: several groups per thread
            float4 a[ROWS], b[ROWS];
            long long row[ROWS];
#pragma unroll
            for (int u = 0; u < ROWS; ++u) {
                row[u] = base + (long long)u * rows_per_cta + rl;
                if (row[u] < n_rows) {
                    const float4* src = csm + (row[u] * S + 4 * q0) / 2;
                    a[u] = __ldcs(src);
                    b[u] = __ldcs(src + 1);
                }
            }
#pragma unroll
            for (int u = 0; u < ROWS; ++u) {
                if (row[u] >= n_rows) continue;
                const long long bf = row[u] / S;
                const int i = (int)(row[u] - bf * S);
                const float pi = sqrtf(power[bf * S + i]);
                const float4 pj = *reinterpret_cast<const float4*>(power + bf * S + 4 * q0);
                const float re[4] = {a[u].x, a[u].z, b[u].x, b[u].z}, im[4] = {a[u].y, a[u].w, b[u].y, b[u].w};
                const float pjs[4] = {pj.x, pj.y, pj.z, pj.w};
                float r4[4], i4[4];
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    float norm = pi * sqrtf(pjs[e]);
                    norm = norm < (float)kEps64 ? (float)kEps64 : norm;  // connectivity.py:649-652
                    const float inv = 1.0f / norm;
                    r4[e] = re[e] * inv;
                    i4[e] = im[e] * inv;
                }
                const int j0 = 4 * q0;
                if (measure == SC_M_COHERENCY) {
                    float4 o0 = make_float4(r4[0], i4[0], r4[1], i4[1]), o1 = make_float4(r4[2], i4[2], r4[3], i4[3]);
                    if (i == j0) o0.x = o0.y = qnan;
                    if (i == j0 + 1) o0.z = o0.w = qnan;
                    if (i == j0 + 2) o1.x = o1.y = qnan;
                    if (i == j0 + 3) o1.z = o1.w = qnan;
                    float4* dst = reinterpret_cast<float4*>(out) + (row[u] * S + j0) / 2;
                    __stcs(dst, o0);
                    __stcs(dst + 1, o1);
                } else {
                    float o[4];
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        if (measure == SC_M_COHERENCE_MAG) {
                            float m = fmaf(r4[e], r4[e], i4[e] * i4[e]);
                            m = m < 0.f ? 0.f : (m > 1.f ? 1.f : m);  // np.clip semantics: NaN stays NaN
                            o[e] = (i == j0 + e) ? qnan : m;
                        } else if (measure == SC_M_COHERENCE_PHASE) {
                            o[e] = (i == j0 + e) ? qnan : atan2f(i4[e], r4[e]);
                        } else {
                            const float m = fabsf(i4[e]);
                            o[e] = m > 1.f ? 1.f : m;
                        }
                    }
                    __stcs(reinterpret_cast<float4*>(out + row[u] * S + j0), make_float4(o[0], o[1], o[2], o[3]));
                }
            }
        }
    }
}

// phase_slope_index (connectivity.py:1587-1650): Im sum_{f1 < f2} conj(c[f1]) c[f2] over the selected bins =
// sum_{f2} Im(conj(prefix(f2)) c[f2]) -- one pass with a running prefix instead of the reference's n(n-1)/2
// products.  One thread per (b, i, j), coalesced along j; fp64 accumulators.
__global__ void psi_kernel(const float2* __restrict__ coh, long long B, long long F, long long SS,
                           const int* __restrict__ fidx, int nsel, float* __restrict__ out) {
    for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < B * SS;
         e += (long long)gridDim.x * blockDim.x) {
        const long long b = e / SS, ij = e - b * SS;
        double px = 0.0, py = 0.0, acc = 0.0;
        for (int q = 0; q < nsel; ++q) {
            const float2 c = coh[(b * F + fidx[q]) * SS + ij];
            acc += px * c.y - py * c.x;
            px += c.x;
            py += c.y;
        }
        out[e] = (float)acc;
    }
}

unsigned grid_for(long long total, int threads) {
    long long blocks = (total + threads - 1) / threads;
    const long long cap = (long long)sc_num_sms() * 32;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    return (unsigned)blocks;
}

}  // namespace

extern "C" int sc_power(const float* xp, int64_t B, int64_t F, int64_t R, int64_t S, float scale, float* out,
                        void* stream) {
    SC_CHECK_ARG(xp && out, "sc_power: null pointer");
    SC_CHECK_ARG(B > 0 && F > 0 && R > 0 && S > 0, "sc_power: non-positive size");
    power_kernel<<<grid_for(B * F * S, 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(xp, B * F, R, S, scale,
                                                                                              out);
    SC_LAUNCH_OK();
    return SC_OK;
}

int sc_csm_tc_supported(int64_t R, int64_t S);
int sc_csm_tc_launch(const float* xp, int64_t B, int64_t F, int64_t R, int64_t S, float scale, void* out,
                     cudaStream_t st);

extern "C" int sc_csm_simt(const float* xp, int64_t B, int64_t F, int64_t R, int64_t S, float scale, int mode,
                           void* out, void* stream) {
    SC_CHECK_ARG(xp && out, "sc_csm: null pointer");
    SC_CHECK_ARG(B > 0 && F > 0 && R > 0 && S > 0, "sc_csm: non-positive size");
    SC_CHECK_ARG(mode >= 0 && mode <= 2, "sc_csm: unknown mode %d", mode);
    const int ntile = (int)((S + TILE - 1) / TILE);
    const long long npair = (long long)ntile * (ntile + 1) / 2;
    const long long grid = npair * B * F;
    SC_CHECK_ARG(grid < (1LL << 31), "sc_csm: grid too large (%lld CTAs); split the batch", grid);
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    float* o = reinterpret_cast<float*>(out);
    if (mode == SC_CSM_CROSS) csm_kernel<SC_CSM_CROSS><<<(unsigned)grid, kThreads, 0, st>>>(xp, B * F, R, S, scale, o, ntile);
    else if (mode == SC_CSM_PLV) csm_kernel<SC_CSM_PLV><<<(unsigned)grid, kThreads, 0, st>>>(xp, B * F, R, S, scale, o, ntile);
    else csm_kernel<SC_CSM_PLI><<<(unsigned)grid, kThreads, 0, st>>>(xp, B * F, R, S, scale, o, ntile);
    SC_LAUNCH_OK();
    return SC_OK;
}

extern "C" int sc_csm(const float* xp, int64_t B, int64_t F, int64_t R, int64_t S, float scale, int mode, void* out,
                      void* stream) {
    if (mode == SC_CSM_CROSS && xp && out && sc_csm_tc_supported(R, S))
        return sc_csm_tc_launch(xp, B, F, R, S, scale, out, reinterpret_cast<cudaStream_t>(stream));
    return sc_csm_simt(xp, B, F, R, S, scale, mode, out, stream);
}

extern "C" int sc_pairwise_epilogue(int measure, const void* in0, const float* in1, int64_t B, int64_t F, int64_t S,
                                    double n_observations, void* out, void* stream) {
    SC_CHECK_ARG(in0 && out, "sc_pairwise_epilogue: null pointer");
    SC_CHECK_ARG(measure >= SC_M_COHERENCY && measure <= SC_M_DWPLI2, "sc_pairwise_epilogue: unknown measure %d", measure);
    SC_CHECK_ARG(measure > SC_M_IMAG_COHERENCE || in1, "sc_pairwise_epilogue: measure %d needs the power array", measure);
    SC_CHECK_ARG(B > 0 && F > 0 && S > 0, "sc_pairwise_epilogue: non-positive size");
    const long long rows = B * F * S;
    if (measure <= SC_M_IMAG_COHERENCE && S % 4 == 0 && S <= 1024 && 1024 % S == 0) {
        constexpr int kRows = 4;
        const int rows_per_cta = (int)(1024 / S);
        const long long ctas = (rows + (long long)rows_per_cta * kRows - 1) / ((long long)rows_per_cta * kRows);
        const long long cap_v = (long long)sc_num_sms() * 16;
        coherence_epilogue_vec_kernel<kRows><<<(unsigned)(ctas < cap_v ? ctas : cap_v), 256, 0,
                                               reinterpret_cast<cudaStream_t>(stream)>>>(
            measure, reinterpret_cast<const float4*>(in0), in1, rows, (int)S, reinterpret_cast<float*>(out));
        SC_LAUNCH_OK();
        return SC_OK;
    }
    const int threads = S >= 256 ? 256 : (S >= 128 ? 128 : (S >= 64 ? 64 : 32));
    const long long cap = (long long)sc_num_sms() * 64;
    epilogue_kernel<<<(unsigned)(rows < cap ? rows : cap), threads, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
        measure, reinterpret_cast<const float*>(in0), in1, B * F, S, n_observations, reinterpret_cast<float*>(out));
    SC_LAUNCH_OK();
    return SC_OK;
}

extern "C" int sc_phase_slope_index(const void* coherency_c64, int64_t B, int64_t F, int64_t S, const int* freq_index,
                                    int n_selected, float* out, void* stream) {
    SC_CHECK_ARG(coherency_c64 && freq_index && out, "sc_phase_slope_index: null pointer");
    SC_CHECK_ARG(B > 0 && F > 0 && S > 0 && n_selected >= 0, "sc_phase_slope_index: bad size");
    psi_kernel<<<grid_for(B * S * S, 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
        reinterpret_cast<const float2*>(coherency_c64), B, F, S * S, freq_index, n_selected, out);
    SC_LAUNCH_OK();
    return SC_OK;
}

// power = real diagonal of the expected cross-spectral matrix (connectivity.py:441-445: E[|X_i|^2] = E[X_i conj X_i])
__global__ void power_from_csm_kernel(const float2* __restrict__ csm, long long BF, long long S, float* __restrict__ out) {
    for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < BF * S;
         e += (long long)gridDim.x * blockDim.x) {
        const long long bf = e / S, i = e - bf * S;
        out[e] = csm[(bf * S + i) * S + i].x;
    }
}

extern "C" int sc_power_from_csm(const void* csm_c64, int64_t BF, int64_t S, float* out, void* stream) {
    SC_CHECK_ARG(csm_c64 && out && BF > 0 && S > 0, "sc_power_from_csm: bad argument");
    power_from_csm_kernel<<<grid_for(BF * S, 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
        reinterpret_cast<const float2*>(csm_c64), BF, S, out);
    SC_LAUNCH_OK();
    return SC_OK;
}

// Upper triangle (diagonal included) of a SYMMETRIC real measure, row-major packed: element (i, j >= i) of matrix bf goes
// to out[bf * S (S + 1) / 2 + i S - i (i - 1) / 2 + (j - i)].  Halves the device -> host bytes of coherence-type results
// (the end-to-end step of the headline workload is bound by the host link, not by the kernels).  One CTA per matrix
// walks the rows; thread t copies column i + t of row i, so reads and writes are both contiguous.
__global__ void pack_upper_kernel(const float* __restrict__ in, long long BF, int S, float* __restrict__ out) {
    const long long tri = (long long)S * (S + 1) / 2;
    for (long long bf = blockIdx.x; bf < BF; bf += gridDim.x) {
        const float* m = in + bf * S * S;
        float* o = out + bf * tri;
        for (int i = 0; i < S; ++i) {
            const long long off = (long long)i * S - (long long)i * (i - 1) / 2 - i;  // + j
            for (int j = i + threadIdx.x; j < S; j += blockDim.x) o[off + j] = __ldcs(&m[(long long)i * S + j]);
        }
    }
}

extern "C" int sc_pack_upper(const float* in, int64_t BF, int64_t S, float* out, void* stream) {
    SC_CHECK_ARG(in && out && BF > 0 && S > 0 && S < (1 << 15), "sc_pack_upper: bad argument");
    const long long grid = BF < 148LL * 16 ? BF : 148LL * 16;
    pack_upper_kernel<<<(unsigned)grid, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(in, BF, (int)S, out);
    SC_LAUNCH_OK();
    return SC_OK;
}
