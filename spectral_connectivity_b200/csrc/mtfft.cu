// Fused multitaper FFT: sliding-window gather + detrend + DPSS taper product + real FFT.
// Replaces transforms.py:1147-1171 (Multitaper.fft) = _sliding_window (:1311-1374) +
// detrend (:1798-1915) + _multitaper_fft (:1377-1405), which in the reference materialise a
// (W,T,S,n) window copy, a (W,T,S,n,K) float64 product and a two-sided complex128 FFT.
//
// One CTA owns one (window, trial, tile of TS signals).  The n x TS slab is loaded once
// (coalesced along the signal axis), detrended in shared memory, and for every taper the
// TS real series are packed two-per-complex-FFT, transformed by the CTA-cooperative Stockham
// FFT (fft_device.cuh), unpacked to half (or full) spectra and stored with the signal axis
// fastest, so the CSM stage reads [F][2][R][S] tiles directly.
#include <cuda.h>
#include <stdlib.h>
#include <string.h>

#include <type_traits>

#include "fft_device.cuh"
#include "sc_common.cuh"

namespace {

constexpr int kThreads = 256;
constexpr int kWsCtas = 296;
constexpr int kTmaBoxRows = 256;  // rows (samples) per TMA box: the hardware limit of a box dimension

// ---- TMA (cp.async.bulk.tensor) staging of the n x TS input slab --------------------------------------------
// The slab of one (window, trial, signal tile) is a 3-D box {TS signals, 1 trial, rows} of the (N, T, S) series.
// One elected thread issues ceil(n / 256) bulk tensor copies that land densely ([row][TS]) in shared memory and
// complete on an mbarrier: the load costs no registers and no issue slots of the other 255 threads, the signal
// tail (S not a multiple of TS) and rows beyond the series are zero-filled by the hardware.
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    const uint32_t addr = smem_u32(bar);
    uint32_t done = 0;
    for (uint32_t spin = 0; !done; ++spin) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(addr), "r"(parity)
            : "memory");
        if (spin > (1u << 28)) __trap();  // watchdog: a broken pipeline must not hang the GPU
    }
}
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cta.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
        ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}

// first radix / remaining stages of a compile-time plan (the first stage is fused with the taper product)
template <typename PLAN> struct PlanHead;
template <int N, int R0, int... REST> struct PlanHead<ScStaticPlan<N, R0, REST...>> {
    static constexpr int n = N, r0 = R0;
    template <int NB, typename SYNC>
    static __device__ __forceinline__ cx<float>* rest(cx<float>* src, cx<float>* dst, const cx<float>* tws, int tid,
                                                      int nthreads, SYNC sync) {
        return ScStaticStages<float, N, R0, REST...>::template run<NB>(src, dst, tws, false, tid, nthreads, sync);
    }
};

struct MtParams {
    const float* x;
    long long N, T, S;
    const float* tapers;
    int n, K, step;
    long long w0, W, w_out0;
    int nfft, detrend;
    float scale;
    const cx<float>* tw;
    int layout, nfo;
    ScMap map;
    long long R;
    void* out;
    cx<float>* ws;
    ScFftPlan plan;
};

struct CtaSync {
    __device__ __forceinline__ void operator()() const { __syncthreads(); }
};
struct DynPlan {};  // runtime plan (any nfft)

// Shared-memory slab tile[j][TS] (one row of TS floats per sample): the taper product reads one float2 (two
// series) per thread with the SAMPLE index along the lanes, i.e. at a stride of one row -- 8 lanes per bank
// pair for TS = 8.  XOR-swizzling the float2 slot of a row with the sample bits just above the bank period makes
// those reads conflict-free while the row-wise loads/stores stay a permutation within each row.
template <int TS>
__device__ __forceinline__ int tile_ix(int j, int s) {
    constexpr int NP = TS / 2;
    if (NP <= 1) return j * TS + s;
    constexpr int SH = NP == 2 ? 3 : (NP == 4 ? 2 : (NP == 8 ? 1 : 0));  // row stride = NP bank pairs of 16
    return j * TS + ((((s >> 1) ^ ((j >> SH) & (NP - 1))) << 1) | (s & 1));
}

// slab element (sample j, series s): XOR-swizzled for the register-loaded tile, dense for the TMA-staged one (whose
// fused first FFT stage reads whole 16-byte quarter rows, which is conflict-free without a swizzle)
template <int TS, bool TMA>
__device__ __forceinline__ int tile_at(int j, int s) {
    return TMA ? j * TS + s : tile_ix<TS>(j, s);
}

template <int TS, bool WS, typename PLAN, bool TMA = false>
__global__ void __launch_bounds__(kThreads, WS ? 1 : 2) mt_fft_kernel(const __grid_constant__ CUtensorMap tmap,
                                                                      const MtParams p) {
    constexpr bool STATIC = !std::is_same<PLAN, DynPlan>::value;
    static_assert(!TMA || (STATIC && !WS && TS == 8), "the TMA path is the compile-time-plan, 8-series tile kernel");
    extern __shared__ __align__(128) unsigned char smem_raw[];
    constexpr int NP = TS / 2;  // complex FFTs per taper
    const int n = p.n, nfft = p.nfft;
    const int ncopy = n < nfft ? n : nfft;
    const long long tiles = (p.S + TS - 1) / TS;
    const long long total = p.W * p.T * tiles;

    float* tile = nullptr;
    cx<float>*bufA, *bufB;
    const cx<float>* tw;
    __shared__ double red[2][kThreads];
    __shared__ float trend_a[TS], trend_b[TS];
    __shared__ __align__(8) uint64_t slab_bar;
    uint32_t slab_parity = 0;
    if (TMA) {
        if (threadIdx.x == 0) {
            mbar_init(&slab_bar, 1);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
    }
    if (WS) {
        bufA = p.ws + (size_t)blockIdx.x * 2 * NP * nfft;
        bufB = bufA + (size_t)NP * nfft;
        tw = p.tw;
    } else {
        // TMA: the slab is the 128-byte aligned head of the buffer, padded to whole boxes
        const int tile_rows = TMA ? ((n + kTmaBoxRows - 1) / kTmaBoxRows) * kTmaBoxRows : 0;
        bufA = reinterpret_cast<cx<float>*>(smem_raw + (size_t)tile_rows * TS * 4);
        bufB = bufA + (size_t)NP * nfft;
        cx<float>* tws = bufB + (size_t)NP * nfft;
        tile = TMA ? reinterpret_cast<float*>(smem_raw) : reinterpret_cast<float*>(tws + nfft);
        if constexpr (STATIC) ScStaticFft<float, PLAN>::fill(tws, p.tw, threadIdx.x, kThreads);
        else for (int q = threadIdx.x; q < nfft; q += kThreads) tws[q] = p.tw[q];
        tw = tws;
        __syncthreads();
    }

    // one elected thread stages the slab of `it` with bulk tensor copies (see the TMA helpers above)
    auto issue_slab = [&](long long it) {
        if (threadIdx.x != 0) return;
        const long long ti = it % tiles, tt = (it / tiles) % p.T, ww = p.w0 + it / (tiles * p.T);
        const int nbox = (n + kTmaBoxRows - 1) / kTmaBoxRows;
        // earlier generic-proxy accesses of the slab (ordered by the barrier before this call) must be visible to
        // the async proxy before it overwrites the buffer
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        mbar_expect_tx(&slab_bar, (uint32_t)(nbox * kTmaBoxRows * TS * 4));
        for (int bx = 0; bx < nbox; ++bx)
            tma_load_3d(tile + (size_t)bx * kTmaBoxRows * TS, &tmap, &slab_bar, (int)(ti * TS), (int)tt,
                        (int)(ww * p.step) + bx * kTmaBoxRows);
    };
    for (long long item = blockIdx.x; item < total; item += gridDim.x) {
        const long long tile_i = item % tiles;
        const long long t = (item / tiles) % p.T;
        const long long wl = item / (tiles * p.T);
        const long long w = p.w0 + wl;
        const long long s0 = tile_i * TS;
        const float* xw = p.x + ((w * p.step) * p.T + t) * p.S + s0;  // + j*T*S + s
        const long long row = p.T * p.S;
        const int sl = threadIdx.x % TS;
        const int jr = threadIdx.x / TS;
        constexpr int JSTEP = kThreads / TS;
        const bool s_ok = (s0 + sl) < p.S;

        // ---- load slab + detrend statistics --------------------------------
        double sum = 0.0, sumu = 0.0;
        const double ubar = (n + 1.0) / (2.0 * n);
        // 8 independent loads in flight per thread before any dependent use (the slab load is the
        // exposed DRAM latency of the kernel)
        constexpr int LD = 8;
        int bad = 0;
        if (TMA) {
            if (item == blockIdx.x) issue_slab(item);  // later slabs were prefetched under the previous item's last taper
            mbar_wait(&slab_bar, slab_parity);
            slab_parity ^= 1;
            // detrend statistics + NaN/Inf scan from shared memory (8 lanes per row: conflict-free)
            for (int j = jr; j < n; j += JSTEP) {
                const float v = tile[j * TS + sl];
                bad |= !isfinite(v);
                sum += v;
                if (p.detrend == SC_DETREND_LINEAR) sumu += ((j + 1.0) / n - ubar) * v;
            }
        }
        for (int j0 = jr; j0 < n && !TMA; j0 += LD * JSTEP) {
            float v[LD];
#pragma unroll
            for (int u = 0; u < LD; ++u) {
                const int j = j0 + u * JSTEP;
                v[u] = (s_ok && j < n) ? __ldg(xw + (long long)j * row + sl) : 0.f;
            }
#pragma unroll
            for (int u = 0; u < LD; ++u) {
                const int j = j0 + u * JSTEP;
                if (j < n) {
                    if (!WS) tile[tile_ix<TS>(j, sl)] = v[u];
                    bad |= !isfinite(v[u]);
                    sum += v[u];
                    if (p.detrend == SC_DETREND_LINEAR) sumu += ((j + 1.0) / n - ubar) * v[u];
                }
            }
        }
        // A NaN/Inf would leak from one series into the one it shares a packed FFT with; such slabs
        // (the reference only warns about them, transforms.py:754-774) take the unpacked path below,
        // one series per complex FFT, so that exactly the affected channels come out non-finite.
        const int unpack = __syncthreads_or(bad);
        if (p.detrend != SC_DETREND_NONE) {
            red[0][threadIdx.x] = sum;
            red[1][threadIdx.x] = sumu;
            __syncthreads();
            if (threadIdx.x < TS) {
                double a = 0.0, b = 0.0;
                for (int q = 0; q < JSTEP; ++q) {
                    a += red[0][q * TS + threadIdx.x];
                    b += red[1][q * TS + threadIdx.x];
                }
                const double mean = a / n;
                double slope = 0.0;
                if (p.detrend == SC_DETREND_LINEAR && n > 1) slope = b / ((double(n) * n - 1.0) / (12.0 * n));
                // x' = x - slope*u - icpt, u = (j+1)/n
                trend_a[threadIdx.x] = (float)slope;
                trend_b[threadIdx.x] = (float)(mean - slope * ubar);
            }
        } else if (threadIdx.x < TS) {
            trend_a[threadIdx.x] = 0.f;
            trend_b[threadIdx.x] = 0.f;
        }
        __syncthreads();
        if (!WS && p.detrend != SC_DETREND_NONE && (!TMA || unpack)) {
            const float ta = trend_a[sl], tb = trend_b[sl];
            const float invn = 1.0f / n;
            for (int j = jr; j < ncopy; j += JSTEP) tile[tile_at<TS, TMA>(j, sl)] -= ta * ((j + 1) * invn) + tb;
            __syncthreads();
        }

        for (int k = 0; k < p.K; ++k) {
            const float* h = p.tapers + (size_t)k * n;
          for (int half = 0; half < (unpack ? 2 : 1); ++half) {
            // ---- taper product, two real series per complex sequence --------
            if (WS || unpack) {
                for (int idx = threadIdx.x; idx < NP * nfft; idx += kThreads) {
                    const int pp = idx % NP;
                    const int j = idx / NP;
                    cx<float> z = cmake<float>(0.f, 0.f);
                    if (j < ncopy) {
                        const float hv = __ldg(h + j);
                        const int sa = unpack ? half * NP + pp : 2 * pp;  // first (or only) series of this sequence
                        float v0, v1 = 0.f;
                        if (WS) {
                            const float u = (j + 1.0f) / n;
                            v0 = (s0 + sa < p.S) ? __ldg(xw + (long long)j * row + sa) : 0.f;
                            v0 -= trend_a[sa] * u + trend_b[sa];
                            if (!unpack) {
                                v1 = (s0 + sa + 1 < p.S) ? __ldg(xw + (long long)j * row + sa + 1) : 0.f;
                                v1 -= trend_a[sa + 1] * u + trend_b[sa + 1];
                            }
                        } else {
                            v0 = tile[tile_at<TS, TMA>(j, sa)];
                        }
                        z = cmake<float>(v0 * hv, v1 * hv);
                    }
                    bufA[(size_t)pp * nfft + j] = z;
                }
            } else if constexpr (TMA) {
                // Taper product FUSED into the first FFT stage: butterfly j of the first stage needs the samples
                // j + t*M (t < R0) of its sequence, so it reads them straight from the slab -- one 16-byte quarter
                // row (4 series = 2 packed sequences) per sample, detrend and taper applied in registers -- and
                // writes the stage output; the tapered sequences never exist in shared memory (one write + one
                // read of NP*nfft complex values and one barrier less per taper).
                constexpr int R0 = PlanHead<PLAN>::r0;
                constexpr int M = PlanHead<PLAN>::n / R0;
                const float invn = 1.0f / n;
                for (int idx = threadIdx.x; idx < 2 * M; idx += kThreads) {
                    const int hq = idx / M, j = idx - hq * M;  // quarter-row pair (series 4*hq .. 4*hq+3), butterfly
                    const float ta0 = trend_a[4 * hq], ta1 = trend_a[4 * hq + 1], ta2 = trend_a[4 * hq + 2],
                                ta3 = trend_a[4 * hq + 3];
                    const float tb0 = trend_b[4 * hq], tb1 = trend_b[4 * hq + 1], tb2 = trend_b[4 * hq + 2],
                                tb3 = trend_b[4 * hq + 3];
                    cx<float> va[R0], vb[R0];
#pragma unroll
                    for (int tt = 0; tt < R0; ++tt) {
                        const int row = j + tt * M;
                        float4 xv = make_float4(0.f, 0.f, 0.f, 0.f);
                        float hv = 0.f;
                        if (row < ncopy) {
                            xv = *reinterpret_cast<const float4*>(tile + row * TS + 4 * hq);
                            hv = __ldg(h + row);
                        }
                        const float u = (row + 1) * invn;
                        va[tt] = cmake<float>((xv.x - (ta0 * u + tb0)) * hv, (xv.y - (ta1 * u + tb1)) * hv);
                        vb[tt] = cmake<float>((xv.z - (ta2 * u + tb2)) * hv, (xv.w - (ta3 * u + tb3)) * hv);
                    }
                    sc_dft<float, R0>(va, false, (const cx<float>*)0, PlanHead<PLAN>::n);
                    sc_dft<float, R0>(vb, false, (const cx<float>*)0, PlanHead<PLAN>::n);
                    cx<float>* da = bufA + (size_t)(2 * hq) * nfft + j * R0;
                    cx<float>* db = da + nfft;
#pragma unroll
                    for (int q = 0; q < R0; ++q) {
                        da[q] = va[q];
                        db[q] = vb[q];
                    }
                }
            } else {
                // one thread per sample j: the taper value is loaded once and applied to all TS series
#pragma unroll 2
                for (int j = threadIdx.x; j < nfft; j += kThreads) {
                    const float hv = j < ncopy ? __ldg(h + j) : 0.f;
                    const int jt = j < ncopy ? j : 0;
#pragma unroll
                    for (int pp = 0; pp < NP; ++pp) {
                        const float2 v = *reinterpret_cast<const float2*>(tile + tile_ix<TS>(jt, 2 * pp));
                        bufA[(size_t)pp * nfft + j] = cmake<float>(v.x * hv, v.y * hv);
                    }
                }
            }
            __syncthreads();
            if (TMA && !unpack && k == p.K - 1 && item + gridDim.x < total)
                issue_slab(item + gridDim.x);  // the slab is dead after the last taper's first stage: prefetch the next
            const cx<float>* res;
            if constexpr (TMA) {
                if (unpack) res = ScStaticFft<float, PLAN>::template run<NP>(bufA, bufB, tw, false, threadIdx.x, kThreads, CtaSync());
                else res = PlanHead<PLAN>::template rest<NP>(bufA, bufB, tw, threadIdx.x, kThreads, CtaSync());
            } else if constexpr (STATIC)
                res = ScStaticFft<float, PLAN>::template run<NP>(bufA, bufB, tw, false, threadIdx.x, kThreads, CtaSync());
            else
                res = sc_cta_fft<float>(bufA, bufB, NP, nfft, p.plan, tw, false);

            const long long wo = p.w_out0 + wl;
            if (unpack) {
                // ---- one series per sequence: X_s(f) = Z(f) --------------------------
                const long long plane = p.R * p.S;
                const long long b = wo * p.map.bw + t * p.map.bt + k * p.map.bk;
                const long long r = wo * p.map.rw + t * p.map.rt + k * p.map.rk;
                for (int idx = threadIdx.x; idx < p.nfo * NP; idx += kThreads) {
                    const int pp = idx % NP;
                    const int f = idx / NP;
                    const int sgl = half * NP + pp;
                    if (s0 + sgl >= p.S) continue;
                    const cx<float> z = res[(size_t)pp * nfft + f];
                    if (p.layout == SC_LAYOUT_PLANAR) {
                        float* out = reinterpret_cast<float*>(p.out);
                        out[((b * p.nfo + f) * 2 + 0) * plane + r * p.S + s0 + sgl] = p.scale * z.x;
                        out[((b * p.nfo + f) * 2 + 1) * plane + r * p.S + s0 + sgl] = p.scale * z.y;
                    } else {
                        float2* out = reinterpret_cast<float2*>(p.out);
                        out[((((wo * p.T + t) * p.K + k) * p.nfo) + f) * p.S + s0 + sgl] = make_float2(p.scale * z.x, p.scale * z.y);
                    }
                }
                __syncthreads();
                continue;
            }
            // ---- unpack the two real spectra and store -----------------------
            if (p.layout == SC_LAYOUT_PLANAR) {
                const long long b = wo * p.map.bw + t * p.map.bt + k * p.map.bk;
                const long long r = wo * p.map.rw + t * p.map.rt + k * p.map.rk;
                float* out = reinterpret_cast<float*>(p.out);
                const long long plane = p.R * p.S;
                const float sc = 0.5f * p.scale;
                if (TS % 4 == 0 && (p.S & 3) == 0 && s0 + TS <= p.S) {
                    constexpr int QV = TS / 4 > 0 ? TS / 4 : 1;  // float4 stores per (f, plane) row
                    for (int idx = threadIdx.x; idx < p.nfo * 2 * QV; idx += kThreads) {
                        const int qv = idx % QV;
                        const int c = (idx / QV) & 1;
                        const int f = idx / (2 * QV);
                        const int fm = f == 0 ? 0 : nfft - f;
                        const cx<float> a1 = res[(size_t)(2 * qv) * nfft + f], a2 = res[(size_t)(2 * qv) * nfft + fm];
                        const cx<float> b1 = res[(size_t)(2 * qv + 1) * nfft + f], b2 = res[(size_t)(2 * qv + 1) * nfft + fm];
                        float4 v;
                        if (c == 0) v = make_float4(a1.x + a2.x, a1.y + a2.y, b1.x + b2.x, b1.y + b2.y);
                        else v = make_float4(a1.y - a2.y, a2.x - a1.x, b1.y - b2.y, b2.x - b1.x);
                        v.x *= sc; v.y *= sc; v.z *= sc; v.w *= sc;
                        *reinterpret_cast<float4*>(out + ((b * p.nfo + f) * 2 + c) * plane + r * p.S + s0 + 4 * qv) = v;
                    }
                } else {
                    for (int idx = threadIdx.x; idx < p.nfo * 2 * TS; idx += kThreads) {
                        const int s = idx % TS;
                        const int c = (idx / TS) & 1;
                        const int f = idx / (2 * TS);
                        if (s0 + s >= p.S) continue;
                        const cx<float> z1 = res[(size_t)(s >> 1) * nfft + f];
                        const cx<float> z2 = res[(size_t)(s >> 1) * nfft + (f == 0 ? 0 : nfft - f)];
                        float v;
                        if ((s & 1) == 0) v = c == 0 ? (z1.x + z2.x) : (z1.y - z2.y);
                        else v = c == 0 ? (z1.y + z2.y) : (z2.x - z1.x);
                        out[((b * p.nfo + f) * 2 + c) * plane + r * p.S + s0 + s] = sc * v;
                    }
                }
            } else {
                float2* out = reinterpret_cast<float2*>(p.out);
                const long long base = ((wo * p.T + t) * p.K + k) * p.nfo;
                for (int idx = threadIdx.x; idx < p.nfo * TS; idx += kThreads) {
                    const int s = idx % TS;
                    const int f = idx / TS;
                    if (s0 + s >= p.S) continue;
                    const cx<float> z1 = res[(size_t)(s >> 1) * nfft + f];
                    const cx<float> z2 = res[(size_t)(s >> 1) * nfft + (f == 0 ? 0 : nfft - f)];
                    float2 v;
                    if ((s & 1) == 0) v = make_float2(z1.x + z2.x, z1.y - z2.y);
                    else v = make_float2(z1.y + z2.y, z2.x - z1.x);
                    v.x *= 0.5f * p.scale;
                    v.y *= 0.5f * p.scale;
                    out[(base + f) * p.S + s0 + s] = v;
                }
            }
            __syncthreads();
          }
        }
        if (TMA && unpack && item + gridDim.x < total) issue_slab(item + gridDim.x);  // unpacked slabs: no prefetch
    }
}

size_t mt_smem_bytes(int ts, int n, int nfft) {
    return (size_t)2 * (ts / 2) * nfft * 8 + (size_t)nfft * 8 + (size_t)n * ts * 4;
}

int mt_pick_ts(int n, int nfft) {
    const size_t cap = (size_t)sc_max_smem_optin() - 20 * 1024;  // static smem + slack
    const int cand[3] = {8, 4, 2};
    for (int i = 0; i < 3; ++i)
        if (mt_smem_bytes(cand[i], n, nfft) <= cap / 2) return cand[i];
    for (int i = 0; i < 3; ++i)
        if (mt_smem_bytes(cand[i], n, nfft) <= cap) return cand[i];
    return 0;  // workspace mode
}

template <int TS, bool WS, typename PLAN = DynPlan, bool TMA = false>
int mt_launch(const MtParams& p, size_t smem, long long grid, cudaStream_t st, const CUtensorMap* tmap = nullptr) {
    if (smem > 48 * 1024)
        SC_CUDA_OK(cudaFuncSetAttribute(mt_fft_kernel<TS, WS, PLAN, TMA>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        (int)smem));
    CUtensorMap none;
    if (!tmap) {
        memset(&none, 0, sizeof(none));
        tmap = &none;
    }
    mt_fft_kernel<TS, WS, PLAN, TMA><<<(unsigned)grid, kThreads, smem, st>>>(*tmap, p);
    SC_LAUNCH_OK();
    return SC_OK;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn mt_get_encode() {
    static EncodeTiledFn fn = nullptr;
    static bool tried = false;
    if (tried) return fn;
    tried = true;
    const char* off = getenv("SC_B200_DISABLE_TMA_FFT");
    if (off && off[0] == '1') return nullptr;
    void* sym = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &qres) != cudaSuccess ||
        qres != cudaDriverEntryPointSuccess) {
        (void)cudaGetLastError();
        return nullptr;
    }
    fn = reinterpret_cast<EncodeTiledFn>(sym);
    return fn;
}

// tensor map of the series x (N, T, S) float32 with box {8 signals, 1 trial, 256 samples}; false when the series
// does not meet the TMA constraints (16-byte aligned base and row strides)
bool mt_make_tmap(const MtParams& p, CUtensorMap* tmap) {
    EncodeTiledFn enc = mt_get_encode();
    if (!enc || (p.S & 3) != 0 || (reinterpret_cast<uintptr_t>(p.x) & 15) != 0) return false;
    if (p.S >= (1LL << 31) || p.T >= (1LL << 31) || p.N >= (1LL << 31)) return false;
    const cuuint64_t gdim[3] = {(cuuint64_t)p.S, (cuuint64_t)p.T, (cuuint64_t)p.N};
    const cuuint64_t gstr[2] = {(cuuint64_t)p.S * 4, (cuuint64_t)p.T * p.S * 4};  // bytes, dims 1..2
    const cuuint32_t box[3] = {8, 1, (cuuint32_t)kTmaBoxRows};
    const cuuint32_t estr[3] = {1, 1, 1};
    return enc(tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(p.x), gdim, gstr, box, estr,
               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

size_t mt_smem_bytes_tma(int n, int nfft) {
    const size_t rows = (size_t)((n + kTmaBoxRows - 1) / kTmaBoxRows) * kTmaBoxRows;
    return rows * 8 * 4 + (size_t)2 * 4 * nfft * 8 + (size_t)nfft * 8;
}

__global__ void repack_kernel(const float2* __restrict__ coef, long long W, long long T, long long K, long long nfft,
                              long long S, int nfo, ScMap map, long long R, float* __restrict__ out) {
    const long long total = W * T * K * nfo * S;
    const long long plane = R * S;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        const long long s = i % S;
        long long q = i / S;
        const long long f = q % nfo;
        q /= nfo;
        const long long k = q % K;
        q /= K;
        const long long t = q % T;
        const long long w = q / T;
        const float2 v = coef[(((w * T + t) * K + k) * nfft + f) * S + s];
        const long long b = w * map.bw + t * map.bt + k * map.bk;
        const long long r = w * map.rw + t * map.rt + k * map.rk;
        const long long o = ((b * nfo + f) * 2) * plane + r * S + s;
        out[o] = v.x;
        out[o + plane] = v.y;
    }
}



// NaN/Inf scan of the input (transforms.py:754-774), one pass at HBM rate: flag[0] |= 1 if any sample is non-finite.
// `head` scalar samples bring the pointer to 16-byte alignment, then float4 loads, then the scalar tail.
__global__ void nonfinite_kernel(const float* __restrict__ x, long long n, int head, int* flag) {
    const float4* x4 = reinterpret_cast<const float4*>(x + head);
    const long long n4 = (n - head) >> 2;
    int bad = 0;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
        const float4 v = __ldcs(x4 + i);
        // (v - v) is 0 for finite values and NaN for Inf / NaN
        const float t = (v.x - v.x) + (v.y - v.y) + (v.z - v.z) + (v.w - v.w);
        bad |= !(t == 0.f);
    }
    if (blockIdx.x == 0) {
        const long long tail0 = head + (n4 << 2);
        for (long long i = threadIdx.x; i < head + (n - tail0); i += blockDim.x) {
            const float v = x[i < head ? i : tail0 + (i - head)];
            bad |= !(v - v == 0.f);
        }
    }
    if (__syncthreads_or(bad) && threadIdx.x == 0) atomicOr(flag, 1);
}

}  // namespace

extern "C" int sc_nonfinite_flag(const float* x, int64_t n, int* flag, void* stream) {
    SC_CHECK_ARG(flag && n >= 0, "sc_nonfinite_flag: bad argument");
    if (n == 0) return SC_OK;
    SC_CHECK_ARG(x && (reinterpret_cast<uintptr_t>(x) & 3) == 0, "sc_nonfinite_flag: x must be a 4-byte aligned pointer");
    long long head = ((16 - (long long)(reinterpret_cast<uintptr_t>(x) & 15)) & 15) / 4;
    if (head > n) head = n;
    long long blocks = (((n - head) >> 2) + 255) / 256;
    const long long cap = (long long)sc_num_sms() * 16;
    blocks = blocks > cap ? cap : (blocks < 1 ? 1 : blocks);
    nonfinite_kernel<<<(unsigned)blocks, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(x, n, (int)head, flag);
    SC_LAUNCH_OK();
    return SC_OK;
}

namespace {
}  // namespace

extern "C" int64_t sc_mt_fft_workspace_bytes(int n, int nfft) {
    if (n < 1 || nfft < 1) return 0;
    if (mt_pick_ts(n, nfft) != 0) return 0;
    return (int64_t)kWsCtas * 2 * nfft * 8;
}

extern "C" int sc_mt_fft(const float* x, int64_t N, int64_t T, int64_t S, const float* tapers, int n, int K, int step,
                         int64_t w0, int64_t W, int64_t w_out0, int nfft, int detrend, float scale,
                         const void* twiddle, int layout, int n_freq_out, const int64_t* map, int64_t n_reduce,
                         void* out, void* workspace, int64_t workspace_bytes, void* stream) {
    SC_CHECK_ARG(x && tapers && twiddle && out, "sc_mt_fft: null pointer");
    SC_CHECK_ARG(N > 0 && T > 0 && S > 0 && n > 0 && K > 0 && step > 0 && nfft > 0, "sc_mt_fft: non-positive size");
    SC_CHECK_ARG(W >= 0 && w0 >= 0 && (w0 + W - 1) * (int64_t)step + n <= N || W == 0,
                 "sc_mt_fft: windows [%lld, %lld) exceed the series (N=%lld, n=%d, step=%d)", (long long)w0,
                 (long long)(w0 + W), (long long)N, n, step);
    SC_CHECK_ARG(n_freq_out >= 1 && n_freq_out <= nfft, "sc_mt_fft: n_freq_out %d outside [1, %d]", n_freq_out, nfft);
    SC_CHECK_ARG(detrend >= 0 && detrend <= 2, "sc_mt_fft: unknown detrend mode %d", detrend);
    SC_CHECK_ARG(layout == SC_LAYOUT_PLANAR || layout == SC_LAYOUT_REFERENCE, "sc_mt_fft: unknown layout %d", layout);
    SC_CHECK_ARG(layout != SC_LAYOUT_PLANAR || (map && n_reduce > 0), "sc_mt_fft: planar layout needs map and n_reduce");
    if (W == 0) return SC_OK;
    MtParams p;
    p.x = x; p.N = N; p.T = T; p.S = S; p.tapers = tapers; p.n = n; p.K = K; p.step = step;
    p.w0 = w0; p.W = W; p.w_out0 = w_out0; p.nfft = nfft; p.detrend = detrend; p.scale = scale;
    p.tw = reinterpret_cast<const cx<float>*>(twiddle);
    p.layout = layout; p.nfo = n_freq_out;
    if (map) p.map = ScMap{map[0], map[1], map[2], map[3], map[4], map[5]};
    else p.map = ScMap{0, 0, 0, 0, 0, 0};
    p.R = n_reduce; p.out = out; p.ws = nullptr;
    if (sc_fft_make_plan(nfft, &p.plan)) {
        sc_set_error("sc_mt_fft: cannot factorise nfft=%d", nfft);
        return SC_ERR_UNSUPPORTED;
    }
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const int ts = mt_pick_ts(n, nfft);
    const int tsz = ts ? ts : 2;
    const long long tiles = (S + tsz - 1) / tsz;
    const long long total = W * T * tiles;
    if (ts == 0) {
        const int64_t need = (int64_t)kWsCtas * 2 * nfft * 8;
        if (!workspace || workspace_bytes < need) {
            sc_set_error("sc_mt_fft: window of %d samples (nfft %d) needs a %lld-byte workspace", n, nfft,
                         (long long)need);
            return SC_ERR_WORKSPACE;
        }
        p.ws = reinterpret_cast<cx<float>*>(workspace);
        return mt_launch<2, true>(p, 0, total < kWsCtas ? total : kWsCtas, st);
    }
    const size_t smem = mt_smem_bytes(ts, n, nfft);
    const long long maxgrid = 1LL << 30;
    const long long grid = total < maxgrid ? total : maxgrid;
    if (ts == 8 && (nfft == 1000 || nfft == 120)) {
        CUtensorMap tmap;
        const size_t smem_tma = mt_smem_bytes_tma(n, nfft);
        if (smem_tma <= ((size_t)sc_max_smem_optin() - 20 * 1024) / 2 && mt_make_tmap(p, &tmap)) {
            if (nfft == 1000) return mt_launch<8, false, ScPlan1000, true>(p, smem_tma, grid, st, &tmap);
            return mt_launch<8, false, ScPlan120, true>(p, smem_tma, grid, st, &tmap);
        }
    }
    if (ts == 8 && nfft == 1000) return mt_launch<8, false, ScPlan1000>(p, smem, grid, st);
    if (ts == 8 && nfft == 120) return mt_launch<8, false, ScPlan120>(p, smem, grid, st);
    switch (ts) {
        case 8: return mt_launch<8, false>(p, smem, grid, st);
        case 4: return mt_launch<4, false>(p, smem, grid, st);
        default: return mt_launch<2, false>(p, smem, grid, st);
    }
}

extern "C" int sc_repack_coefficients(const void* coef_c64, int64_t W, int64_t T, int64_t K, int64_t nfft, int64_t S,
                                      int n_freq_out, const int64_t* map, int64_t n_reduce, float* out,
                                      void* stream) {
    SC_CHECK_ARG(coef_c64 && out && map, "sc_repack_coefficients: null pointer");
    SC_CHECK_ARG(W > 0 && T > 0 && K > 0 && nfft > 0 && S > 0 && n_reduce > 0, "sc_repack_coefficients: bad size");
    SC_CHECK_ARG(n_freq_out >= 1 && n_freq_out <= nfft, "sc_repack_coefficients: n_freq_out out of range");
    const long long total = W * T * K * (long long)n_freq_out * S;
    long long blocks = (total + 255) / 256;
    const long long cap = (long long)sc_num_sms() * 16;
    if (blocks > cap) blocks = cap;
    repack_kernel<<<(unsigned)blocks, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
        reinterpret_cast<const float2*>(coef_c64), W, T, K, nfft, S, n_freq_out,
        ScMap{map[0], map[1], map[2], map[3], map[4], map[5]}, n_reduce, out);
    SC_LAUNCH_OK();
    return SC_OK;
}
