// Shared pieces of the Wilson / Granger kernels (wilson.cu, granger_herm.cu).
#pragma once
#include "fft_device.cuh"
#include "sc_common.cuh"

namespace scw {

typedef cx<double> cd;
#ifndef SCW_THREADS
#define SCW_THREADS 256  // threads per CTA of the Wilson / Granger kernels (one CTA = one problem)
#endif
constexpr int kThreads = SCW_THREADS;
constexpr int kWarps = kThreads / 32;
constexpr double kEps64 = 2.220446049250313e-16;
constexpr double kTikhonov = 1e-12;  // connectivity.py:79

struct W2Params {
    const void* csm;
    const float* power;
    long long B;
    int F, nfft, herm;
    long long S;
    const int* pairs;
    long long n_pairs;
    double tol;
    int max_iter;
    const cd* tw;
    void* out;
    int* iters;
    int* flags;
    unsigned char* ws;
    int use_smem;
    int tail;  // 1: sum the geometric tail of the reference iteration in closed form (Hermitian kernel)
    int mixed; // 1: first iterations in fp32 on the row-scaled problem (Hermitian kernel, needs tw32)
    const cx<float>* tw32;
    unsigned long long* exec_counters;  // [4] or NULL: fp32 iterations, fp64 iterations, tail steps, problems
    ScFftPlan plan;
};

__device__ __forceinline__ cd cdiv1(cd a) {  // 1/a
    const double d = a.x * a.x + a.y * a.y;
    return cmake<double>(a.x / d, -a.y / d);
}
__device__ __forceinline__ cd cmulc(cd a, cd b) {  // a * conj(b)
    return cmake<double>(a.x * b.x + a.y * b.y, a.y * b.x - a.x * b.y);
}
__device__ __forceinline__ cd cneg(cd a) { return cmake<double>(-a.x, -a.y); }

template <int NV>
__device__ __forceinline__ void block_sum(double* v, double* red) {
#pragma unroll
    for (int q = 0; q < NV; ++q)
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v[q] += __shfl_xor_sync(0xffffffffu, v[q], o);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    __syncthreads();
    if (lane == 0)
#pragma unroll
        for (int q = 0; q < NV; ++q) red[q * kWarps + warp] = v[q];
    __syncthreads();
#pragma unroll
    for (int q = 0; q < NV; ++q) {
        double s = 0.0;
#pragma unroll
        for (int w = 0; w < kWarps; ++w) s += red[q * kWarps + w];
        v[q] = s;
    }
}

// Block sums of up to FOUR doubles per thread with a transposing butterfly: the first two exchange rounds halve the
// number of values a lane carries (lane l ends up owning value 2 (l & 1) + ((l >> 1) & 1)), three more rounds finish
// the warp, four lanes per warp publish, ONE barrier (double-buffered like block_max_nonneg), and the cross-warp
// stage is again a butterfly over the lane bits that index the warp.  26 shuffles + 9 additions per thread instead of
// 40 shuffles, 32 shared loads and 52 additions (block_sum<4>): the Granger kernel's fixed part spent 30 % of its
// instructions there (profiles/r02_ncu_granger_fixed_part.txt).  Every thread returns all four totals.
__device__ __forceinline__ double shfl_xor_f64(double x, int m) { return __shfl_xor_sync(0xffffffffu, x, m); }
__device__ __forceinline__ void block_sum4(double (&v)[4], double* red /* [2 * 4 * kWarps] */, int& phase) {
    static_assert(kWarps <= 8 && (kWarps & (kWarps - 1)) == 0, "the cross-warp butterfly indexes warps by lane bits 2..4");
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const bool b0 = lane & 1, b1 = lane & 2;
    double k0 = b0 ? v[2] : v[0], k1 = b0 ? v[3] : v[1];
    k0 += shfl_xor_f64(b0 ? v[0] : v[2], 1);
    k1 += shfl_xor_f64(b0 ? v[1] : v[3], 1);
    double k = (b1 ? k1 : k0) + shfl_xor_f64(b1 ? k0 : k1, 2);
    k += shfl_xor_f64(k, 4);
    k += shfl_xor_f64(k, 8);
    k += shfl_xor_f64(k, 16);
    const int q = 2 * (lane & 1) + ((lane >> 1) & 1);
    double* r = red + phase * 4 * kWarps;
    phase ^= 1;
    if (lane < 4) r[q * kWarps + warp] = k;
    __syncthreads();
    k = r[q * kWarps + ((lane >> 2) & (kWarps - 1))];
#pragma unroll
    for (int m = 4; m < 4 * kWarps; m <<= 1) k += shfl_xor_f64(k, m);
    v[0] = __shfl_sync(0xffffffffu, k, 0);
    v[1] = __shfl_sync(0xffffffffu, k, 2);
    v[2] = __shfl_sync(0xffffffffu, k, 1);
    v[3] = __shfl_sync(0xffffffffu, k, 3);
}

template <int NV>
__device__ __forceinline__ void block_maxn(double (&v)[NV], double* red) {
#pragma unroll
    for (int q = 0; q < NV; ++q)
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v[q] = fmax(v[q], __shfl_xor_sync(0xffffffffu, v[q], o));
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    __syncthreads();
    if (lane == 0)
#pragma unroll
        for (int q = 0; q < NV; ++q) red[q * kWarps + warp] = v[q];
    __syncthreads();
#pragma unroll
    for (int q = 0; q < NV; ++q) {
        double m = red[q * kWarps];
#pragma unroll
        for (int w = 1; w < kWarps; ++w) m = fmax(m, red[q * kWarps + w]);
        v[q] = m;
    }
}

__device__ __forceinline__ double block_max(double v, double* red) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    __syncthreads();
    if (lane == 0) red[warp] = v;
    __syncthreads();
    double m = red[0];
#pragma unroll
    for (int w = 1; w < kWarps; ++w) m = fmax(m, red[w]);
    return m;
}

// ---- single-barrier maxima of NON-NEGATIVE values ---------------------------------------------------------
// For non-negative IEEE numbers the unsigned bit pattern is order preserving, so the warp level is one
// redux.sync (two for a 64-bit pattern: high words, then the low words of the lanes that hold the maximal high
// word) instead of five shuffle+compare+select rounds.  Per-warp results go through a double-buffered shared
// array (`phase` toggles), which makes ONE __syncthreads per call sufficient: a buffer is rewritten two calls
// later, i.e. after every thread has passed the barrier of the call in between.  NaNs are dropped (fmax rule).
__device__ __forceinline__ unsigned long long warp_max_u64(unsigned long long x) {
    const unsigned hi = (unsigned)(x >> 32), lo = (unsigned)x;
    const unsigned mh = __reduce_max_sync(0xffffffffu, hi);
    const unsigned ml = __reduce_max_sync(0xffffffffu, hi == mh ? lo : 0u);
    return ((unsigned long long)mh << 32) | ml;
}

__device__ __forceinline__ float block_max_nonneg(float v, unsigned* red /* [2 * kWarps] */, int& phase) {
    const unsigned u = __reduce_max_sync(0xffffffffu, __float_as_uint(v == v ? v : 0.f));
    unsigned* r = red + phase * kWarps;
    phase ^= 1;
    if ((threadIdx.x & 31) == 0) r[threadIdx.x >> 5] = u;
    __syncthreads();
    unsigned m = r[0];
#pragma unroll
    for (int w = 1; w < kWarps; ++w) m = max(m, r[w]);
    return __uint_as_float(m);
}

template <int NV>
__device__ __forceinline__ void block_maxn_nonneg(double (&v)[NV], unsigned long long* red /* [2 * NV * kWarps] */,
                                                  int& phase) {
    static_assert(kWarps <= 32, "second level is one warp-wide redux");
    unsigned long long* r = red + phase * NV * kWarps;
    phase ^= 1;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
    for (int q = 0; q < NV; ++q) {
        const unsigned long long u = warp_max_u64((unsigned long long)__double_as_longlong(v[q] == v[q] ? v[q] : 0.0));
        if (lane == 0) r[q * kWarps + warp] = u;
    }
    __syncthreads();
#pragma unroll
    for (int q = 0; q < NV; ++q)
        v[q] = __longlong_as_double((long long)warp_max_u64(r[q * kWarps + (lane & (kWarps - 1))]));
}

__device__ __forceinline__ void decode_pair(long long k, long long S, int& i, int& j) {
    // k-th pair of combinations(range(S), 2) in lexicographic order
    const double t = 2.0 * S - 1.0;
    long long ii = (long long)floor((t - sqrt(t * t - 8.0 * (double)k)) * 0.5);
    if (ii < 0) ii = 0;
    if (ii > S - 2) ii = S - 2;
    while (ii > 0 && ii * (2 * S - ii - 1) / 2 > k) --ii;
    while ((ii + 1) * (2 * S - ii - 2) / 2 <= k) ++ii;
    const long long start = ii * (2 * S - ii - 1) / 2;
    i = (int)ii;
    j = (int)(ii + 1 + (k - start));
}


}  // namespace scw
