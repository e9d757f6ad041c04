// Shared pieces of the Wilson / Granger kernels (wilson.cu, granger_herm.cu).
#pragma once
#include "fft_device.cuh"
#include "sc_common.cuh"

namespace scw {

typedef cx<double> cd;
constexpr int kThreads = 256;
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
