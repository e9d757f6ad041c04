// Measurement aid: SIMT FMA throughput of the device (fp32 / fp64), the denominators bench.py uses for the
// Wilson / Granger stage ("report achieved FLOP/s against the FP64 / FP32 vector peak";
// MEASURED_PEAKS.json holds only HBM and bf16 tensor numbers).  Not on the product path.
#include "sc_common.cuh"

namespace {

constexpr int kPeakThreads = 256;
constexpr int kPeakCtasPerSm = 8;   // 2048 threads = every warp slot of an SM
constexpr int kChains = 8;          // independent dependent-FMA chains per thread (covers the 4-cycle FMA latency)

template <typename R>
__global__ void __launch_bounds__(kPeakThreads) fma_peak_kernel(R* __restrict__ out, int n, R a, R b) {
    R v[kChains];
#pragma unroll
    for (int q = 0; q < kChains; ++q) v[q] = (R)(threadIdx.x + q) * (R)1e-3;
    for (int i = 0; i < n; ++i) {
#pragma unroll
        for (int q = 0; q < kChains; ++q) v[q] = fma(v[q], a, b);
    }
    R s = (R)0;
#pragma unroll
    for (int q = 0; q < kChains; ++q) s += v[q];
    out[(size_t)blockIdx.x * kPeakThreads + threadIdx.x] = s;  // keeps the chains live
}

}  // namespace

extern "C" int64_t sc_simt_peak_scratch_bytes(void) {
    return (int64_t)sc_num_sms() * kPeakCtasPerSm * kPeakThreads * 8;
}

extern "C" int sc_simt_peak(int dtype, int fma_per_chain, void* scratch, double* out_flops, void* stream) {
    SC_CHECK_ARG(scratch && out_flops, "sc_simt_peak: null pointer");
    SC_CHECK_ARG(dtype == 0 || dtype == 1, "sc_simt_peak: dtype must be 0 (float32) or 1 (float64)");
    SC_CHECK_ARG(fma_per_chain > 0, "sc_simt_peak: fma_per_chain must be positive");
    const unsigned grid = (unsigned)(sc_num_sms() * kPeakCtasPerSm);
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    // a, b chosen so the chains stay finite (|a| < 1)
    if (dtype == 0)
        fma_peak_kernel<float><<<grid, kPeakThreads, 0, st>>>(reinterpret_cast<float*>(scratch), fma_per_chain, 0.999f, 1e-3f);
    else
        fma_peak_kernel<double><<<grid, kPeakThreads, 0, st>>>(reinterpret_cast<double*>(scratch), fma_per_chain, 0.999, 1e-3);
    SC_LAUNCH_OK();
    *out_flops = 2.0 * kChains * (double)fma_per_chain * kPeakThreads * (double)grid;
    return SC_OK;
}
