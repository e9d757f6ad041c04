// Fused reduce-scatter + epilogue over NVLink peer memory (trial-sharded mode, SURVEY.md section 8e partition B).
//
// Every rank of the reduce group holds the PARTIAL expected cross-spectral matrices of a window chunk (its share
// of the trials x tapers observations, already scaled by 1 / n_observations) in a buffer that its peers can map
// (torch symmetric memory: CUDA VMM allocations exchanged over the process group; NVLink 5 through NVSwitch gives
// every peer the full 900 GB/s).  Instead of an NCCL reduce_scatter followed by a diagonal-extraction kernel and an
// epilogue kernel (three passes over the reduced chunk), ONE kernel per rank PULLS the rows that rank owns from all
// peers with 16-byte loads over the fabric, adds them in rank order (deterministic, unlike a ring whose order
// depends on the chunk schedule), writes the reduced matrix for the Wilson / Granger stage, extracts the power
// (the real diagonal) and -- when asked -- applies one coherence-family epilogue on the fly.  One CTA owns one
// (window, frequency) matrix: the S diagonal sums first (power, kept in shared memory for the normalisation),
// then the S x S elements in float4 = two complex entries per thread and step.
//
// Peer loads bypass L1 (ld.global.cg): the buffers are rewritten every third chunk by another GPU and this SM's L1
// is not coherent with remote writes.  Ordering between "partials written" and "peers read" is the caller's
// (a device-side barrier of the symmetric-memory handle on the same stream, connectivity.py).
#include "sc_common.cuh"

namespace {

constexpr int kMaxPeers = 16;
constexpr double kEps64 = 2.220446049250313e-16;

struct PeerPtrs {
    const float4* p[kMaxPeers];
};

__device__ __forceinline__ float4 ld_peer(const float4* p) { return __ldcg(p); }

// A SMALL persistent grid (like an NCCL kernel: it must fit beside the Wilson / Granger kernel of the previous
// chunk, which occupies every register of the SMs it runs on) of 512-thread CTAs; every thread keeps U x WORLD
// 16-byte peer loads in flight (U = 4 for 2 peers ... 1 for 8+), ~4-8 MB over the whole grid, which covers the
// bandwidth-delay product of the NVLink fabric (~2 MB).
constexpr int kPeerThreads = 512;
constexpr int kPeerCtas = 64;

template <int WORLD>
__global__ void __launch_bounds__(kPeerThreads) peer_reduce_csm_kernel(PeerPtrs peers, int world, long long mat0, long long n_mat,
                                                                       int S, float4* __restrict__ out_csm,
                                                                       float* __restrict__ out_power, int measure,
                                                                       float* __restrict__ out_measure) {
    extern __shared__ float spow[];  // [S] sqrt of the reduced power of this matrix
    const float qnan = __int_as_float(0x7fc00000);
    constexpr int NW = WORLD > 0 ? WORLD : kMaxPeers;
    constexpr int U = WORLD == 2 ? 4 : (WORLD == 4 ? 2 : 1);
    const int nw = WORLD > 0 ? WORLD : world;
    const long long quads = (long long)S * S / 2;  // float4 = 2 complex entries
    for (long long m = blockIdx.x; m < n_mat; m += gridDim.x) {
        const long long src0 = (mat0 + m) * quads;
        const long long dst0 = m * quads;
        // ---- power = real diagonal of the reduced matrix ----
        if (out_power || measure >= 0) {
            for (int i = threadIdx.x; i < S; i += blockDim.x) {
                const long long e = (long long)i * S + i;  // complex index within the matrix
                float part[NW];
#pragma unroll
                for (int r = 0; r < NW; ++r)
                    part[r] = r < nw ? __ldcg(&reinterpret_cast<const float2*>(peers.p[r] + src0)[e]).x : 0.f;
                float d = 0.f;
#pragma unroll
                for (int r = 0; r < NW; ++r) d += part[r];
                if (out_power) out_power[m * S + i] = d;
                spow[i] = sqrtf(d);
            }
            __syncthreads();
        }
        // ---- the matrix: sum over the peers in rank order, optional coherence-family epilogue ----
        for (long long q0 = (long long)threadIdx.x; q0 < quads; q0 += (long long)U * blockDim.x) {
            float4 v[NW][U];
#pragma unroll
            for (int r = 0; r < NW; ++r)
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    const long long q = q0 + (long long)u * blockDim.x;
                    v[r][u] = (r < nw && q < quads) ? ld_peer(peers.p[r] + src0 + q) : make_float4(0.f, 0.f, 0.f, 0.f);
                }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const long long q = q0 + (long long)u * blockDim.x;
                if (q >= quads) break;
                float4 acc = v[0][u];
#pragma unroll
                for (int r = 1; r < NW; ++r) {
                    acc.x += v[r][u].x; acc.y += v[r][u].y; acc.z += v[r][u].z; acc.w += v[r][u].w;
                }
                if (out_csm) __stcs(out_csm + dst0 + q, acc);
                if (measure >= 0) {
                    const long long c0 = 2 * q;               // complex index of acc.xy; acc.zw is c0 + 1 (same row: S even)
                    const int i = (int)(c0 / S), j = (int)(c0 - (long long)i * S);
                    const float pi = spow[i];
                    float re[2] = {acc.x, acc.z}, im[2] = {acc.y, acc.w}, o[2], o2[2];
#pragma unroll
                    for (int e = 0; e < 2; ++e) {
                        float norm = pi * spow[j + e];
                        norm = norm < (float)kEps64 ? (float)kEps64 : norm;  // connectivity.py:649-652
                        const float inv = 1.0f / norm;
                        const float cr = re[e] * inv, ci = im[e] * inv;
                        const bool diag = i == j + e;
                        if (measure == SC_M_COHERENCY) {
                            o[e] = diag ? qnan : cr;
                            o2[e] = diag ? qnan : ci;
                        } else if (measure == SC_M_COHERENCE_MAG) {
                            float mg = fmaf(cr, cr, ci * ci);
                            mg = mg < 0.f ? 0.f : (mg > 1.f ? 1.f : mg);
                            o[e] = diag ? qnan : mg;
                        } else if (measure == SC_M_COHERENCE_PHASE) {
                            o[e] = diag ? qnan : atan2f(ci, cr);
                        } else {
                            const float a = fabsf(ci);
                            o[e] = a > 1.f ? 1.f : a;
                        }
                    }
                    if (measure == SC_M_COHERENCY)
                        __stcs(reinterpret_cast<float4*>(out_measure) + dst0 + q, make_float4(o[0], o2[0], o[1], o2[1]));
                    else
                        __stcs(reinterpret_cast<float2*>(out_measure) + dst0 + q, make_float2(o[0], o[1]));
                }
            }
        }
        __syncthreads();  // spow is rewritten by the next matrix
    }
}

template <int WORLD>
__global__ void __launch_bounds__(256) peer_reduce_flat_kernel(PeerPtrs peers, int world, long long q0, long long n_quads,
                                                               float4* __restrict__ out) {
    const int nw = WORLD > 0 ? WORLD : world;
    for (long long q = blockIdx.x * (long long)blockDim.x + threadIdx.x; q < n_quads; q += (long long)gridDim.x * blockDim.x) {
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int r = 0; r < (WORLD > 0 ? WORLD : kMaxPeers); ++r) {
            if (r >= nw) break;
            const float4 v = ld_peer(peers.p[r] + q0 + q);
            acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
        }
        __stcs(out + q, acc);
    }
}

}  // namespace

extern "C" int sc_peer_reduce_csm(const void* const* peers /* host array of device pointers */, int world, int64_t mat0,
                                  int64_t n_mat, int S, void* out_csm_c64, float* out_power, int measure,
                                  void* out_measure, void* stream) {
    SC_CHECK_ARG(peers && world >= 1 && world <= kMaxPeers, "sc_peer_reduce_csm: world size %d outside [1, %d]", world, kMaxPeers);
    SC_CHECK_ARG(n_mat >= 0 && mat0 >= 0 && S > 0 && S % 2 == 0, "sc_peer_reduce_csm: bad shape (S must be even)");
    if (n_mat == 0) return SC_OK;  // a rank may own no row of a ragged last chunk
    SC_CHECK_ARG(measure == -1 || (measure >= SC_M_COHERENCY && measure <= SC_M_IMAG_COHERENCE),
                 "sc_peer_reduce_csm: measure %d is not a coherence-family epilogue", measure);
    SC_CHECK_ARG(measure == -1 || out_measure, "sc_peer_reduce_csm: measure needs an output");
    PeerPtrs pp;
    for (int r = 0; r < kMaxPeers; ++r) pp.p[r] = r < world ? reinterpret_cast<const float4*>(peers[r]) : nullptr;
    for (int r = 0; r < world; ++r)
        SC_CHECK_ARG(pp.p[r] && (reinterpret_cast<uintptr_t>(pp.p[r]) & 15) == 0, "sc_peer_reduce_csm: peer %d pointer unaligned", r);
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const unsigned grid = (unsigned)(n_mat < kPeerCtas ? n_mat : kPeerCtas);
    const size_t smem = (size_t)S * sizeof(float);
    float4* oc = reinterpret_cast<float4*>(out_csm_c64);
    float* om = reinterpret_cast<float*>(out_measure);
    switch (world) {
        case 2: peer_reduce_csm_kernel<2><<<grid, kPeerThreads, smem, st>>>(pp, world, mat0, n_mat, S, oc, out_power, measure, om); break;
        case 4: peer_reduce_csm_kernel<4><<<grid, kPeerThreads, smem, st>>>(pp, world, mat0, n_mat, S, oc, out_power, measure, om); break;
        case 8: peer_reduce_csm_kernel<8><<<grid, kPeerThreads, smem, st>>>(pp, world, mat0, n_mat, S, oc, out_power, measure, om); break;
        default: peer_reduce_csm_kernel<0><<<grid, kPeerThreads, smem, st>>>(pp, world, mat0, n_mat, S, oc, out_power, measure, om);
    }
    SC_LAUNCH_OK();
    return SC_OK;
}

extern "C" int sc_peer_reduce(const void* const* peers /* host array of device pointers */, int world, int64_t offset_bytes,
                              int64_t n_bytes, void* out, void* stream) {
    SC_CHECK_ARG(peers && out && world >= 1 && world <= kMaxPeers, "sc_peer_reduce: bad argument");
    SC_CHECK_ARG(offset_bytes >= 0 && n_bytes >= 0 && offset_bytes % 16 == 0 && n_bytes % 16 == 0,
                 "sc_peer_reduce: offset and size must be multiples of 16 bytes");
    if (n_bytes == 0) return SC_OK;
    PeerPtrs pp;
    for (int r = 0; r < kMaxPeers; ++r) pp.p[r] = r < world ? reinterpret_cast<const float4*>(peers[r]) : nullptr;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const long long nq = n_bytes / 16, q0 = offset_bytes / 16;
    long long blocks = (nq + 255) / 256;
    const long long cap = (long long)sc_num_sms() * 16;
    if (blocks > cap) blocks = cap;
    float4* o = reinterpret_cast<float4*>(out);
    switch (world) {
        case 2: peer_reduce_flat_kernel<2><<<(unsigned)blocks, 256, 0, st>>>(pp, world, q0, nq, o); break;
        case 4: peer_reduce_flat_kernel<4><<<(unsigned)blocks, 256, 0, st>>>(pp, world, q0, nq, o); break;
        case 8: peer_reduce_flat_kernel<8><<<(unsigned)blocks, 256, 0, st>>>(pp, world, q0, nq, o); break;
        default: peer_reduce_flat_kernel<0><<<(unsigned)blocks, 256, 0, st>>>(pp, world, q0, nq, o);
    }
    SC_LAUNCH_OK();
    return SC_OK;
}
