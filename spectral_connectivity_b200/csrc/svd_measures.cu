// canonical_coherence and global_coherence from the expected cross-spectral matrix
// (connectivity.py:745-895, 1953-2032, 2245-2279).
//
// The reference whitens every signal group with a batched SVD of its (S_g x T*K) coefficient matrix
// (U V^H), multiplies pairs of whitened groups and takes the top singular value; global coherence is the
// top singular value of the full (S x T*K) matrix inside a Python double loop over (window, frequency).
// Both only depend on the cross-spectral matrix C = X X^H / n:
//   canonical coherence(a,b) = sigma_max^2( C_aa^-1/2 C_ab C_bb^-1/2 ) = lambda_max( M M^H ),
//       M = L_a^-1 C_ab L_b^-H with C_gg = L_g L_g^H (any square root gives the same singular values)
//   global coherence         = lambda_max( C ), with its eigenvector
// so they are computed from the (already expectation-reduced, trial-shardable) CSM: one CTA per
// (window, frequency, group pair) does two Cholesky factorisations, two triangular solves and one
// Hermitian product in fp64 shared memory; the top eigenpair comes from repeated squaring of the
// trace-normalised matrix (P <- P^2 / tr P^2 converges to v v^H at rate (l2/l1)^(2^k)) followed by a
// Rayleigh quotient against the original matrix.
#include "wilson_common.cuh"

namespace {

using namespace scw;

constexpr int kMaxN = 64;       // largest group (canonical) / signal count (global) held in shared memory
constexpr int kMaxSquarings = 26;

// P (n x n, Hermitian PSD, trace-normalised in place) -> dominant eigenvector direction in `vec` (n entries),
// using Q as the ping-pong buffer.  All threads of the CTA participate.
__device__ void top_vector_by_squaring(cd* P, cd* Q, int n, cd* vec, double* red) {
    const int nn = n * n;
    __shared__ double tr_sh;
    __shared__ int best_sh;
    // normalise by the trace
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int i = 0; i < n; ++i) t += P[i * n + i].x;
        tr_sh = t;
    }
    __syncthreads();
    {
        const double t = tr_sh;
        const double s = t > 0.0 ? 1.0 / t : 0.0;
        for (int e = threadIdx.x; e < nn; e += blockDim.x) P[e] = cscale(P[e], s);
    }
    __syncthreads();
    cd* src = P;
    cd* dst = Q;
    for (int it = 0; it < kMaxSquarings; ++it) {
        // P is Hermitian: only the upper triangle is evaluated and mirrored (diagonal forced real), which also
        // keeps the iterate EXACTLY Hermitian -- an anti-Hermitian rounding component (e.g. the imaginary noise of
        // an fp32-born diagonal, or of a deflated matrix) would otherwise double with every squaring and run away
        for (int e = threadIdx.x; e < nn; e += blockDim.x) {
            const int i = e / n, j = e - i * n;
            if (j < i) continue;
            cd acc = cmake<double>(0.0, 0.0);
            for (int k = 0; k < n; ++k) acc = cadd(acc, cmul(src[i * n + k], src[k * n + j]));
            if (i == j) acc.y = 0.0;
            dst[e] = acc;
            if (i != j) dst[j * n + i] = cconj(acc);
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            double t = 0.0;
            for (int i = 0; i < n; ++i) t += dst[i * n + i].x;
            tr_sh = t;
        }
        __syncthreads();
        const double t = tr_sh;  // = tr(P^2) with tr(P) = 1: reaches 1 when P has rank one
        const double s = t > 0.0 ? 1.0 / t : 0.0;
        for (int e = threadIdx.x; e < nn; e += blockDim.x) dst[e] = cscale(dst[e], s);
        __syncthreads();
        cd* tmp = src;
        src = dst;
        dst = tmp;
        if (!(t > 0.0) || 1.0 - t < 1e-13) break;
    }
    if (threadIdx.x == 0) {
        int best = 0;
        double bv = -1.0;
        for (int i = 0; i < n; ++i)
            if (src[i * n + i].x > bv) {
                bv = src[i * n + i].x;
                best = i;
            }
        best_sh = best;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < n; i += blockDim.x) vec[i] = src[i * n + best_sh];
    __syncthreads();
    (void)red;
}

// Rayleigh quotient v^H A v / v^H v (A Hermitian n x n); result broadcast to all threads.
__device__ double rayleigh(const cd* A, const cd* v, int n, double* red) {
    double num = 0.0, den = 0.0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        cd acc = cmake<double>(0.0, 0.0);
        for (int k = 0; k < n; ++k) acc = cadd(acc, cmul(A[i * n + k], v[k]));
        num += v[i].x * acc.x + v[i].y * acc.y;  // Re(conj(v_i) * (A v)_i)
        den += v[i].x * v[i].x + v[i].y * v[i].y;
    }
    double vals[2] = {num, den};
    block_sum<2>(vals, red);
    return vals[1] > 0.0 ? vals[0] / vals[1] : 0.0;
}

// in-place lower Cholesky of the Hermitian n x n matrix A (lower triangle read); returns false if not PD
__device__ bool cta_cholesky(cd* A, int n) {
    __shared__ int bad_sh;
    if (threadIdx.x == 0) bad_sh = 0;
    __syncthreads();
    for (int k = 0; k < n; ++k) {
        __shared__ double dk;
        if (threadIdx.x == 0) {
            const double d = A[k * n + k].x;
            if (!(d > 0.0) || !isfinite(d)) bad_sh = 1;
            dk = sqrt(d > 0.0 ? d : 1.0);
            A[k * n + k] = cmake<double>(dk, 0.0);
        }
        __syncthreads();
        const double inv = 1.0 / dk;
        for (int i = k + 1 + threadIdx.x; i < n; i += blockDim.x) A[i * n + k] = cscale(A[i * n + k], inv);
        __syncthreads();
        const int m = n - k - 1;
        for (int e = threadIdx.x; e < m * m; e += blockDim.x) {
            const int i = k + 1 + e / m, j = k + 1 + e % m;
            if (j <= i) A[i * n + j] = csub(A[i * n + j], cmulc(A[i * n + k], A[j * n + k]));
        }
        __syncthreads();
    }
    const bool ok = bad_sh == 0;
    __syncthreads();  // the next call resets bad_sh
    return ok;
}

struct CanonParams {
    const float2* csm;  // [B][F][S][S]
    long long BF;
    int S, G;
    const int* gidx;  // [S] signal indices sorted by group
    const int* goff;  // [G+1]
    float* out;       // [B][F][G][G]
    int* flags;       // [B][F] or null
    int nmax;
};

__global__ void __launch_bounds__(256) canonical_kernel(const CanonParams p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ double red[4 * 8];
    const int npair = p.G * (p.G - 1) / 2;
    const long long bf = blockIdx.x / npair;
    int pk = (int)(blockIdx.x % npair);
    int ga = 0;
    while (pk >= p.G - 1 - ga) {
        pk -= p.G - 1 - ga;
        ++ga;
    }
    const int gb = ga + 1 + pk;
    const int na = p.goff[ga + 1] - p.goff[ga], nb = p.goff[gb + 1] - p.goff[gb];
    const int* ia = p.gidx + p.goff[ga];
    const int* ib = p.gidx + p.goff[gb];
    const size_t cap = (size_t)p.nmax * p.nmax;
    cd* Caa = reinterpret_cast<cd*>(smem_raw);
    cd* Cbb = Caa + cap;
    cd* Cab = Cbb + cap;
    cd* vec = Cab + cap;
    const float2* m = p.csm + (size_t)bf * p.S * p.S;
    for (int e = threadIdx.x; e < na * na; e += blockDim.x) {
        const float2 v = m[(size_t)ia[e / na] * p.S + ia[e % na]];
        Caa[e] = cmake<double>(v.x, v.y);
    }
    for (int e = threadIdx.x; e < nb * nb; e += blockDim.x) {
        const float2 v = m[(size_t)ib[e / nb] * p.S + ib[e % nb]];
        Cbb[e] = cmake<double>(v.x, v.y);
    }
    for (int e = threadIdx.x; e < na * nb; e += blockDim.x) {
        const float2 v = m[(size_t)ia[e / nb] * p.S + ib[e % nb]];
        Cab[e] = cmake<double>(v.x, v.y);
    }
    __syncthreads();
    const bool ok_a = cta_cholesky(Caa, na);
    const bool ok_b = cta_cholesky(Cbb, nb);
    // Y = La^-1 Cab (forward substitution, all columns in parallel)
    for (int k = 0; k < na; ++k) {
        const double inv = 1.0 / Caa[k * na + k].x;
        for (int c = threadIdx.x; c < nb; c += blockDim.x) Cab[k * nb + c] = cscale(Cab[k * nb + c], inv);
        __syncthreads();
        const int rows = na - k - 1;
        for (int e = threadIdx.x; e < rows * nb; e += blockDim.x) {
            const int i = k + 1 + e / nb, c = e % nb;
            Cab[i * nb + c] = csub(Cab[i * nb + c], cmul(Caa[i * na + k], Cab[k * nb + c]));
        }
        __syncthreads();
    }
    // M = Y Lb^-H : M Lb^H = Y, column by column
    for (int k = 0; k < nb; ++k) {
        const double inv = 1.0 / Cbb[k * nb + k].x;
        for (int r = threadIdx.x; r < na; r += blockDim.x) Cab[r * nb + k] = cscale(Cab[r * nb + k], inv);
        __syncthreads();
        const int cols = nb - k - 1;
        for (int e = threadIdx.x; e < na * cols; e += blockDim.x) {
            const int r = e / cols, c = k + 1 + e % cols;
            Cab[r * nb + c] = csub(Cab[r * nb + c], cmulc(Cab[r * nb + k], Cbb[c * nb + k]));
        }
        __syncthreads();
    }
    // N = M M^H (na x na) into Caa, P copy into Cbb
    for (int e = threadIdx.x; e < na * na; e += blockDim.x) {
        const int i = e / na, j = e % na;
        cd acc = cmake<double>(0.0, 0.0);
        for (int k = 0; k < nb; ++k) acc = cadd(acc, cmulc(Cab[i * nb + k], Cab[j * nb + k]));
        Caa[e] = acc;
        Cbb[e] = acc;
    }
    __syncthreads();
    top_vector_by_squaring(Cbb, Cab, na, vec, red);
    const double lam = rayleigh(Caa, vec, na, red);
    if (threadIdx.x == 0) {
        const float val = (ok_a && ok_b) ? (float)lam : __int_as_float(0x7fc00000);
        float* o = p.out + (size_t)bf * p.G * p.G;
        o[ga * p.G + gb] = val;
        o[gb * p.G + ga] = val;
        if (p.flags && !(ok_a && ok_b)) atomicOr(p.flags + bf, SC_FLAG_NOT_SPD);
    }
}

__global__ void __launch_bounds__(256) global_coherence_kernel(const float2* csm, long long BF, int S, float* value,
                                                               float2* vector) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ double red[4 * 8];
    const long long bf = blockIdx.x;
    const size_t nn = (size_t)S * S;
    cd* A = reinterpret_cast<cd*>(smem_raw);
    cd* P = A + nn;
    cd* Q = P + nn;
    cd* vec = Q + nn;
    const float2* m = csm + (size_t)bf * nn;
    for (int e = threadIdx.x; e < (int)nn; e += blockDim.x) {
        const cd v = cmake<double>(m[e].x, m[e].y);
        A[e] = v;
        P[e] = v;
    }
    __syncthreads();
    top_vector_by_squaring(P, Q, S, vec, red);
    const double lam = rayleigh(A, vec, S, red);
    double nrm = 0.0;
    for (int i = threadIdx.x; i < S; i += blockDim.x) nrm += vec[i].x * vec[i].x + vec[i].y * vec[i].y;
    double vals[1] = {nrm};
    block_sum<1>(vals, red);
    const double inv = vals[0] > 0.0 ? 1.0 / sqrt(vals[0]) : 0.0;
    if (threadIdx.x == 0) value[bf] = (float)lam;
    for (int i = threadIdx.x; i < S; i += blockDim.x)
        vector[(size_t)bf * S + i] = make_float2((float)(vec[i].x * inv), (float)(vec[i].y * inv));
}

}  // namespace

extern "C" int sc_canonical_coherence(const void* csm_c64, int64_t B, int F, int S, const int* group_index,
                                      const int* group_offsets, int n_groups, int max_group_size, float* out,
                                      int* out_flags, void* stream) {
    SC_CHECK_ARG(csm_c64 && group_index && group_offsets && out, "sc_canonical_coherence: null pointer");
    SC_CHECK_ARG(B > 0 && F > 0 && S > 0 && n_groups >= 2, "sc_canonical_coherence: need at least two groups");
    if (max_group_size < 1 || max_group_size > kMaxN) {
        sc_set_error("sc_canonical_coherence: group size %d outside [1, %d]", max_group_size, kMaxN);
        return SC_ERR_UNSUPPORTED;
    }
    CanonParams p;
    p.csm = reinterpret_cast<const float2*>(csm_c64); p.BF = B * (int64_t)F; p.S = S; p.G = n_groups;
    p.gidx = group_index; p.goff = group_offsets; p.out = out; p.flags = out_flags; p.nmax = max_group_size;
    const long long grid = p.BF * (n_groups * (n_groups - 1) / 2);
    SC_CHECK_ARG(grid < (1LL << 31), "sc_canonical_coherence: grid too large; split the batch");
    const size_t smem = ((size_t)3 * max_group_size * max_group_size + max_group_size) * sizeof(cd);
    if (smem > 48 * 1024)
        SC_CUDA_OK(cudaFuncSetAttribute(canonical_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    canonical_kernel<<<(unsigned)grid, 256, smem, reinterpret_cast<cudaStream_t>(stream)>>>(p);
    SC_LAUNCH_OK();
    return SC_OK;
}

int64_t sc_global_coherence_blocked_workspace(int64_t BF, int S);
int sc_global_coherence_blocked(const void* csm_c64, int64_t BF, int S, float* out_value, void* out_vector_c64,
                                void* workspace, int64_t workspace_bytes, void* stream);

extern "C" int64_t sc_global_coherence_workspace_bytes(int64_t BF, int S) {
    if (BF < 1 || S <= kMaxN) return 0;
    return sc_global_coherence_blocked_workspace(BF, S);
}

extern "C" int sc_global_coherence(const void* csm_c64, int64_t BF, int S, float* out_value, void* out_vector_c64,
                                   void* workspace, int64_t workspace_bytes, void* stream) {
    SC_CHECK_ARG(csm_c64 && out_value && out_vector_c64 && BF > 0, "sc_global_coherence: bad argument");
    SC_CHECK_ARG(BF < (1LL << 31), "sc_global_coherence: batch too large");
    if (S < 1) {
        sc_set_error("sc_global_coherence: S=%d", S);
        return SC_ERR_UNSUPPORTED;
    }
    if (S > kMaxN)  // tiled c128 GEMM squarings in global memory (wilson_general.cu)
        return sc_global_coherence_blocked(csm_c64, BF, S, out_value, out_vector_c64, workspace, workspace_bytes, stream);
    const size_t smem = ((size_t)3 * S * S + S) * sizeof(cd);
    if (smem > 48 * 1024)
        SC_CUDA_OK(cudaFuncSetAttribute(global_coherence_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    global_coherence_kernel<<<(unsigned)BF, 256, smem, reinterpret_cast<cudaStream_t>(stream)>>>(
        reinterpret_cast<const float2*>(csm_c64), BF, S, out_value, reinterpret_cast<float2*>(out_vector_c64));
    SC_LAUNCH_OK();
    return SC_OK;
}


// Deflation step of global_coherence(max_rank > 1) (connectivity.py:2245-2279 keeps the max_rank largest singular
// values): C <- C - lambda v v^H removes the eigenpair just found, so that the next call of sc_global_coherence on the
// same buffer returns the next one.  One thread per matrix element, c64 in place.
namespace {
__global__ void deflate_kernel(float2* __restrict__ csm, long long BF, int S, const float* __restrict__ value,
                               const float2* __restrict__ vec) {
    const long long nn = (long long)S * S;
    for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < BF * nn;
         e += (long long)gridDim.x * blockDim.x) {
        const long long bf = e / nn;
        const int i = (int)((e - bf * nn) / S), j = (int)((e - bf * nn) % S);
        const float lam = value[bf];
        const float2 vi = vec[bf * S + i], vj = vec[bf * S + j];
        float2 c = csm[e];
        // v_i conj(v_j) with individually rounded products (no FMA contraction): the update of (i, j) is then the
        // exact conjugate of the update of (j, i) and the deflated matrix stays exactly Hermitian
        c.x -= lam * __fadd_rn(__fmul_rn(vi.x, vj.x), __fmul_rn(vi.y, vj.y));
        c.y -= lam * __fadd_rn(__fmul_rn(vi.y, vj.x), -__fmul_rn(vi.x, vj.y));
        csm[e] = c;
    }
}
}  // namespace

extern "C" int sc_hermitian_deflate(void* csm_c64, int64_t BF, int S, const float* value, const void* vector_c64,
                                    void* stream) {
    SC_CHECK_ARG(csm_c64 && value && vector_c64 && BF > 0 && S > 0, "sc_hermitian_deflate: bad argument");
    const long long total = BF * (long long)S * S;
    long long blocks = (total + 255) / 256;
    const long long cap = (long long)sc_num_sms() * 32;
    if (blocks > cap) blocks = cap;
    deflate_kernel<<<(unsigned)blocks, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
        reinterpret_cast<float2*>(csm_c64), BF, S, value, reinterpret_cast<const float2*>(vector_c64));
    SC_LAUNCH_OK();
    return SC_OK;
}
