// Large-matrix (S > 32) building blocks of the full-matrix Wilson factorisation and the MVAR family:
// batched complex-fp64 GEMM, blocked in-place Gauss-Jordan inversion with partial pivoting, lag-0 Cholesky.
// Included by wilson_general.cu (inside its anonymous namespace) -- BASELINE config 5 asks for the directed
// transfer function of 512 channels, i.e. 512 x 512 solves per (window, frequency) and iteration
// (minimum_phase_decomposition.py:218-224 uses two LAPACK solves; connectivity.py:585-588, 1742-1748 two more).
//
// Matrices are row-major c128, batch-contiguous [count][S][S].  A "window state" array (0 = active) lets
// the Wilson loop skip converged / failed windows: matrix b belongs to window b / n_inner.
//
// Inversion.  In-place blocked Gauss-Jordan (panel width nb = 32 up to S = 420, 24 at S = 512, ... so that the
// S x nb panel fits in shared memory): for each column panel J
//   zb_panel_kernel   one CTA per matrix: unblocked in-place GJ with partial pivoting (rows >= current column)
//                     on the panel in shared memory -> T_J (S x nb: rows J hold A_JJ^-1, the others
//                     -A_iJ A_JJ^-1) and the pivot rows; then the same CTA gathers the row block
//                     RB = (P M)[J, :] and scatters the displaced rows (a pure gather/scatter: rows that leave
//                     J always land outside J and vice versa, so no ordering hazards between columns)
//   zb_update_kernel  64 x 64 tiles: M[i, c] <- (i in J ? 0 : M[i, c]) + T_J[i, :] RB[:, c] for c outside J,
//                     M[:, J] <- T_J   (rank-nb update: S^3 complex multiply-adds per inverse in total)
// and finally zb_unscramble_kernel undoes the row interchanges as a column permutation (out of place).
#pragma once

struct ZGemmParams {
    const cd* A;
    const cd* Bm;
    cd* C;
    long long a_so, a_si, b_so, b_si, c_so, c_si;  // element offset of batch (outer, inner) = outer*so + inner*si
    int n_inner;                                   // batch = outer * n_inner + inner
    int S;
    int conj_b;        // 0: C = A B     1: C = A B^H
    int add_identity;  // C += I
    const int* state;  // per outer index, may be null; != 0 -> skip (or copy A through, see passthrough)
    int passthrough;   // skipped batches copy A to C (keeps a ping-pong buffer pair consistent)
    double* err;       // per outer index, may be null: max |C - A| as an order-preserving bit pattern
    int upper_only;    // skip tiles strictly below the diagonal (the consumer reads i <= j only)
};

constexpr int kZT = 64;   // C tile edge
constexpr int kZK = 8;    // k tile
constexpr int kZP = kZT + 1;

// C = A op(B) (+ I): 64 x 64 tile per CTA, 256 threads as 16 x 16, 4 x 4 interleaved micro-tile per thread
// (rows ty + 16 i, columns tx + 16 j: conflict-free 16-byte shared loads, coalesced stores).
__global__ void __launch_bounds__(256, 2) zb_gemm_kernel(const ZGemmParams p) {
    __shared__ cd As[2][kZK][kZP];
    __shared__ cd Bs[2][kZK][kZP];
    const int S = p.S;
    const int tiles = (S + kZT - 1) / kZT;
    const long long batch = blockIdx.x / (tiles * tiles);
    const int tile = (int)(blockIdx.x - batch * tiles * tiles);
    const int m0 = (tile / tiles) * kZT, n0 = (tile % tiles) * kZT;
    if (p.upper_only && m0 > n0) return;
    const long long outer = batch / p.n_inner, inner = batch - outer * p.n_inner;
    const cd* A = p.A + outer * p.a_so + inner * p.a_si;
    const cd* Bm = p.Bm + outer * p.b_so + inner * p.b_si;
    cd* C = p.C + outer * p.c_so + inner * p.c_si;
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    if (p.state && p.state[outer] != 0) {
        if (p.passthrough)
            for (int e = tid; e < kZT * kZT; e += 256) {
                const int m = m0 + e / kZT, n = n0 + e % kZT;
                if (m < S && n < S) C[(size_t)m * S + n] = A[(size_t)m * S + n];
            }
        return;
    }
    const cd zero = cmake<double>(0.0, 0.0);
    cd acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = zero;

    // global -> register staging: two elements of each operand tile per thread
    cd ra[2], rb[2];
    auto fetch = [&](int k0) {
#pragma unroll
        for (int q = 0; q < 2; ++q) {
            const int e = tid + 256 * q;
            {   // A tile: rows m (64) x k (8), k contiguous in memory
                const int m = e / kZK, k = e % kZK;
                ra[q] = (m0 + m < S && k0 + k < S) ? A[(size_t)(m0 + m) * S + k0 + k] : zero;
            }
            if (p.conj_b) {  // B^H: element (k, n) = conj(Bm[n][k]) -- same access shape as the A tile
                const int n = e / kZK, k = e % kZK;
                rb[q] = (n0 + n < S && k0 + k < S) ? cconj(Bm[(size_t)(n0 + n) * S + k0 + k]) : zero;
            } else {  // rows k (8) x n (64), n contiguous
                const int k = e / kZT, n = e % kZT;
                rb[q] = (k0 + k < S && n0 + n < S) ? Bm[(size_t)(k0 + k) * S + n0 + n] : zero;
            }
        }
    };
    auto stash = [&](int buf) {
#pragma unroll
        for (int q = 0; q < 2; ++q) {
            const int e = tid + 256 * q;
            As[buf][e % kZK][e / kZK] = ra[q];
            if (p.conj_b)
                Bs[buf][e % kZK][e / kZK] = rb[q];
            else
                Bs[buf][e / kZT][e % kZT] = rb[q];
        }
    };
    fetch(0);
    stash(0);
    __syncthreads();
    int buf = 0;
    for (int k0 = 0; k0 < S; k0 += kZK) {
        const bool more = k0 + kZK < S;
        if (more) fetch(k0 + kZK);
#pragma unroll
        for (int k = 0; k < kZK; ++k) {
            cd a[4], b[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) a[i] = As[buf][k][ty + 16 * i];
#pragma unroll
            for (int j = 0; j < 4; ++j) b[j] = Bs[buf][k][tx + 16 * j];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    acc[i][j].x = fma(a[i].x, b[j].x, fma(-a[i].y, b[j].y, acc[i][j].x));
                    acc[i][j].y = fma(a[i].x, b[j].y, fma(a[i].y, b[j].x, acc[i][j].y));
                }
        }
        if (more) stash(buf ^ 1);
        __syncthreads();
        buf ^= 1;
    }
    double err2 = 0.0;
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int m = m0 + ty + 16 * i, n = n0 + tx + 16 * j;
            if (m < S && n < S) {
                cd v = acc[i][j];
                if (p.add_identity && m == n) v.x += 1.0;
                if (p.err) {
                    const cd d = csub(v, A[(size_t)m * S + n]);
                    const double d2 = d.x * d.x + d.y * d.y;
                    err2 = (d2 == d2) ? fmax(err2, d2) : __longlong_as_double(0x7ff0000000000000LL);
                }
                C[(size_t)m * S + n] = v;
            }
        }
    if (p.err) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) err2 = fmax(err2, __shfl_xor_sync(0xffffffffu, err2, o));
        if ((tid & 31) == 0 && err2 > 0.0)
            atomicMax(reinterpret_cast<unsigned long long*>(p.err + outer),
                      (unsigned long long)__double_as_longlong(sqrt(err2)));
    }
}

struct ZInvParams {
    cd* M;        // [count][S][S], inverted in place (up to the final column permutation)
    cd* TJ;       // [count][S][nb]
    cd* RB;       // [count][nb][S]
    int* ipiv;    // [count][S]
    int* perm;    // [count][S] column permutation that undoes the row interchanges (written by the last panel)
    int* bad;     // [count] set to 1 when a zero / non-finite pivot is met (may be null)
    const int* state;
    int n_inner;
    int S, nb, j0, w;
};

// One CTA per matrix: in-place Gauss-Jordan on the S x w panel (columns j0 .. j0+w-1) in shared memory with
// partial pivoting, then the row interchange of every other column as a gather (RB) / scatter.
__global__ void __launch_bounds__(256) zb_panel_kernel(const ZInvParams p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const long long b = blockIdx.x;
    if (p.state && p.state[b / p.n_inner] != 0) return;
    const int S = p.S, w = p.w, j0 = p.j0, ld = p.nb + 1;
    cd* pan = reinterpret_cast<cd*>(smem_raw);            // [S][ld]
    cd* rowp = pan + (size_t)S * ld;                      // [nb]
    double* redv = reinterpret_cast<double*>(rowp + p.nb);  // [8]
    int* redi = reinterpret_cast<int*>(redv + 8);         // [8]
    int* piv_s = redi + 8;                                // [nb] pivot row of each panel column
    int* jsrc = piv_s + p.nb;                             // [nb] original row that ends at position j0+k
    int* orow = jsrc + p.nb;                              // [nb] outside rows that receive a panel row
    int* osrc = orow + p.nb;                              // [nb] ... and which original (panel) row
    __shared__ int n_out, singular;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    cd* M = p.M + (size_t)b * S * S;
    for (int e = tid; e < S * w; e += 256) {
        const int r = e / w, c = e - r * w;
        pan[r * ld + c] = M[(size_t)r * S + j0 + c];
    }
    if (tid == 0) singular = 0;
    __syncthreads();
    for (int c = 0; c < w; ++c) {
        const int prow = j0 + c;
        double bestv = -1.0;
        int best = prow;
        for (int r = prow + tid; r < S; r += 256) {
            const cd v = pan[r * ld + c];
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
        if (lane == 0) {
            redv[warp] = bestv;
            redi[warp] = best;
        }
        __syncthreads();
        if (warp == 0) {
            bestv = redv[0];
            best = redi[0];
            for (int q = 1; q < 8; ++q)
                if (redv[q] > bestv || (redv[q] == bestv && redi[q] < best)) {
                    bestv = redv[q];
                    best = redi[q];
                }
            if (!(bestv > 0.0) || !isfinite(bestv)) {
                if (lane == 0) singular = 1;
                best = prow;
            }
            if (lane == 0) piv_s[c] = best;
            if (best != prow)
                for (int cc = lane; cc < w; cc += 32) {
                    const cd t = pan[prow * ld + cc];
                    pan[prow * ld + cc] = pan[best * ld + cc];
                    pan[best * ld + cc] = t;
                }
            __syncwarp();
            const cd pinv = cdiv1(pan[prow * ld + c]);
            for (int cc = lane; cc < w; cc += 32) rowp[cc] = (cc == c) ? pinv : cmul(pan[prow * ld + cc], pinv);
        }
        __syncthreads();
        const cd pinv = rowp[c];
        for (int r = tid; r < S; r += 256) {
            cd* row = pan + r * ld;
            if (r == prow) {
                for (int cc = 0; cc < w; ++cc) row[cc] = rowp[cc];
            } else {
                const cd f = row[c];
                for (int cc = 0; cc < w; ++cc) row[cc] = (cc == c) ? cneg(cmul(f, pinv)) : csub(row[cc], cmul(f, rowp[cc]));
            }
        }
        __syncthreads();
    }
    // T_J and the pivots
    cd* TJ = p.TJ + (size_t)b * S * p.nb;
    for (int e = tid; e < S * w; e += 256) {
        const int r = e / w, c = e - r * w;
        TJ[(size_t)r * p.nb + c] = pan[r * ld + c];
    }
    int* ipiv = p.ipiv + (size_t)b * S;
    if (tid < w) ipiv[j0 + tid] = piv_s[tid];
    if (tid == 0) {
        if (singular && p.bad) p.bad[b] = 1;
        // net effect of the w interchanges on the rows: jsrc[k] = original row now at j0+k; displaced panel rows
        int no = 0;
        for (int k = 0; k < w; ++k) jsrc[k] = j0 + k;
        for (int k = 0; k < w; ++k) {
            const int pv = piv_s[k];
            if (pv == j0 + k) continue;
            if (pv < j0 + w) {
                const int t = jsrc[k];
                jsrc[k] = jsrc[pv - j0];
                jsrc[pv - j0] = t;
            } else {
                int m = 0;
                while (m < no && orow[m] != pv) ++m;
                const int cur = (m < no) ? osrc[m] : pv;
                if (m == no) orow[no++] = pv;
                osrc[m] = jsrc[k];
                jsrc[k] = cur;
            }
        }
        n_out = no;
    }
    __syncthreads();
    cd* RB = p.RB + (size_t)b * p.nb * S;
    const int no = n_out;
    for (int c = tid; c < S; c += 256) {
        if (c >= j0 && c < j0 + w) continue;
        for (int k = 0; k < w; ++k) RB[(size_t)k * S + c] = M[(size_t)jsrc[k] * S + c];
        // displaced rows come from ORIGINAL panel rows, which nobody overwrites in this phase
        for (int m = 0; m < no; ++m) M[(size_t)orow[m] * S + c] = M[(size_t)osrc[m] * S + c];
    }
    if (j0 + w >= S) {
        // Undoing the row interchanges = permuting the columns of the result: replaying the interchanges
        // backwards on an identity index vector gives perm(c) = s_{S-1}(...s_1(s_0(c))), s_k = (k <-> ipiv[k]).
        // Each column is traced independently (s_k only moves k and ipiv[k] >= k).
        __syncthreads();
        int* ip = reinterpret_cast<int*>(pan);  // the panel is dead
        for (int k = tid; k < S; k += 256) ip[k] = (k >= j0) ? piv_s[k - j0] : ipiv[k];
        __syncthreads();
        int* perm = p.perm + (size_t)b * S;
        for (int c = tid; c < S; c += 256) {
            int v = c;
            for (int k = 0; k < S; ++k) {
                const int pv = ip[k];
                v = (v == k) ? pv : ((v == pv) ? k : v);
            }
            perm[c] = v;
        }
    }
}

// M[i, c] <- (i in J ? 0 : M[i, c]) + sum_k T_J[i, k] RB[k, c]  (c outside J);  M[:, J] <- T_J
__global__ void __launch_bounds__(256, 2) zb_update_kernel(const ZInvParams p) {
    __shared__ cd Ts[8][kZP];
    __shared__ cd Rs[8][kZP];
    const int S = p.S, w = p.w, j0 = p.j0;
    const int tiles = (S + kZT - 1) / kZT;
    const long long b = blockIdx.x / (tiles * tiles);
    if (p.state && p.state[b / p.n_inner] != 0) return;
    const int tile = (int)(blockIdx.x - b * tiles * tiles);
    const int m0 = (tile / tiles) * kZT, n0 = (tile % tiles) * kZT;
    cd* M = p.M + (size_t)b * S * S;
    const cd* TJ = p.TJ + (size_t)b * S * p.nb;
    const cd* RB = p.RB + (size_t)b * p.nb * S;
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const cd zero = cmake<double>(0.0, 0.0);
    cd acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int m = m0 + ty + 16 * i, n = n0 + tx + 16 * j;
            const bool in_j = m >= j0 && m < j0 + w;
            acc[i][j] = (m < S && n < S && !in_j) ? M[(size_t)m * S + n] : zero;
        }
    for (int k0 = 0; k0 < w; k0 += 8) {  // panel widths are multiples of 8 (the last panel is zero-filled)
        __syncthreads();
        for (int e = tid; e < 8 * kZT; e += 256) {
            {
                const int m = e / 8, k = e % 8;
                Ts[k][m] = (m0 + m < S && k0 + k < w) ? TJ[(size_t)(m0 + m) * p.nb + k0 + k] : zero;
            }
            {
                const int k = e / kZT, n = e % kZT;
                Rs[k][n] = (n0 + n < S && k0 + k < w) ? RB[(size_t)(k0 + k) * S + n0 + n] : zero;
            }
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            cd a[4], bb[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) a[i] = Ts[k][ty + 16 * i];
#pragma unroll
            for (int j = 0; j < 4; ++j) bb[j] = Rs[k][tx + 16 * j];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    acc[i][j].x = fma(a[i].x, bb[j].x, fma(-a[i].y, bb[j].y, acc[i][j].x));
                    acc[i][j].y = fma(a[i].x, bb[j].y, fma(a[i].y, bb[j].x, acc[i][j].y));
                }
        }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int m = m0 + ty + 16 * i, n = n0 + tx + 16 * j;
            if (m < S && n < S)
                M[(size_t)m * S + n] = (n >= j0 && n < j0 + w) ? TJ[(size_t)m * p.nb + (n - j0)] : acc[i][j];
        }
}

// dst[r][c] = M[r][perm[c]]
__global__ void __launch_bounds__(256) zb_unscramble_kernel(const cd* Msrc, cd* dst, const int* perm_all, int S,
                                                             int rows_per_cta, const int* state, int n_inner) {
    extern __shared__ int perm[];  // [S]
    const int chunks = (S + rows_per_cta - 1) / rows_per_cta;
    const long long b = blockIdx.x / chunks;
    if (state && state[b / n_inner] != 0) return;
    const int r0 = (int)(blockIdx.x - b * chunks) * rows_per_cta;
    for (int c = threadIdx.x; c < S; c += 256) perm[c] = perm_all[(size_t)b * S + c];
    __syncthreads();
    const cd* M = Msrc + (size_t)b * S * S;
    cd* D = dst + (size_t)b * S * S;
    for (int e = threadIdx.x; e < rows_per_cta * S; e += 256) {
        const int r = r0 + e / S, c = e % S;
        if (r < S) D[(size_t)r * S + c] = M[(size_t)r * S + perm[c]];
    }
}

// dst = src + lambda I  (src c128 or, with src_real, f64 promoted to c128)
__global__ void zb_shift_copy_kernel(const cd* src, const double* src_real, double lam_host, const double* lam_dev,
                                     long long count, int S, cd* dst) {
    const double lam = lam_dev ? lam_dev[0] : lam_host;
    const long long n = count * S * S;
    for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < n; e += (long long)gridDim.x * blockDim.x) {
        const int rc = (int)(e % ((long long)S * S));
        cd v = src ? src[e] : cmake<double>(src_real[e], 0.0);
        if (rc / S == rc % S) v.x += lam;
        dst[e] = v;
    }
}

// widest panel (multiple of 8, at most 32) whose S x (nb + 1) c128 image fits in shared memory:
// 32 up to S = 420, 24 at S = 512, 8 at S = 1024
inline int zb_panel_width(int S) {
    int nb = (int)(((size_t)sc_max_smem_optin() - 4096) / ((size_t)S * sizeof(cd))) - 1;
    nb &= ~7;
    return nb > 32 ? 32 : (nb < 8 ? 8 : nb);
}

inline size_t zb_panel_smem(int S, int nb) {
    return (size_t)S * (nb + 1) * sizeof(cd) + (size_t)nb * sizeof(cd) + 8 * sizeof(double) + 8 * sizeof(int) +
           (size_t)4 * nb * sizeof(int);
}

// bytes of the TJ / RB / pivot scratch of a batched inversion
inline int64_t zb_inverse_scratch_bytes(int64_t count, int S) {
    const int nb = 32;  // upper bound of zb_panel_width: the query must not depend on the device
    return count * ((int64_t)2 * S * nb * (int64_t)sizeof(cd) + (int64_t)S * 8 + 16) + 256;
}

// Inverts the `count` matrices in `work` (destroyed) into `dst` (may not alias work).
inline int zb_inverse(cd* work, cd* dst, int64_t count, int S, const int* state, int n_inner, int* bad,
                      unsigned char* scratch, cudaStream_t st) {
    const int nb = zb_panel_width(S);
    ZInvParams q;
    q.M = work;
    q.TJ = reinterpret_cast<cd*>(scratch);
    q.RB = q.TJ + (size_t)count * S * nb;
    q.ipiv = reinterpret_cast<int*>(q.RB + (size_t)count * nb * S);
    q.perm = q.ipiv + (size_t)count * S;
    q.bad = bad;
    q.state = state;
    q.n_inner = n_inner;
    q.S = S;
    q.nb = nb;
    const size_t smem = zb_panel_smem(S, nb);
    if (smem > (size_t)sc_max_smem_optin()) {
        sc_set_error("blocked inverse: S=%d needs %zu bytes of shared memory for the panel", S, smem);
        return SC_ERR_UNSUPPORTED;
    }
    SC_CUDA_OK(cudaFuncSetAttribute(zb_panel_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int tiles = (S + kZT - 1) / kZT;
    for (int j0 = 0; j0 < S; j0 += nb) {
        q.j0 = j0;
        q.w = (S - j0 < nb) ? S - j0 : nb;
        zb_panel_kernel<<<(unsigned)count, 256, smem, st>>>(q);
        zb_update_kernel<<<(unsigned)(count * tiles * tiles), 256, 0, st>>>(q);
    }
    const int rows = 8, chunks = (S + rows - 1) / rows;
    zb_unscramble_kernel<<<(unsigned)(count * chunks), 256, (size_t)S * sizeof(int), st>>>(work, dst, q.perm, S, rows,
                                                                                          state, n_inner);
    SC_LAUNCH_OK();
    return SC_OK;
}

inline void zb_gemm(const ZGemmParams& g, long long batch, cudaStream_t st) {
    const int tiles = (g.S + kZT - 1) / kZT;
    zb_gemm_kernel<<<(unsigned)(batch * tiles * tiles), 256, 0, st>>>(g);
}

// In-place lower Cholesky of a real symmetric S x S matrix per window (global memory, one CTA per window);
// state[w] = 2 when the matrix is not positive definite (minimum_phase_decomposition.py:48-93).
__global__ void __launch_bounds__(1024) zb_cholesky_kernel(double* a_all, int S, int* state, int* iters, double* err) {
    extern __shared__ double col[];  // [S]
    __shared__ int bad;
    const long long w = blockIdx.x;
    double* a = a_all + (size_t)w * S * S;
    if (threadIdx.x == 0) bad = 0;
    __syncthreads();
    for (int j = 0; j < S; ++j) {
        const double d = a[(size_t)j * S + j];
        __syncthreads();  // everyone has read the pivot before it is overwritten by its square root
        if (!(d > 0.0) || !isfinite(d)) {  // uniform: every thread read the same value
            if (threadIdx.x == 0) bad = 1;
            break;
        }
        const double l = sqrt(d);
        for (int i = j + threadIdx.x; i < S; i += blockDim.x) {
            const double v = (i == j) ? l : a[(size_t)i * S + j] / l;
            col[i] = v;
            a[(size_t)i * S + j] = v;
        }
        __syncthreads();
        // trailing rows j+1 .. S-1, lower triangle: one warp per row, lanes along the row
        for (int i = j + 1 + (threadIdx.x >> 5); i < S; i += (blockDim.x >> 5)) {
            const double ci = col[i];
            for (int k = j + 1 + (threadIdx.x & 31); k <= i; k += 32) a[(size_t)i * S + k] -= ci * col[k];
        }
        __syncthreads();
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        state[w] = bad ? 2 : 0;
        iters[w] = 0;
        err[w] = 0.0;
    }
}

// G0[f] = L^T for every frequency (NaN when the Cholesky failed)
__global__ void zb_init_g_kernel(const double* l_all, const int* state, int F, int S, cd* g) {
    const long long w = blockIdx.y;
    const double* l = l_all + (size_t)w * S * S;
    const bool bad = state[w] == 2;
    const double qnan = __longlong_as_double(0x7ff8000000000000LL);
    const size_t n = (size_t)F * S * S;
    for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < n; idx += (size_t)gridDim.x * blockDim.x) {
        const int e = (int)(idx % ((size_t)S * S));
        const int i = e / S, j = e % S;
        double v = (j >= i) ? l[(size_t)j * S + i] : 0.0;
        if (bad) v = qnan;
        g[(size_t)w * n + idx] = cmake<double>(v, bad ? qnan : 0.0);
    }
}

// sigma[w] = H0 H0^T (real), one thread per entry
__global__ void zb_sigma_kernel(const double* h0, long long B, int S, double* sigma) {
    const long long n = B * S * S;
    for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < n; e += (long long)gridDim.x * blockDim.x) {
        const long long w = e / ((long long)S * S);
        const int rc = (int)(e - w * S * S), i = rc / S, j = rc % S;
        const double* h = h0 + (size_t)w * S * S;
        double acc = 0.0;
        for (int k = 0; k < S; ++k) acc += h[i * S + k] * h[j * S + k];
        sigma[e] = acc;
    }
}

// Row / column sums for the MVAR normalisations of large matrices (one CTA per (window, frequency)):
// rs[i] = (measure 1: nv_i) sum_k |H_ik|^2,  cs[j] = sum_k |A_kj|^2 / (measure 3: nv_k)
__global__ void __launch_bounds__(256) zb_mvar_sums_kernel(int measure, const cd* h, const cd* a, const double* sigma,
                                                            int F, int S, double* rs, double* cs) {
    const long long wf = blockIdx.x, w = wf / F;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (h)
        for (int i = warp; i < S; i += 8) {
            double acc = 0.0;
            for (int k = lane; k < S; k += 32) {
                const cd v = h[(size_t)wf * S * S + (size_t)i * S + k];
                acc += v.x * v.x + v.y * v.y;
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
            if (lane == 0) rs[(size_t)wf * S + i] = (measure == 1 ? sigma[(size_t)w * S * S + (size_t)i * S + i] : 1.0) * acc;
        }
    if (a)
        for (int j = threadIdx.x; j < S; j += 256) {
            double acc = 0.0;
            for (int k = 0; k < S; ++k) {
                const cd v = a[(size_t)wf * S * S + (size_t)k * S + j];
                const double nv = (measure == 3) ? sigma[(size_t)w * S * S + (size_t)k * S + k] : 1.0;
                acc += (v.x * v.x + v.y * v.y) / nv;
            }
            cs[(size_t)wf * S + j] = acc;
        }
}

__global__ void zb_mvar_measure_kernel(int measure, const cd* h, const cd* a, const double* sigma,
                                       const double* inflow_all, const double* rs, const double* cs, long long BF, int F,
                                       int S, float* out) {
    const long long n = BF * S * S;
    for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < n; e += (long long)gridDim.x * blockDim.x) {
        const long long wf = e / ((long long)S * S), w = wf / F;
        const int rc = (int)(e - wf * S * S), i = rc / S, j = rc % S;
        double h2 = 0.0, a2 = 0.0;
        if (h) {
            const cd v = h[e];
            h2 = v.x * v.x + v.y * v.y;
        }
        if (a) {
            const cd v = a[e];
            a2 = v.x * v.x + v.y * v.y;
        }
        const double nv_i = sigma ? sigma[(size_t)w * S * S + (size_t)i * S + i] : 1.0;
        double v;
        switch (measure) {
            case 0: v = h2 / rs[wf * S + i]; break;
            case 1: v = sqrt(nv_i) * h2 / sqrt(rs[wf * S + i]); break;
            case 2: v = a2 / cs[wf * S + j]; break;
            case 3: v = a2 / nv_i / cs[wf * S + j]; break;
            default: v = sqrt(h2 / inflow_all[w * S + i]) * sqrt(a2 / cs[wf * S + j]); break;
        }
        out[e] = (float)v;
    }
}

// Causal projection (plus operator, mpd.py:96-142) for large matrices: one CTA owns a 4 x 4 tile of matrix
// entries, so that every (frequency, tile row) access is a 64-byte segment instead of wg_plus_kernel's single
// 16-byte elements at a stride of S*S*16 bytes.  The 16 lag sequences are transformed as one batch (sequence
// stride N + 1: conflict-free when the lanes run over the entries).  For real time series (herm) only tiles on or
// above the diagonal are launched; the tile's entries i <= j also produce their mirror images (j, i).
constexpr int kPT = 4;
constexpr int kPE = kPT * kPT;

inline size_t zb_plus_smem(int nfft) { return ((size_t)2 * kPE * (nfft + 1) + nfft) * sizeof(cd); }

__global__ void __launch_bounds__(kThreads) wg_plus_tile_kernel(const WgParams p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int N = p.nfft, S = p.S, NS = N + 1;
    const size_t SS = (size_t)S * S;
    cd* ZA = reinterpret_cast<cd*>(smem_raw);
    cd* ZB = ZA + (size_t)kPE * NS;
    cd* tws = ZB + (size_t)kPE * NS;
    const long long w = blockIdx.y;
    if (p.state[w] != 0) return;
    const int T = (S + kPT - 1) / kPT;
    int ti, tj;
    if (p.herm) {  // blockIdx.x enumerates ti <= tj
        int k = blockIdx.x;
        ti = 0;
        while (k >= T - ti) {
            k -= T - ti;
            ++ti;
        }
        tj = ti + k;
    } else {
        ti = blockIdx.x / T;
        tj = blockIdx.x % T;
    }
    const int i0 = ti * kPT, j0 = tj * kPT;
    for (int q = threadIdx.x; q < N; q += kThreads) tws[q] = p.tw[q];
    cd* base = p.bp + (size_t)w * p.F * SS;
    const cd zero = cmake<double>(0.0, 0.0);
    for (int idx = threadIdx.x; idx < p.F * kPE; idx += kThreads) {
        const int f = idx / kPE, e = idx % kPE;
        const int i = i0 + e / kPT, j = j0 + e % kPT;
        const bool valid = i < S && j < S && (!p.herm || i <= j);
        const cd v = valid ? base[(size_t)f * SS + (size_t)i * S + j] : zero;
        ZA[e * NS + f] = v;
        if (p.herm && f != 0 && 2 * f != N) ZA[e * NS + N - f] = cconj(v);
    }
    __syncthreads();
    cd* c = sc_cta_fft<double, true>(ZA, ZB, kPE, NS, p.plan, tws, true);
    cd* o = (c == ZA) ? ZB : ZA;
    const double inv_n = 1.0 / N;
    const int kcut = (N + 1) / 2;
    for (int idx = threadIdx.x; idx < N * kPE; idx += kThreads) {
        const int k = idx / kPE, e = idx % kPE;
        const int i = i0 + e / kPT, j = j0 + e % kPT;
        cd y = zero;
        if (k < kcut) {
            const double wgt = k == 0 ? 0.5 * inv_n : inv_n;
            if (p.herm) {
                const double cij = c[e * NS + k].x;
                const double cji = (i == j) ? 0.0 : (k == 0 ? 0.0 : c[e * NS + N - k].x);
                y = cmake<double>(cij * wgt, cji * wgt);
            } else {
                y = (k == 0 && i > j) ? zero : cscale(c[e * NS + k], wgt);
            }
        }
        o[e * NS + k] = y;
    }
    __syncthreads();
    const cd* Q = sc_cta_fft<double, true>(o, c, kPE, NS, p.plan, tws, false);
    for (int idx = threadIdx.x; idx < p.F * kPE; idx += kThreads) {
        const int f = idx / kPE, e = idx % kPE;
        const int i = i0 + e / kPT, j = j0 + e % kPT;
        if (i >= S || j >= S) continue;
        if (p.herm) {
            if (i > j) continue;
            const cd a = Q[e * NS + f], m = Q[e * NS + (f == 0 ? 0 : N - f)];
            base[(size_t)f * SS + (size_t)i * S + j] = cmake<double>(0.5 * (a.x + m.x), 0.5 * (a.y - m.y));
            if (i != j)
                base[(size_t)f * SS + (size_t)j * S + i] = cmake<double>(0.5 * (a.y + m.y), 0.5 * (m.x - a.x));
        } else {
            base[(size_t)f * SS + (size_t)i * S + j] = Q[e * NS + f];
        }
    }
}
