/* sc_b200.h -- C ABI of the B200-native multitaper spectral-connectivity hot path.
 *
 * The reference (Eden-Kramer-Lab/spectral_connectivity, pure Python) has no FFI;
 * its only backend seam is the module-level ``xp`` alias chosen at import time
 * (transforms.py:405-439, connectivity.py:31-65, minimum_phase_decomposition.py:14-26).
 * Each entry point below replaces the NumPy/SciPy (or CuPy) call sequence of one
 * reference function; the file:line of what it replaces is cited per function.
 * INTEGRATION.md shows the ctypes stub a reference maintainer would add.
 *
 * Conventions
 *  - every pointer is a DEVICE pointer unless it says "host"; the caller owns all
 *    memory (inputs, outputs, workspaces); the library never allocates or frees;
 *  - all work is enqueued asynchronously on ``stream`` (a cudaStream_t passed as void*);
 *  - return value 0 = success, negative = error (SC_ERR_*); sc_last_error() gives the
 *    message of the last failure on the calling thread;
 *  - "planar coefficients" Xp: float32 [B][F][2][R][S] -- B kept (batch) index,
 *    F frequency bins, plane 0 = real / 1 = imaginary, R reduced (observation) index,
 *    S signals (fastest).  (B, R) are built from (window, trial, taper) by the linear
 *    map  b = w*map[0] + t*map[1] + k*map[2],  r = w*map[3] + t*map[4] + k*map[5],
 *    which expresses all seven ``expectation_type``s of connectivity.py:67-75;
 *  - complex outputs are interleaved (re, im) float32 ("c64") or float64 ("c128").
 */
#ifndef SC_B200_H
#define SC_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SC_OK 0
#define SC_ERR_INVALID_ARGUMENT (-1)
#define SC_ERR_UNSUPPORTED (-2)
#define SC_ERR_WORKSPACE (-3)
#define SC_ERR_CUDA (-4)

/* detrend modes of Multitaper(detrend_type=...) -- transforms.py:1798-1915 */
#define SC_DETREND_NONE 0
#define SC_DETREND_CONSTANT 1
#define SC_DETREND_LINEAR 2

/* sc_mt_fft output layouts */
#define SC_LAYOUT_PLANAR 0    /* float32 [B][F][2][R][S], see above                         */
#define SC_LAYOUT_REFERENCE 1 /* c64 (W,T,K,F,S): the logical layout of Multitaper.fft()  */

/* sc_csm modes: what is accumulated over the R observations of each (b, f, i, j) */
#define SC_CSM_CROSS 0 /* sum x_i conj(x_j)                        -> c64 [B][F][S][S]      */
#define SC_CSM_PLV 1   /* sum x_i conj(x_j)/|x_i conj(x_j)|        -> c64 [B][F][S][S]      */
#define SC_CSM_PLI 2   /* 4 real planes [4][B][F][S][S]: sum sign(Im), sum |Im|,
                          sum Im^2, sum Im (diagonal Im forced to 0)                       */

/* sc_pairwise_epilogue measures */
#define SC_M_COHERENCY 0       /* in0 = csm c64, in1 = power -> c64, NaN diagonal           */
#define SC_M_COHERENCE_MAG 1   /* -> f32 |coherency|^2 clipped to [0,1], NaN diagonal       */
#define SC_M_COHERENCE_PHASE 2 /* -> f32 angle(coherency), NaN diagonal                     */
#define SC_M_IMAG_COHERENCE 3  /* -> f32 |Im csm|/norm clipped to [0,1]                     */
#define SC_M_PLV 4             /* in0 = E[x/|x|] c64 -> f32 magnitude                       */
#define SC_M_PPC 5             /* in0 = E[x/|x|] c64 -> f32 pairwise phase consistency      */
#define SC_M_PLI 6             /* in0 = PLI planes (means) -> f32 E[sign Im]                */
#define SC_M_WPLI 7            /* -> f32 E[Im]/E[|Im|] (weights < eps -> 1)                 */
#define SC_M_DPLI2 8           /* -> f32 debiased squared PLI                               */
#define SC_M_DWPLI2 9          /* -> f32 debiased squared wPLI (zero weights -> NaN)        */

/* Wilson / Granger per-problem status flags */
#define SC_FLAG_NOT_CONVERGED 1 /* max_iterations reached (minimum_phase_decomposition.py:318-322) */
#define SC_FLAG_NOT_SPD 2       /* lag-0 covariance not positive definite (:78-82); output NaN      */

int sc_version(void);
const char* sc_last_error(void);

/* Measurement aid (bench.py roofline denominators; SURVEY.md section 8d asks for the Wilson stage against the
 * FP32 / FP64 vector peak, which MEASURED_PEAKS.json does not hold): launches one kernel of dependent-chain FMAs,
 * 8 independent chains per thread, every SM fully occupied.  dtype 0 = float32, 1 = float64.  *out_flops (host)
 * receives the number of floating-point operations the launch performs (2 per FMA); the caller times the launch
 * with CUDA events on ``stream``.  scratch: device buffer of at least sc_simt_peak_scratch_bytes() bytes. */
int sc_simt_peak(int dtype, int fma_per_chain, void* scratch, double* out_flops /* host */, void* stream);
int64_t sc_simt_peak_scratch_bytes(void);

/* Multitaper.fft() -- transforms.py:1147-1171: sliding-window gather (:1311-1374),
 * detrend (:1798-1915), DPSS taper product + FFT(n=nfft)/fs (:1377-1405), fused.
 *  x        float32 (N,T,S), S fastest
 *  tapers   float32 [K][n], already multiplied by sqrt(fs) (transforms.py:1440)
 *  windows  w0 .. w0+W-1 of the series (window w starts at sample w*step) are transformed;
 *           they are written at output window index w_out0 + (w - w0)
 *  nfft     FFT length: zero-pads (nfft > n) or crops (nfft < n) like scipy.fft.fft(n=)
 *  scale    multiplies the result (1/fs in the reference, :1405)
 *  twiddle  c64 [nfft], exp(-2 pi i q/nfft)
 *  layout   SC_LAYOUT_PLANAR: out float32 [B][n_freq_out][2][n_reduce][S] via map[6]
 *           SC_LAYOUT_REFERENCE: out c64 (windows,T,K,n_freq_out,S), row = output window index
 *  n_freq_out  number of leading bins written (nfft//2+1 or nfft)
 *  workspace   only for windows too long for shared memory: sc_mt_fft_workspace_bytes()
 */
int sc_mt_fft(const float* x, int64_t N, int64_t T, int64_t S, const float* tapers, int n, int K, int step,
              int64_t w0, int64_t W, int64_t w_out0, int nfft, int detrend, float scale, const void* twiddle,
              int layout, int n_freq_out, const int64_t* map /* host [6] */, int64_t n_reduce, void* out,
              void* workspace, int64_t workspace_bytes, void* stream);
int64_t sc_mt_fft_workspace_bytes(int n, int nfft);

/* Connectivity(fourier_coefficients=...) boundary -- connectivity.py:277-364: re-lays a
 * (W,T,K,Nfft,S) c64 array (the Multitaper.fft() layout) into planar coefficients,
 * keeping the first n_freq_out bins. */
int sc_repack_coefficients(const void* coef_c64, int64_t W, int64_t T, int64_t K, int64_t nfft, int64_t S,
                           int n_freq_out, const int64_t* map /* host [6] */, int64_t n_reduce, float* out,
                           void* stream);

/* Connectivity._power -- connectivity.py:441-445: out[b][f][s] = scale * sum_r |X|^2 */
int sc_power(const float* xp, int64_t B, int64_t F, int64_t R, int64_t S, float scale, float* out, void* stream);

/* Connectivity._expectation_cross_spectral_matrix(fcn) -- connectivity.py:463-526 with
 * _cross_spectral_matrix (:447-461) / _complex_inner_product (:1799-1822) fused in, and the
 * per-observation fcn of PLV (:899-903) and the PLI family (:970-980, :1010-1028, :1090-1127).
 * Results are multiplied by ``scale`` (1/n_observations for the expectation). */
int sc_csm(const float* xp, int64_t B, int64_t F, int64_t R, int64_t S, float scale, int mode, void* out,
           void* stream);

/* Same contract as sc_csm, always the SIMT fp32 kernel (sc_csm dispatches SC_CSM_CROSS to the
 * tcgen05 tensor-core kernel when the shape qualifies); kept public so tests can cross-check. */
int sc_csm_simt(const float* xp, int64_t B, int64_t F, int64_t R, int64_t S, float scale, int mode, void* out,
                void* stream);

/* NaN/Inf scan of the float32 input (the UserWarning of transforms.py:754-774, connectivity.py:340): ORs 1 into
 * flag[0] (int32, device, zeroed by the caller) when any of the n samples is not finite. */
int sc_nonfinite_flag(const float* x, int64_t n, int* flag, void* stream);

/* power as the real diagonal of an expected cross-spectral matrix that has been computed anyway
 * (connectivity.py:441-445: E[|X_i|^2] = E[X_i conj X_i]): csm c64 [BF][S][S] -> f32 [BF][S].  Saves the
 * second pass over the coefficients that sc_power needs. */
int sc_power_from_csm(const void* csm_c64, int64_t BF, int64_t S, float* out, void* stream);

/* Packed upper triangle (diagonal included, row-major: (i, j >= i) -> i S - i (i - 1) / 2 + j - i) of a symmetric real
 * measure [BF][S][S] -> [BF][S (S + 1) / 2]: an opt-in result format that halves the device -> host bytes of
 * coherence-type results (no reference counterpart: the reference returns the full symmetric array, connectivity.py:675-702). */
int sc_pack_upper(const float* in, int64_t BF, int64_t S, float* out, void* stream);

/* coherency / coherence_magnitude / coherence_phase / imaginary_coherence (connectivity.py:632-743),
 * phase_locking_value / pairwise_phase_consistency (:905-931, :1129-1159), phase_lag_index,
 * weighted / debiased variants (:933-1127): element-wise epilogues on [B][F][S][S]. */
int sc_pairwise_epilogue(int measure, const void* in0, const float* in1, int64_t B, int64_t F, int64_t S,
                         double n_observations, void* out, void* stream);

/* phase_slope_index (connectivity.py:1587-1650, _inner_combination :1652-1676): for every (b, i, j)
 * Im sum_{q1 < q2} conj(c[f_q1]) c[f_q2] over the n_selected bins freq_index[q] (int32, device; the host
 * applies the reference's band-pass and independent-frequency subsampling) of the coherency c64
 * [B][F][S][S] -> f32 [B][S][S] (NaN diagonal inherited from the coherency). */
int sc_phase_slope_index(const void* coherency_c64, int64_t B, int64_t F, int64_t S, const int* freq_index,
                         int n_selected, float* out, void* stream);

/* minimum_phase_decomposition(csm, tolerance, max_iterations) -- minimum_phase_decomposition.py:227-322
 * for 2x2 matrices: csm c128 [B][nfft][2][2] two-sided -> G c128 same shape.  Each b is an
 * independent unit of convergence (frozen at its first iterate with max|dG| < tol, :310-315).
 *  twiddle  c128 [nfft];  iters/flags int32 [B] (may be NULL). */
int sc_wilson2(const void* csm_c128, int64_t B, int nfft, double tolerance, int max_iterations,
               const void* twiddle, void* out_g_c128, int* out_iters, int* out_flags, void* workspace,
               int64_t workspace_bytes, void* stream);

/* pairwise_spectral_granger_prediction / subset_... -- connectivity.py:1161-1213, 2282-2340 with
 * _estimate_transfer_function (:1712-1748), _estimate_noise_covariance (:1679-1709),
 * _remove_instantaneous_causality (:1825-1848), _estimate_predictive_power (:1751-1779) fused.
 *  csm      c64 [B][F][S][S]; power f32 [B][F][S]
 *  F        nfft (two-sided spectrum) or nfft/2+1 with hermitian_half=1 (real time series:
 *           S(-f) = conj S(f), the from_multitaper path)
 *  pairs    int32 [n_pairs][2] or NULL for all i<j (then n_pairs is ignored)
 *           (all pairs with S % 4 == 0 and 16-byte aligned csm / power / out_gc take the grouped problem order: one
 *           staging pass per (window, row, 4 columns) instead of strided gathers per pair; results are identical)
 *  out_gc   f32 [B][nfft/2+1][S][S]; entries of processed pairs are overwritten ([i][j] =
 *           influence j -> i); the caller pre-fills the rest (NaN, connectivity.py:2310, 2336-2338)
 *  tail_extrapolation  0: run the reference iteration verbatim.  1 (hermitian_half only): once the
 *           update has collapsed onto the reference's geometric lag-0 mode (its plus-operator halves
 *           the lag-0 coefficient and then zeroes the lower triangle, mpd.py:132-138, so the lag-0
 *           off-diagonal residual halves per iteration), the remaining iterations are summed in
 *           closed form up to the iterate the reference stops at; out_iters reports that iterate.
 *  mixed_precision  1 (hermitian_half only, needs twiddle_c64 = c64 [nfft]): the first iterations, while the
 *           non-constant part of the update is still above 4e-4 of |G|, run in fp32 on the row-scaled problem; every later
 *           iteration, the stopping test and the Granger epilogue are fp64.  Moves the result by
 *           ~5e-7 relative (the fp32-born CSM already carries 3e-7).  0: fp64 throughout.
 *  out_iters/out_flags  int32 [n_pairs][B] or NULL
 *  out_exec_counters  uint64 [4] (device, accumulated with atomicAdd -- the caller zeroes it) or NULL: work the
 *           hermitian_half kernel actually EXECUTED, summed over all problems: [0] fp32-phase iterations,
 *           [1] fp64-phase iterations, [2] closed-form tail steps, [3] problems.  out_iters reports the
 *           reference-equivalent iteration count instead (= [0]+[1]+[2] summed).  Used for roofline accounting. */
int sc_granger_pairwise(const void* csm_c64, const float* power, int64_t B, int F, int nfft, int hermitian_half,
                        int64_t S, const int* pairs, int64_t n_pairs, double tolerance, int max_iterations,
                        int tail_extrapolation, int mixed_precision, const void* twiddle_c128,
                        const void* twiddle_c64, float* out_gc, int* out_iters, int* out_flags,
                        uint64_t* out_exec_counters, void* workspace, int64_t workspace_bytes, void* stream);
int64_t sc_wilson_workspace_bytes(int nfft);

/* minimum_phase_decomposition for S x S matrices, 1 <= S <= 1024 (minimum_phase_decomposition.py:227-322):
 * csm c128 [B][F][S][S] -> G c128 same shape; F = nfft (two-sided) or nfft/2+1 with hermitian_half=1
 * (real time series).  One iteration = a few stream-ordered kernels over the batch; convergence state
 * lives on the device (no host synchronisation), every b converges independently (:310-315).
 * S <= 32: one warp per (b, f) matrix in shared memory.  S > 32 (BASELINE config 5: 512 channels): blocked
 * in-place Gauss-Jordan inversion with partial pivoting + tiled c128 GEMMs in global memory.
 * workspace: sc_wilson_general_workspace_bytes(B, F, S). */
int sc_wilson(const void* csm_c128, int64_t B, int F, int nfft, int hermitian_half, int S, double tolerance,
              int max_iterations, const void* twiddle_c128, void* out_g_c128, int* out_iters, int* out_flags,
              void* workspace, int64_t workspace_bytes, void* stream);
int64_t sc_wilson_general_workspace_bytes(int64_t B, int F, int S);

/* MVAR quantities from the factor G (connectivity.py:567-588, 1679-1748):
 *  sc_mvar_lag0      H0 = Re ifft(G)[lag 0], f64 [B][S][S]
 *  sc_mvar_transfer  H = G (H0 + lambda I)^-1 for the first n_freq_out bins (c128 [B][n_freq_out][S][S]) and
 *                    the noise covariance H0 H0^T (f64 [B][S][S], may be NULL); lambda is the caller's
 *                    Tikhonov term 1e-12 * mean(H0^2) over all windows (:1742-1746)
 *  sc_mvar_inverse   A = (H + lambda I)^-1 per (b, f): the MVAR Fourier coefficients (:580-588)
 *  lambda_device (both): NULL, or a DEVICE scalar that overrides ``lambda`` -- the caller can then form the Tikhonov
 *                    term with a device reduction and needs no host synchronisation between Wilson and the measures
 * workspace (S > 32 only, else NULL/0): sc_mvar_workspace_bytes(B, 2, S) for sc_mvar_transfer,
 * sc_mvar_workspace_bytes(BF, 1, S) for sc_mvar_inverse. */
int sc_mvar_lag0(const void* g_c128, int64_t B, int F, int nfft, int hermitian_half, int S, double* out_h0,
                 void* stream);
int sc_mvar_transfer(const void* g_c128, const double* h0, double lambda, const double* lambda_device, int64_t B, int F,
                     int n_freq_out, int S, void* out_h_c128, double* out_sigma, void* workspace, int64_t workspace_bytes,
                     void* stream);
int sc_mvar_inverse(const void* h_c128, double lambda, const double* lambda_device, int64_t BF, int S, void* out_a_c128,
                    void* workspace, int64_t workspace_bytes, void* stream);
int64_t sc_mvar_workspace_bytes(int64_t B, int F, int S);

/* directed_transfer_function (0), directed_coherence (1), partial_directed_coherence (2),
 * generalized_partial_directed_coherence (3), direct_directed_transfer_function (4)
 * (connectivity.py:1237-1426): f32 [B][F][S][S] from H and/or A (c128 [B][F][S][S]) and the noise
 * covariance (f64 [B][S][S]); scratch: B*S doubles (measure 4 only) for S <= 32, B*S + 2*B*F*S doubles
 * (every measure) for S > 32. */
int sc_mvar_measure(int measure, const void* h_c128, const void* a_c128, const double* sigma, int64_t B, int F, int S,
                    double* scratch, float* out, void* stream);

/* canonical_coherence(group_labels) -- connectivity.py:745-820, 1979-2032: squared canonical coherence between
 * signal groups from the expected CSM (c64 [B][F][S][S]): per (b, f, group pair) two Cholesky factorisations, two
 * triangular solves and the top eigenvalue of M M^H.  group_index int32 [S] lists the signals sorted by group,
 * group_offsets int32 [n_groups+1] delimits them (both device); groups hold at most 64 signals.
 * out f32 [B][F][n_groups][n_groups]; only off-diagonal entries are written (the caller pre-fills NaN, :797).
 * out_flags int32 [B*F] or NULL: SC_FLAG_NOT_SPD when a group's block is not positive definite (fewer
 * observations than signals). */
int sc_canonical_coherence(const void* csm_c64, int64_t B, int F, int S, const int* group_index,
                           const int* group_offsets, int n_groups, int max_group_size, float* out, int* out_flags,
                           void* stream);

/* global_coherence(max_rank=1) -- connectivity.py:822-895, 2245-2279: largest eigenvalue of each CSM
 * (c64 [BF][S][S]) = top singular value^2 / n_observations, and its unit eigenvector (c64 [BF][S], defined up
 * to a phase like the reference's SVD output).  S <= 64: one CTA per matrix in shared memory, workspace NULL/0.
 * S > 64 (up to 1024): repeated squaring with the tiled c128 GEMM, workspace
 * sc_global_coherence_workspace_bytes(BF, S) -- callers chunk BF to bound it. */
int sc_global_coherence(const void* csm_c64, int64_t BF, int S, float* out_value, void* out_vector_c64,
                        void* workspace, int64_t workspace_bytes, void* stream);
int64_t sc_global_coherence_workspace_bytes(int64_t BF, int S);

/* global_coherence(max_rank > 1) -- connectivity.py:2245-2279 keeps the max_rank largest singular values: the
 * deflation C <- C - lambda v v^H (c64 [BF][S][S], in place) removes the eigenpair sc_global_coherence just returned
 * (value f32 [BF], vector c64 [BF][S]), so that the next call on the same buffer yields the next eigenpair. */
int sc_hermitian_deflate(void* csm_c64, int64_t BF, int S, const float* value, const void* vector_c64, void* stream);

/* Trial-sharded mode (partition B): fused reduce-scatter + epilogue over NVLink peer memory.
 * Every rank holds partial expectation sums in a buffer its peers can address (e.g. torch symmetric memory);
 * ``peers`` is a HOST array of ``world`` device pointers, rank order, one per rank's buffer.  The caller orders
 * "all ranks have written" before the launch and "all ranks have read" before a buffer is rewritten (device-side
 * barrier of the symmetric-memory handle on the same stream).
 *  sc_peer_reduce_csm  buffers hold c64 matrices [..][S][S]; matrices mat0 .. mat0+n_mat-1 are summed over the peers
 *                      in rank order (deterministic) into out_csm (c64 [n_mat][S][S], may be NULL); out_power (f32
 *                      [n_mat][S], may be NULL) receives the real diagonal (connectivity.py:441-445); measure = -1 or
 *                      SC_M_COHERENCY .. SC_M_IMAG_COHERENCE writes that epilogue (connectivity.py:632-743) into
 *                      out_measure in the same pass.  S must be even.
 *  sc_peer_reduce      plain sum of n_bytes (multiple of 16) of float32 data at offset_bytes of every peer buffer. */
int sc_peer_reduce_csm(const void* const* peers, int world, int64_t mat0, int64_t n_mat, int S, void* out_csm_c64,
                       float* out_power, int measure, void* out_measure, void* stream);
int sc_peer_reduce(const void* const* peers, int world, int64_t offset_bytes, int64_t n_bytes, void* out, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* SC_B200_H */
