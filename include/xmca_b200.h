/* xmca_b200 -- C ABI of the B200-native MCA engine (libxmca_b200.so).
 *
 * The reference (nicrie/xmca) has NO FFI: its hot path is numpy/LAPACK calls
 * inside xmca/array.py and xmca/tools/rotation.py.  Each entry point below
 * replaces one of those library-call seams; the comment above it cites the
 * reference lines (relative to the reference repo root).  INTEGRATION.md shows
 * the ctypes stub a maintainer of the reference would add at each seam.
 *
 * Conventions
 *   - every pointer named d_* is a DEVICE pointer owned by the caller (the
 *     Python host passes torch.cuda tensor data_ptr()s); the library never
 *     keeps a pointer after the call returns and never allocates persistent
 *     device memory.  Scratch space is caller-supplied: ask the matching
 *     *_workspace_bytes() function first.
 *   - `stream` is a cudaStream_t passed as void*; all work is enqueued on it (or on a stream
 *     forked from and joined back onto it inside the call: xmca_cholesky).
 *     Calls that need a device->host decision (convergence tests) synchronise
 *     that stream internally and say so.
 *   - matrices are dense, real.  Complex fields are handled by the host as
 *     planar (re, im) pairs / real embeddings.
 *   - return value: 0 ok, 1 bad argument, 2 CUDA error, 3 not converged,
 *     4 numerical failure.  xmca_last_error() gives the message (thread local).
 */
#ifndef XMCA_B200_H
#define XMCA_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum { XMCA_OK = 0, XMCA_BAD_ARG = 1, XMCA_CUDA_ERROR = 2, XMCA_NOT_CONVERGED = 3, XMCA_NUMERIC = 4 };
enum { XMCA_F32 = 0, XMCA_F64 = 1 };

const char* xmca_last_error(void);
int xmca_version(void);
/* number of kernels this library has launched since load (bench.py's gpu_launches) */
long long xmca_launch_count(void);

/* ---- general dense product (SIMT FP64/FP32 cores) ------------------------
 * D[M,N] (row-major, ldd) = alpha * opA(A) * opB(B)  (+ D if accumulate)
 *   a_kmajor=1: A is stored M x K row-major (lda >= K);  0: K x M row-major (lda >= M)
 *   b_kmajor=1: B is stored N x K row-major (ldb >= K);  0: K x N row-major (ldb >= N)
 * acc_dtype selects the accumulator (XMCA_F64 for the parity-critical paths).
 * Replaces the `@` products of array.py:556-566 (kernel), :584 (back-projection),
 * :640 (rotated EOFs), :667-669 (PCs) and rotation.py:128-147 (Promax fit).
 * split_k > 1 needs workspace of xmca_gemm_workspace_bytes(). */
size_t xmca_gemm_workspace_bytes(int64_t M, int64_t N, int split_k, int acc_dtype);
int xmca_gemm(int a_kmajor, int b_kmajor, int64_t M, int64_t N, int64_t K, double alpha,
              const void* d_A, int a_dtype, int64_t lda,
              const void* d_B, int b_dtype, int64_t ldb,
              void* d_D, int d_dtype, int64_t ldd, int accumulate,
              int acc_dtype, int split_k, void* d_workspace, size_t workspace_bytes,
              void* stream);

/* The same product when the operands / the result have structure (M == N, split_k = 1):
 *   XMCA_GEMM_SYMMETRIC : the result is symmetric (X X^T, L^T G L, V W^T + W V^T with D symmetric
 *                         when accumulating): tiles above the diagonal are mirrored, not computed
 *   XMCA_GEMM_A_LOWER_T : opA = L^T with L lower triangular, stored K x M (a_kmajor = 0)
 *   XMCA_GEMM_B_LOWER   : opB = L lower triangular, stored K x N (b_kmajor = 0)
 *   XMCA_GEMM_LOWER_ONLY: only the 128 x 128 tiles on or below the diagonal are computed / written
 *                         (trailing update of the Cholesky factorisation)
 * Used for the T x T Gram matrices and S = L_B^T G_A L_B of the tridiagonal route. */
enum { XMCA_GEMM_SYMMETRIC = 1, XMCA_GEMM_A_LOWER_T = 2, XMCA_GEMM_B_LOWER = 4, XMCA_GEMM_LOWER_ONLY = 8 };
int xmca_gemm_ex(int a_kmajor, int b_kmajor, int64_t M, int64_t N, int64_t K, double alpha,
                 const void* d_A, int a_dtype, int64_t lda,
                 const void* d_B, int b_dtype, int64_t ldb,
                 void* d_D, int d_dtype, int64_t ldd, int accumulate,
                 int acc_dtype, int split_k, void* d_workspace, size_t workspace_bytes,
                 int flags, void* stream);

/* ---- cross-covariance on the tensor cores (tcgen05 + TMA, 3xTF32) ---------
 * Replaces the field SVDs + kernel product of array.py:552-566: forms
 * C = A^T B * alpha directly (S1 x S2, fp32) from fp32 fields.
 * Step 1: xmca_split_tf32 writes the (hi, lo) TF32 planes of X or X^T,
 *         K-major, row pitch ldo (multiple of 4 floats).
 * Step 2: xmca_tc_gemm_nt multiplies two plane pairs:
 *         D[M,N] = alpha * sum_k A[m,k] B[n,k]  with hi*hi + hi*lo + lo*hi.
 * frob2 (optional, device double*) accumulates sum(D^2)
 * (= total_squared_covariance, array.py:596). */
int xmca_split_tf32(const void* d_X, int x_dtype, int64_t rows, int64_t cols, int64_t ldx,
                    int transpose, float* d_hi, float* d_lo, int64_t ldo, void* stream);
int xmca_tc_gemm_nt(int64_t M, int64_t N, int64_t K, float alpha,
                    const float* d_Ahi, const float* d_Alo, int64_t lda,
                    const float* d_Bhi, const float* d_Blo, int64_t ldb,
                    float* d_D, int64_t ldd, double* d_frob2, void* stream);

/* fp64-output variant for long contractions (Gram matrices G = X X^T of fp32 fields, the T x T
 * matrices of the tridiagonal route): large and small split products in separate TMEM accumulators,
 * chunks of 64 k drained into fp64 register sums.  symmetric != 0 (M == N, same operand): only the
 * tiles on or below the diagonal are computed, the rest mirrored. */
int xmca_tc_gemm_nt_f64(int64_t M, int64_t N, int64_t K, double alpha,
                        const float* d_Ahi, const float* d_Alo, int64_t lda,
                        const float* d_Bhi, const float* d_Blo, int64_t ldb,
                        double* d_D, int64_t ldd, int symmetric, void* stream);

/* ---- SVD / symmetric eigen-decomposition: blocked one-sided Jacobi --------
 * Replaces np.linalg.svd of array.py:479 and :570 (and the p x p SVD of
 * rotation.py:59 when called with small n).
 * d_Kc : m x n matrix stored COLUMN-major (column j at d_Kc + j*ldk), fp64,
 *        n_pad = xmca_jacobi_padded_cols(n) columns allocated (extra columns
 *        must be zero).  On return column j holds u_j * sigma_j.
 * d_Jc : optional n_pad x n_pad column-major fp64 (ldj >= n_pad), overwritten
 *        with the accumulated right rotations (right singular vectors).
 * d_sigma : n_pad doubles, column norms on return (NOT sorted; host sorts).
 * tol : sweeps stop once the largest cosine |x_i.x_j| / (|x_i||x_j|) seen BEFORE a sweep
 *       is <= tol (<= 0: 1e-11).  Columns below 1e-11 of the largest norm count as zero.
 * Synchronises `stream` once per sweep to read the convergence measure. */
int64_t xmca_jacobi_padded_cols(int64_t n);
size_t xmca_jacobi_workspace_bytes(int64_t m, int64_t n);
int xmca_jacobi_svd(int64_t m, int64_t n, double* d_Kc, int64_t ldk,
                    double* d_Jc, int64_t ldj, double* d_sigma,
                    int max_sweeps, double tol, int* sweeps_out, double* offnorm_out,
                    void* d_workspace, size_t workspace_bytes, void* stream);

/* ---- Cholesky factorisation + triangular solve (fp64) ---------------------
 * The "Cholesky-QR" route of the engine for T < S: with G_X = X X^T = L_X L_X^T
 * the singular values of C = A^T B are those of the T x T matrix L_A^T L_B, so
 * ONE Jacobi SVD replaces the three LAPACK SVDs of array.py:479 (x2) and :570,
 * and V_X = X^T (L_X^{-T} P) replaces the back-projection of array.py:584.
 * xmca_cholesky: d_A (n x n row-major, lda) is overwritten by its lower Cholesky
 *   factor (strict upper part zeroed); d_invdiag receives the inverses of the
 *   64 x 64 diagonal blocks (xmca_cholesky_invdiag_bytes).  Returns XMCA_NUMERIC
 *   (info_out = 1 + failing column) if a pivot is not above min_pivot (>= 0): the
 *   matrix is not numerically positive definite.
 *   Synchronises `stream` once at the end to read that flag.  Look-ahead: the chain diagonal block ->
 *   panel -> next block column runs on an internal high-priority stream that is forked from and joined
 *   back onto `stream` inside the call (the bulk trailing updates stay on `stream`).
 * xmca_trsm_lt: solves L^T W = R in place (R: n x nrhs row-major, ldr). */
size_t xmca_cholesky_workspace_bytes(int64_t n);
size_t xmca_cholesky_invdiag_bytes(int64_t n);
int xmca_cholesky(int64_t n, double* d_A, int64_t lda, double* d_invdiag, double min_pivot,
                  int* info_out, void* d_workspace, size_t workspace_bytes, void* stream);
size_t xmca_trsm_workspace_bytes(int64_t n, int64_t nrhs);
int xmca_trsm_lt(int64_t n, int64_t nrhs, const double* d_L, int64_t ldl, const double* d_invdiag,
                 double* d_R, int64_t ldr, void* d_workspace, size_t workspace_bytes, void* stream);

/* ---- symmetric eigen-solver by Householder tridiagonalisation (fp64) ------
 * The engine's full-spectrum route for large problems: sigma(C)^2 are the
 * eigenvalues of ONE T x T symmetric matrix S (engine.py), so the three LAPACK
 * SVDs of array.py:479 (x2) and :570 become  S = Q T Q^T  (4/3 n^3 flops, one
 * streaming pass over the trailing matrix per column), bisection for all
 * eigenvalues, and inverse iteration + back-transformation for the vectors
 * that are actually asked for (array.py:584 computes all of them).
 * xmca_sytrd: d_A (n x n, row-major, symmetric, BOTH triangles stored) is
 *   destroyed; on return d_d (n) / d_e (n-1) hold the tridiagonal, row c of d_A
 *   (columns c+1..n-1) holds Householder vector c (leading 1 stored) and d_tau
 *   (n) its scalar:  Q = H(0) H(1) ... H(n-2),  H(c) = I - tau_c v_c v_c^T.
 * xmca_sytrd_batched: batch = 1 or 2 problems of the same size n in ONE sequence of launches, each on half
 *   of the SMs with its own grid barrier: while one streams its trailing matrix the other is in its
 *   latency-bound phases (column update, reflector, reductions).  This is what the independent surrogate
 *   runs of rule_n (array.py:1753-1765) use.  Problem 1 lives stride_a doubles behind d_A and stride_v
 *   doubles behind d_d / d_e / d_tau; d_workspace: batch * xmca_sytrd_workspace_bytes(n).
 * xmca_stebz: all n eigenvalues, DESCENDING, into d_w (d_scratch: 2 n + 8 doubles, 16-byte aligned).
 *   Synchronises `stream` once (Gershgorin bounds come back to the host).
 * xmca_stein: eigenvectors of the tridiagonal for the k eigenvalues d_lambda
 *   (descending), written as ROWS of d_Z (k x n, ldz).  d_cluster_start
 *   (n_clusters + 1 ints, device): members of a cluster are orthogonalised
 *   against each other (modified Gram-Schmidt), clusters run in parallel.
 *   tnorm = max |eigenvalue| (scale of the pivot perturbation).
 * xmca_ormtr: rows of d_Z <- Q * row  (eigenvectors of the original matrix). */
int64_t xmca_sytrd_max_n(void);
size_t xmca_sytrd_workspace_bytes(int64_t n);
int xmca_sytrd(int64_t n, double* d_A, int64_t lda, double* d_d, double* d_e, double* d_tau,
               void* d_workspace, size_t workspace_bytes, void* stream);
int xmca_sytrd_batched(int64_t n, int batch, double* d_A, int64_t lda, int64_t stride_a, double* d_d,
                       double* d_e, double* d_tau, int64_t stride_v, void* d_workspace,
                       size_t workspace_bytes, void* stream);
int xmca_stebz(int64_t n, const double* d_d, const double* d_e, double* d_w, double* d_scratch,
               void* stream);
size_t xmca_stein_workspace_bytes(int64_t n, int64_t n_clusters);
int xmca_stein(int64_t n, const double* d_d, const double* d_e, int64_t k, const double* d_lambda,
               const int* d_cluster_start, int64_t n_clusters, double tnorm, int iterations,
               double* d_Z, int64_t ldz, void* d_workspace, size_t workspace_bytes, void* stream);
int xmca_ormtr(int64_t n, const double* d_A, int64_t lda, const double* d_tau, int64_t k,
               double* d_Z, int64_t ldz, void* stream);

/* ---- two-stage symmetric tridiagonalisation (fp64; csrc/sbr.cu + csrc/sbtrd.cu) --------------
 * Same seam as xmca_sytrd (np.linalg.svd of array.py:479 / :570), compute-bound instead of HBM-bound:
 *   stage 1  dense -> band (bandwidth 64): panels by shifted CholeskyQR3 + Householder reconstruction,
 *            trailing matrix  A22 <- A22 - X Y^T - Y X^T  with Z = A22 Y and the rank-128 update on the
 *            fp64 DMMA pipe (4/3 n^3 flops, all GEMM shaped);
 *   stage 2  band -> tridiagonal by bulge chasing on the L2-resident band (one persistent CTA per
 *            sweep, task windows in registers, release/acquire progress flags between sweeps).
 * xmca_sytrd2: d_A (n x n row-major, symmetric, BOTH triangles) is destroyed.  On return d_d (n) / d_e (n - 1)
 *   hold the tridiagonal; d_A holds the stage-1 block reflectors Y_p (m x 64, dense, in the place of panel p,
 *   below the band) and -- when want_vectors != 0 -- the stage-2 reflectors in its strict upper triangle
 *   (row j: sweep j, leading entry of every reflector replaced by its tau); d_tfac (xmca_sytrd2_tfac_bytes)
 *   receives the 64 x 64 triangular factors T_p of  W_p = I - Y_p T_p Y_p^T.
 *   Returns XMCA_NUMERIC if a panel factorisation broke down (non-finite input / exactly rank-deficient
 *   panel): the caller falls back to xmca_sytrd.  Synchronises `stream` once at the end to read that flag -- unless
 *   bit 3 (value 8) of want_vectors is set: then the call only enqueues, and the caller reads the int32 flag at byte
 *   offset xmca_sytrd2_info_offset(n) of d_workspace after its own synchronisation (non-zero = breakdown).  Two such
 *   calls on two streams overlap on the device (the paired surrogate runs of rule_n, array.py:1753-1765).
 * xmca_ormtr2: rows of d_Z (k x n) <- Q row with Q = Q1 Q2 from xmca_sytrd2 (eigenvectors of the tridiagonal ->
 *   eigenvectors of the original matrix; array.py:584 for the modes that are asked for).  Stage 2: one CTA per
 *   vector (kept in shared memory), reflector data prefetched one sweep ahead; stage 1: three products per panel
 *   over all vectors (workspace xmca_ormtr2_workspace_bytes; without it a per-vector kernel is used). */
size_t xmca_sytrd2_workspace_bytes(int64_t n);
size_t xmca_sytrd2_tfac_bytes(int64_t n);
size_t xmca_sytrd2_info_offset(int64_t n);
int xmca_sytrd2(int64_t n, double* d_A, int64_t lda, double* d_d, double* d_e, double* d_tfac,
                int want_vectors, void* d_workspace, size_t workspace_bytes, void* stream);
size_t xmca_ormtr2_workspace_bytes(int64_t n, int64_t k);
int xmca_ormtr2(int64_t n, const double* d_A, int64_t lda, const double* d_tfac, int64_t k,
                double* d_Z, int64_t ldz, void* d_workspace, size_t workspace_bytes, void* stream);

/* ---- analytic signal (Hilbert transform along time) ----------------------
 * Replaces scipy.signal.hilbert(field, axis=0) of array.py:464 by two linear operators that
 * are applied to the T x S field as GEMMs (xmca_tc_gemm_nt for fp32 fields, xmca_gemm for fp64):
 * xmca_hilbert_matrix: H (T x T, circulant) with  imag(analytic signal) = H x;  d_taps: T doubles.
 * xmca_dft_matrix: F (xmca_dft_rows(T) = 2 floor(T/2) rows x T): stacked [Re; Im] rows of
 *   diag(w) DFT / sqrt(T) for the positive frequencies f = 1..floor(T/2) (w = 2, Nyquist 1): the
 *   one-sided spectrum Z^ = F x satisfies  Z_A^H Z_B = Z^_A^H Z^_B, so solve() runs on Z^.
 * xmca_embed_complex: real embedding [[Zr, -Zi], [Zi, Zr]] of the stacked [Zr; Zi] (2 rows_half x cols). */
int xmca_hilbert_matrix(int64_t T, void* d_H, int h_dtype, int64_t ldh, double* d_taps, void* stream);
/* rows x cols block at (row0, col0) of the length-N operator (d_taps: N doubles): the pieces of the 3T-long
 * transform that the fore/back-cast extension of array.py:378-472 (extend='exp') needs. */
int xmca_hilbert_block(int64_t N, int64_t rows, int64_t cols, int64_t row0, int64_t col0,
                       void* d_H, int h_dtype, int64_t ldh, double* d_taps, void* stream);
int64_t xmca_dft_rows(int64_t T);
int xmca_dft_matrix(int64_t T, void* d_F, int f_dtype, int64_t ldf, void* stream);
int xmca_embed_complex(const void* d_Z, int z_dtype, int64_t ldz, int64_t rows_half, int64_t cols,
                       void* d_E, int e_dtype, int64_t lde, void* stream);

/* ---- element-wise / layout helpers ---------------------------------------*/
/* Y[r,c] = X[r,c] * (col_scale ? col_scale[c] : 1) * (row_scale ? row_scale[r] : 1);
 * row-major, dtypes may differ (conversion kernel).  array.py:553, :640, :667. */
int xmca_scale_copy(const void* d_X, int x_dtype, int64_t ldx, void* d_Y, int y_dtype, int64_t ldy,
                    int64_t rows, int64_t cols, const double* d_col_scale, const double* d_row_scale,
                    void* stream);
/* Y (cols x rows, row-major) = X^T, with dtype conversion. */
int xmca_transpose(const void* d_X, int x_dtype, int64_t rows, int64_t cols, int64_t ldx,
                   void* d_Y, int y_dtype, int64_t ldy, void* stream);
/* out[c] = sum_r X[r,c]^2 for r in [row0,row1)  (fp64 out); array.py:826-830. */
int xmca_col_sumsq(const void* d_X, int x_dtype, int64_t ldx, int64_t row0, int64_t row1,
                   int64_t cols, double* d_out, void* stream);
/* subtract the column mean in place (array.py:199-207), mean returned in d_mean (fp64). */
int xmca_center_columns(void* d_X, int x_dtype, int64_t rows, int64_t cols, int64_t ldx,
                        double* d_mean, void* stream);
/* Constructor pre-processing on the device (array.py:191-240, tools/array.py:26-73):
 * per column mean / std (ddof 0) / "contains a NaN" flag, per row "has a valid value" flag. */
int xmca_field_stats(const void* d_X, int x_dtype, int64_t rows, int64_t cols, int64_t ldx,
                     double* d_mean, double* d_std, int* d_col_nan, int* d_row_valid, void* stream);
/* Y[r, j] = X[r, idx[j]] - mean[idx[j]]: drop the NaN columns and centre (array.py:199-228),
 * subtraction in the precision of X like the reference. */
int xmca_compact_center(const void* d_X, int x_dtype, int64_t rows, int64_t ldx,
                        const int64_t* d_idx, int64_t n_keep, const double* d_mean,
                        void* d_Y, int y_dtype, int64_t ldy, void* stream);
/* fill X (rows x cols) with N(0,1) from Philox4x32-10, counter = element index,
 * key = (seed, stream_id): result independent of launch geometry. array.py:1756. */
int xmca_fill_normal(void* d_X, int x_dtype, int64_t rows, int64_t cols, int64_t ldx,
                     uint64_t seed, uint64_t stream_id, void* stream);

/* Y[i,:] = X[idx[i],:] * row_scale[i]  (mode sorting + normalisation after the Jacobi SVD;
 * array.py:592 argsort / :584). idx: device int64. */
int xmca_gather_rows(const void* d_X, int x_dtype, int64_t ldx, const int64_t* d_idx, int64_t n_out,
                     int64_t cols, const double* d_row_scale, void* d_Y, int y_dtype, int64_t ldy,
                     void* stream);
/* out[r] = sum_c X[r,c]^2 (communalities, rotation.py:46 / :115). */
int xmca_row_sumsq(const void* d_X, int x_dtype, int64_t ldx, int64_t rows, int64_t cols,
                   double* d_out, void* stream);
/* out[c] = max_r |X[r,c] * row_scale[r]|  (rotation.py:121). */
int xmca_col_absmax(const void* d_X, int x_dtype, int64_t ldx, int64_t rows, int64_t cols,
                    const double* d_row_scale, double* d_out, void* stream);
/* Promax target (rotation.py:115-124): Xout = X * row_scale[r]; Pout = Xn |Xn|^(power-1),
 * Xn = Xout / colmax[c]. */
int xmca_promax_target(const double* d_X, int64_t ldx, int64_t rows, int64_t cols,
                       const double* d_row_scale, const double* d_colmax, double power,
                       double* d_Xout, double* d_Pout, int64_t ldo, void* stream);

/* complex (planar fp64) versions of the two Promax helpers above (rotation.py:115-124, complex dtype) */
int xmca_col_absmax_complex(const double* d_Xr, const double* d_Xi, int64_t ldx, int64_t rows, int64_t cols,
                            const double* d_row_scale, double* d_out, void* stream);
int xmca_promax_target_complex(const double* d_Xr, const double* d_Xi, int64_t ldx, int64_t rows, int64_t cols,
                               const double* d_row_scale, const double* d_colmax, double power,
                               double* d_Xor, double* d_Xoi, double* d_Por, double* d_Poi, int64_t ldo,
                               void* stream);

/* ---- fused Varimax / Promax rotation --------------------------------------
 * Replaces xmca/tools/rotation.py:15-78 (varimax) driven from array.py:823.
 * d_L : n x p loadings, row-major (ld = ldl), fp32 or fp64.  Real case.
 * Runs the Kaiser-normalised fixed point with a persistent cooperative
 * kernel: per iteration one streaming pass over the normalised loadings,
 * p x p polar factor on one CTA, device-side convergence test
 * |d - d_old| / d < tol (rotation.py:62).  All p x p state is fp64.
 * Outputs: d_R (p x p row-major fp64), d_B (n x p row-major fp64: rotated
 * loadings, de-normalised, rotation.py:74-77), iterations.
 * d_out: 16 doubles of statistics: [0] iterations, [1] converged, [2] sum of singular values, [3] Jacobi sweeps,
 * [4..9] phase clocks of CTA 0, [10] rotations applied, [11] / [12] pairs seen with a cosine above 1e-4 / 1e-3.
 * Returns XMCA_NOT_CONVERGED after max_iter (rotation.py:66-71). */
size_t xmca_varimax_workspace_bytes(int64_t n, int p);
int xmca_varimax(const void* d_L, int l_dtype, int64_t n, int p, int64_t ldl,
                 double gamma, int max_iter, double tol,
                 double* d_B, int64_t ldb, double* d_R, int* iterations_out, double* d_out,
                 void* d_workspace, size_t workspace_bytes, void* stream);

/* Complex loadings (complex MCA): planar (re, im) n x p inputs in storage precision, p <= 32;
 * outputs planar fp64 B (rotated loadings) and R (p x p unitary rotation).  Same fixed point
 * evaluated with complex arithmetic (rotation.py:46-62 with complex dtype: |b|^2 b, A^H, polar
 * factor of a complex p x p matrix). */
size_t xmca_varimax_complex_workspace_bytes(int64_t n, int p);
int xmca_varimax_complex(const void* d_Lr, const void* d_Li, int l_dtype, int64_t n, int p, int64_t ldl,
                         double gamma, int max_iter, double tol,
                         double* d_Br, double* d_Bi, int64_t ldb, double* d_Rr, double* d_Ri,
                         int* iterations_out, double* d_out,
                         void* d_workspace, size_t workspace_bytes, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* XMCA_B200_H */
