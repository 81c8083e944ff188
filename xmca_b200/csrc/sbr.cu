// Stage 1 of the two-stage symmetric tridiagonalisation (fp64): dense symmetric S -> symmetric band matrix of
// bandwidth 64 by blocked orthogonal similarity transformations, every O(n^3) part GEMM shaped (fp64 DMMA pipe),
// plus the drivers of the whole two-stage solver (stage 2: sbtrd.cu).
//
// Replaces np.linalg.svd of array.py:479 / :570 (engine.py: sigma(C)^2 = eigenvalues of ONE symmetric matrix)
// -- the one-stage xmca_sytrd (tridiag.cu) streams the trailing matrix once per column and is bound by HBM and
// per-column latency; here the 4/3 n^3 flops are products with K = 64 / 128.
//
// Panel k (columns kb .. kb + 63, rows r0 = kb + 64 .. n - 1, m rows):
//   P = Q R by shifted CholeskyQR3 (Fukaya et al. 2020): three passes of Gram matrix (deterministic two-level
//       sum) -> 64 x 64 Cholesky (one CTA) -> triangular solve per row; the tiny shift of the first pass bounds the
//       condition number the second pass sees, so no pass can break down on a numerically full-rank panel;
//   compact WY form of an orthogonal W = I - Y T Y^T with W [I; 0] = Q diag(s) by Householder RECONSTRUCTION
//       (Ballard et al. 2015): LU of [I; 0] - Q diag(s) with the signs chosen so that every pivot is >= 1;
//   two-sided update of the trailing matrix  A22 <- W^T A22 W = A22 - X Y^T - Y X^T,
//       Z = A22 Y (DMMA, split-K), X = (Z - 1/2 Y T^T (Y^T Z)) T, rank-128 update on xmca_gemm_ex (DMMA, lower tiles
//       computed, upper mirrored).
//   R diag(s) goes to the band array; Y stays in the panel's place in S (dense m x 64), T in the caller's `tfac`.
// The last panel (m < 64 rows) is done by one CTA with plain Householder reflectors.
#include "common.cuh"
#include "small64.cuh"
#include <math.h>
#include <stdlib.h>
#include <vector>

namespace xmca {

// (PB = 64: panel width = bandwidth, PLD = 65: shared-memory pitch -- small64.cuh)
constexpr int BLD = 2 * PB;         // doubles per band column (same as sbtrd.cu)
constexpr int PT = 256;             // threads of the chunk kernels
constexpr int NRED = 256;           // CTAs of the reduction kernels (16 elements each, 16 threads per element)

int sb_chase(double* AB, int n, int* counters, double* V2, int64_t ldv, double* d, double* e, cudaStream_t st,
             long long* prof = nullptr, void* ll = nullptr);
size_t sb_chase_ll_bytes(int n);
int sb_apply_q2(const double* V2, int64_t ldv, int n, double* Z, int64_t ldz, int kvec, cudaStream_t st);

__device__ __forceinline__ void bar64() { asm volatile("bar.sync 1, 64;" ::: "memory"); }


// ------------------------------------------------------------------ small dense helpers (one CTA)
// 64 x 64 (pitch PLD) from / to global (pitch 64), all threads of the CTA
__device__ __forceinline__ void load64(double (*s)[PLD], const double* __restrict__ g, int tid, int nthreads) {
  for (int e = tid; e < PB * PB; e += nthreads) s[e >> 6][e & 63] = g[e];
}

// X <- X L^-T (rows of X; L lower in Ls, Dinv = inverses of its diagonal blocks).  Thread (r = tid >> 2, kq = tid & 3):
// the four threads of a row compute the diagonal-block step redundantly and share the trailing blocks.
__device__ __forceinline__ void rows_trsm_lt(double (*Xs)[PLD], const double (*Ls)[PLD], const double* Dinv, int tid) {
  const int r = tid >> 2, kq = tid & 3;
  for (int cb = 0; cb < NB8; ++cb) {
    const int j0 = 8 * cb;
    double x[8], q[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) x[k] = Xs[r][j0 + k];
#pragma unroll
    for (int jj = 0; jj < 8; ++jj) {
      double sacc = 0.0;
#pragma unroll
      for (int k = 0; k <= jj; ++k) sacc = fma(x[k], Dinv[cb * 64 + jj * 8 + k], sacc);
      q[jj] = sacc;
    }
    __syncwarp();
    if (kq == 0) {
#pragma unroll
      for (int jj = 0; jj < 8; ++jj) Xs[r][j0 + jj] = q[jj];
    }
    for (int tb = cb + 1 + kq; tb < NB8; tb += 4) {
      const int t0 = 8 * tb;
      double y[8];
#pragma unroll
      for (int kk = 0; kk < 8; ++kk) y[kk] = Xs[r][t0 + kk];
#pragma unroll
      for (int kk = 0; kk < 8; ++kk)
#pragma unroll
        for (int jj = 0; jj < 8; ++jj) y[kk] = fma(-q[jj], Ls[t0 + kk][j0 + jj], y[kk]);
#pragma unroll
      for (int kk = 0; kk < 8; ++kk) Xs[r][t0 + kk] = y[kk];
    }
    __syncwarp();
  }
}
// X <- X U^-1 (U upper in Us, Uinv = inverses of its diagonal blocks)
__device__ __forceinline__ void rows_trsm_u(double (*Xs)[PLD], const double (*Us)[PLD], const double* Uinv, int tid) {
  const int r = tid >> 2, kq = tid & 3;
  for (int cb = 0; cb < NB8; ++cb) {
    const int j0 = 8 * cb;
    double x[8], q[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) x[k] = Xs[r][j0 + k];
#pragma unroll
    for (int jj = 0; jj < 8; ++jj) {
      double sacc = 0.0;
#pragma unroll
      for (int k = 0; k <= jj; ++k) sacc = fma(x[k], Uinv[cb * 64 + k * 8 + jj], sacc);
      q[jj] = sacc;
    }
    __syncwarp();
    if (kq == 0) {
#pragma unroll
      for (int jj = 0; jj < 8; ++jj) Xs[r][j0 + jj] = q[jj];
    }
    for (int tb = cb + 1 + kq; tb < NB8; tb += 4) {
      const int t0 = 8 * tb;
      double y[8];
#pragma unroll
      for (int kk = 0; kk < 8; ++kk) y[kk] = Xs[r][t0 + kk];
#pragma unroll
      for (int jj = 0; jj < 8; ++jj)
#pragma unroll
        for (int kk = 0; kk < 8; ++kk) y[kk] = fma(-q[jj], Us[j0 + jj][t0 + kk], y[kk]);
#pragma unroll
      for (int kk = 0; kk < 8; ++kk) Xs[r][t0 + kk] = y[kk];
    }
    __syncwarp();
  }
}

// LU without pivoting of  A diag(s) + I  with the signs s_j chosen on the fly so that every pivot is >= 1
// (Householder reconstruction).  In place: strict lower part = L (unit diagonal implied), upper part = U;
// Linv / Uinv: inverses of the diagonal blocks of L / U; sg: the 64 signs.
__device__ __forceinline__ void lu64_sign_blocked(double (*A)[PLD], double* Linv, double* Uinv, double* sg, int tid) {
  const int ti = tid >> 4, tk = tid & 15;
  for (int cb = 0; cb < NB8; ++cb) {
    const int j0 = 8 * cb;
    double a[8][8], li[8][8], ui[8][8], sv[8];
#pragma unroll
    for (int ii = 0; ii < 8; ++ii)
#pragma unroll
      for (int jj = 0; jj < 8; ++jj) a[ii][jj] = A[j0 + ii][j0 + jj];
#pragma unroll
    for (int jj = 0; jj < 8; ++jj) {
      const double s = a[jj][jj] >= 0.0 ? 1.0 : -1.0;
      sv[jj] = s;
#pragma unroll
      for (int ii = 0; ii < 8; ++ii) a[ii][jj] *= s;                 // column jj scaled in every row of the block
      a[jj][jj] += 1.0;
      const double rp = fast_rcp(a[jj][jj]);
      ui[jj][jj] = rp;
#pragma unroll
      for (int ii = jj + 1; ii < 8; ++ii) a[ii][jj] *= rp;
#pragma unroll
      for (int ii = jj + 1; ii < 8; ++ii)
#pragma unroll
        for (int kk = jj + 1; kk < 8; ++kk) a[ii][kk] = fma(-a[ii][jj], a[jj][kk], a[ii][kk]);
    }
    // Li = inverse of the unit lower block, Ui = inverse of the upper block
#pragma unroll
    for (int jj = 0; jj < 8; ++jj) {
      li[jj][jj] = 1.0;
#pragma unroll
      for (int ii = jj + 1; ii < 8; ++ii) {
        double sacc = 0.0;
#pragma unroll
        for (int k = jj; k < ii; ++k) sacc = fma(a[ii][k], li[k][jj], sacc);
        li[ii][jj] = -sacc;
      }
    }
#pragma unroll
    for (int jj = 0; jj < 8; ++jj)                                    // column jj of Ui: rows jj - 1 .. 0
#pragma unroll
      for (int ii = jj - 1; ii >= 0; --ii) {
        double sacc = 0.0;
#pragma unroll
        for (int k = ii + 1; k <= jj; ++k) sacc = fma(a[ii][k], ui[k][jj], sacc);
        ui[ii][jj] = -sacc * ui[ii][ii];
      }
    if (tid == 0) {
#pragma unroll
      for (int ii = 0; ii < 8; ++ii) {
        sg[j0 + ii] = sv[ii];
#pragma unroll
        for (int jj = 0; jj < 8; ++jj) {
          Linv[cb * 64 + ii * 8 + jj] = jj <= ii ? li[ii][jj] : 0.0;
          Uinv[cb * 64 + ii * 8 + jj] = jj >= ii ? ui[ii][jj] : 0.0;
        }
      }
    }
    if (tid < PB) {
      if (tid >= j0 + 8) {                             // column panel: L[i][blk] = (A[i][blk] * s) Ukk^-1
        double x[8], l[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) x[k] = A[tid][j0 + k] * sv[k];
#pragma unroll
        for (int jj = 0; jj < 8; ++jj) {
          double sacc = 0.0;
#pragma unroll
          for (int k = 0; k <= jj; ++k) sacc = fma(x[k], ui[k][jj], sacc);
          l[jj] = sacc;
        }
#pragma unroll
        for (int jj = 0; jj < 8; ++jj) A[tid][j0 + jj] = l[jj];
      } else if (tid < j0) {                           // rows above the block: their U entries of these columns take the signs
#pragma unroll
        for (int k = 0; k < 8; ++k) A[tid][j0 + k] *= sv[k];
      }
    } else if (tid < 2 * PB) {
      const int k = tid - PB;
      if (k >= j0 + 8) {                               // row panel: U[blk][k] = Lkk^-1 A[blk][k]
        double y[8], u[8];
#pragma unroll
        for (int ii = 0; ii < 8; ++ii) y[ii] = A[j0 + ii][k];
#pragma unroll
        for (int ii = 0; ii < 8; ++ii) {
          double sacc = 0.0;
#pragma unroll
          for (int k2 = 0; k2 <= ii; ++k2) sacc = fma(li[ii][k2], y[k2], sacc);
          u[ii] = sacc;
        }
#pragma unroll
        for (int ii = 0; ii < 8; ++ii) A[j0 + ii][k] = u[ii];
      }
    }
    __syncthreads();
    if (tid == 0) {                                    // (after the barrier: every thread has read the original block)
#pragma unroll
      for (int ii = 0; ii < 8; ++ii)
#pragma unroll
        for (int jj = 0; jj < 8; ++jj) A[j0 + ii][j0 + jj] = a[ii][jj];
    }
    if (cb < NB8 - 1) {                                // trailing: A[i][k] -= sum_jj L[i][j0 + jj] U[j0 + jj][k]
      double lr[4][8], uc[4][8];
#pragma unroll
      for (int q = 0; q < 4; ++q)
#pragma unroll
        for (int jj = 0; jj < 8; ++jj) {            // (the diagonal block is being written back by thread 0: not read)
          lr[q][jj] = ti + 16 * q >= j0 + 8 ? A[ti + 16 * q][j0 + jj] : 0.0;
          uc[q][jj] = tk + 16 * q >= j0 + 8 ? A[j0 + jj][tk + 16 * q] : 0.0;
        }
#pragma unroll
      for (int qa = 0; qa < 4; ++qa)
#pragma unroll
        for (int qb = 0; qb < 4; ++qb) {
          const int i = ti + 16 * qa, k = tk + 16 * qb;
          if (i >= j0 + 8 && k >= j0 + 8) {
            double sacc = 0.0;
#pragma unroll
            for (int jj = 0; jj < 8; ++jj) sacc = fma(lr[qa][jj], uc[qb][jj], sacc);
            A[i][k] -= sacc;
          }
        }
    }
    __syncthreads();
  }
}

// 64 x 64 += over rows:  G[i][j] = sum_r A[r][i] B[r][j], 256 threads, thread (ty, tx) owns i = ty + 16 a, j = tx + 16 b
__device__ __forceinline__ void tile_atb(const double (*A)[PLD], const double (*B)[PLD], int rows, int tid,
                                         double (&acc)[4][4]) {
  const int tx = tid & 15, ty = tid >> 4;
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b) acc[a][b] = 0.0;
  for (int r = 0; r < rows; ++r) {
    double av[4], bv[4];
#pragma unroll
    for (int a = 0; a < 4; ++a) { av[a] = A[r][ty + 16 * a]; bv[a] = B[r][tx + 16 * a]; }
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int b = 0; b < 4; ++b) acc[a][b] = fma(av[a], bv[b], acc[a][b]);
  }
}

// ------------------------------------------------------------------ (1) optional solve + Gram partial per 64-row chunk
// X: m x 64 (ldx).  If Lf != null: Q = X L^-T is written to Qout (pitch 64) first.  Gpart[chunk] = Q_chunk^T Q_chunk.
// One extra CTA (blockIdx.x == nch) multiplies two lower-triangular factors when Ma != null: Mout = Ma Mb.
__global__ void __launch_bounds__(PT)
sbr_gram_kernel(const double* X, int64_t ldx, int m, const double* __restrict__ Lf,
                double* Qout, double* __restrict__ Gpart, int nch,
                const double* __restrict__ Ma, const double* __restrict__ Mb, double* __restrict__ Mout) {
  pdl_enter();
  extern __shared__ double smem[];
  double (*Xs)[PLD] = reinterpret_cast<double (*)[PLD]>(smem);
  double (*Ls)[PLD] = reinterpret_cast<double (*)[PLD]>(smem + PB * PLD);
  __shared__ double s_dinv[512];
  const int tid = threadIdx.x;
  if ((int)blockIdx.x == nch) {                       // lower x lower product (accumulated R^T of the passes)
    if (!Ma) return;
    load64(Xs, Ma, tid, PT);
    load64(Ls, Mb, tid, PT);
    __syncthreads();
    for (int e = tid; e < PB * PB; e += PT) {
      const int i = e >> 6, c = e & 63;
      double s = 0.0;
      for (int k = c; k <= i; ++k) s = fma(Xs[i][k], Ls[k][c], s);
      Mout[e] = s;
    }
    return;
  }
  const int r0 = blockIdx.x * PB, rows = min(PB, m - r0);
  for (int e = tid; e < PB * PB; e += PT) {
    const int r = e >> 6, c = e & 63;
    Xs[r][c] = r < rows ? X[(int64_t)(r0 + r) * ldx + c] : 0.0;
  }
  if (Lf) {
    load64(Ls, Lf, tid, PT);
    for (int e = tid; e < 512; e += PT) s_dinv[e] = Lf[PB * PB + e];
    __syncthreads();
    rows_trsm_lt(Xs, Ls, s_dinv, tid);
    __syncthreads();
    for (int e = tid; e < PB * PB; e += PT) {
      const int r = e >> 6, c = e & 63;
      if (r < rows) Qout[(int64_t)(r0 + r) * PB + c] = Xs[r][c];
    }
  } else {
    __syncthreads();
  }
  double acc[4][4];
  tile_atb(Xs, Xs, rows, tid, acc);
  double* G = Gpart + (int64_t)blockIdx.x * PB * PB;
  const int tx = tid & 15, ty = tid >> 4;
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b) G[(ty + 16 * a) * PB + tx + 16 * b] = acc[a][b];
}

// ------------------------------------------------------------------ (2) sum of the partials + factorisation (last CTA)
// mode 0: G + shift I = L L^T (first pass)   1: G = L L^T   2: as 1, then Householder reconstruction of the top block
// mode 3: C1 = T^T G (no factorisation; G = Y^T Z)
struct RedParams {
  const double* part; int npart; double* G; int* ticket; int mode; int m;
  double* Lout;            // modes 0-2: Cholesky factor (lower, 64 x 64)
  const double* Qtop;      // mode 2: top 64 rows of the current Q (pitch 64)
  double* LU;              // mode 2: strict lower = Y1, upper = U
  double* sign;            // mode 2: 64 signs
  const double* Tm;        // mode 3
  double* C1;              // mode 3
  int* fail;
};

__global__ void __launch_bounds__(PT) sbr_reduce_kernel(RedParams P) {
  pdl_enter();
  extern __shared__ double smem[];
  double (*Ls)[PLD] = reinterpret_cast<double (*)[PLD]>(smem);
  double (*Ts)[PLD] = reinterpret_cast<double (*)[PLD]>(smem + PB * PLD);
  __shared__ double s_dinv[512], s_linv[512], s_uinv[512], s_sg[PB];
  __shared__ int s_last;
  const int tid = threadIdx.x;
  {
    // 16 elements per CTA, 16 threads per element: each thread loads its partials in batches of 8 independent loads
    // (fixed order -> deterministic), then a shuffle tree
    const int e = blockIdx.x * (PT / 16) + (tid >> 4), part = tid & 15;
    double s = 0.0;
    for (int p0 = part; p0 < P.npart; p0 += 16 * 8) {
      double v[8];
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const int p = p0 + 16 * q;
        v[q] = p < P.npart ? __ldcg(P.part + (int64_t)p * PB * PB + e) : 0.0;
      }
      s += ((v[0] + v[1]) + (v[2] + v[3])) + ((v[4] + v[5]) + (v[6] + v[7]));
    }
    s += __shfl_xor_sync(0xffffffffu, s, 1);
    s += __shfl_xor_sync(0xffffffffu, s, 2);
    s += __shfl_xor_sync(0xffffffffu, s, 4);
    s += __shfl_xor_sync(0xffffffffu, s, 8);
    if (part == 0) P.G[e] = s;
  }
  __threadfence();
  __syncthreads();
  if (tid == 0) s_last = (atomicAdd(P.ticket, 1) == (int)gridDim.x - 1);
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  if (tid == 0) *P.ticket = 0;                        // re-armed for the next launch (stream ordered)
  if (P.mode == 3) {                                  // C1 = T^T C0, T upper triangular
    for (int e = tid; e < PB * PB; e += PT) { Ls[e >> 6][e & 63] = __ldcg(P.G + e); Ts[e >> 6][e & 63] = P.Tm[e]; }
    __syncthreads();
    const int i = tid >> 2, jq = tid & 3;             // row i, columns jq + 4 q
    double acc[16];
#pragma unroll
    for (int q = 0; q < 16; ++q) acc[q] = 0.0;
    for (int k = 0; k <= i; ++k) {
      const double t = Ts[k][i];
#pragma unroll
      for (int q = 0; q < 16; ++q) acc[q] = fma(t, Ls[k][jq + 4 * q], acc[q]);
    }
#pragma unroll
    for (int q = 0; q < 16; ++q) P.C1[i * PB + jq + 4 * q] = acc[q];
    return;
  }
  // ---- modes 0-2: Cholesky of the (shifted) Gram matrix
  for (int e = tid; e < PB * PB; e += PT) Ls[e >> 6][e & 63] = __ldcg(P.G + e);
  __syncthreads();
  if (P.mode == 0) {                                  // shift = 11 (m b + b (b + 1)) u trace(G)
    double tr = 0.0;
    for (int k = 0; k < PB; ++k) tr += Ls[k][k];
    const bool bad = !(tr > 0.0) || !isfinite(tr);
    if (bad && tid == 0) atomicExch(P.fail, 1);
    __syncthreads();
    if (tid < PB)
      Ls[tid][tid] = bad ? 1.0 : Ls[tid][tid] + 11.0 * ((double)P.m * PB + PB * (PB + 1)) * 1.1102230246251565e-16 * tr;
    __syncthreads();
  }
  chol64_blocked(Ls, s_dinv, tid, P.fail);
  for (int e = tid; e < PB * PB; e += PT) {
    const int r = e >> 6, c = e & 63;
    P.Lout[e] = c <= r ? Ls[r][c] : 0.0;
  }
  for (int e = tid; e < 512; e += PT) P.Lout[PB * PB + e] = s_dinv[e];
  if (P.mode != 2) return;
  // ---- Householder reconstruction on the top block: A = -(Qtop L^-T), LU without pivoting of A diag(s) + I
  for (int e = tid; e < PB * PB; e += PT) Ts[e >> 6][e & 63] = P.Qtop[e];
  __syncthreads();
  rows_trsm_lt(Ts, Ls, s_dinv, tid);
  __syncthreads();
  for (int e = tid; e < PB * PB; e += PT) Ts[e >> 6][e & 63] = -Ts[e >> 6][e & 63];
  __syncthreads();
  lu64_sign_blocked(Ts, s_linv, s_uinv, s_sg, tid);
  for (int e = tid; e < PB * PB; e += PT) P.LU[e] = Ts[e >> 6][e & 63];
  for (int e = tid; e < 512; e += PT) { P.LU[PB * PB + e] = s_linv[e]; P.LU[PB * PB + 512 + e] = s_uinv[e]; }
  if (tid < PB) P.sign[tid] = s_sg[tid];
}

// ------------------------------------------------------------------ (3) Y = reconstruction applied to all rows
// chunk CTAs: rows of Y.  Top block (chunk 0): Y1 (unit lower) from LU.  Other rows: y = ((q L3^-T) * (-s)) U^-1.
// Y goes to Ybuf (pitch 64) and to the panel's place in S (pitch lda).
// CTA nch: T = U Y1^-T -> Tout.   CTA nch + 1: R = diag(s) (M12 L3)^T -> band array (rows r0 .., columns kb ..).
__global__ void __launch_bounds__(PT)
sbr_applyfinal_kernel(const double* __restrict__ Q, int m, const double* __restrict__ L3, const double* __restrict__ LU,
                      const double* __restrict__ sign, double* __restrict__ Ybuf, double* __restrict__ Sp, int64_t lda,
                      int nch, double* __restrict__ Tout, const double* __restrict__ M12, double* __restrict__ AB,
                      int kb) {
  pdl_enter();
  extern __shared__ double smem[];
  double (*Ls)[PLD] = reinterpret_cast<double (*)[PLD]>(smem);
  double (*Us)[PLD] = reinterpret_cast<double (*)[PLD]>(smem + PB * PLD);
  double (*Xs)[PLD] = reinterpret_cast<double (*)[PLD]>(smem + 2 * PB * PLD);
  __shared__ double s_inv[512], s_uinv[512], sg[PB];
  const int tid = threadIdx.x;
  if ((int)blockIdx.x == nch) {                       // T: T Y1^T = U  (row solves with the unit lower Y1)
    load64(Ls, LU, tid, PT);
    for (int e = tid; e < PB * PB; e += PT) {
      const int r = e >> 6, c = e & 63;
      Xs[r][c] = c >= r ? LU[e] : 0.0;
    }
    for (int e = tid; e < 512; e += PT) s_inv[e] = LU[PB * PB + e];      // inverses of the unit lower diagonal blocks
    __syncthreads();
    rows_trsm_lt(Xs, Ls, s_inv, tid);
    __syncthreads();
    for (int e = tid; e < PB * PB; e += PT) Tout[e] = Xs[e >> 6][e & 63];
    return;
  }
  if ((int)blockIdx.x == nch + 1) {                   // R = diag(s) (M12 L3)^T, written into the band
    load64(Ls, M12, tid, PT);
    load64(Us, L3, tid, PT);
    if (tid < PB) sg[tid] = sign[tid];
    __syncthreads();
    for (int e = tid; e < PB * PB; e += PT) {
      const int c = e >> 6, i = e & 63;               // M[c][i], i <= c  ->  R[i][c]
      if (i > c) continue;
      double s = 0.0;
      for (int k = i; k <= c; ++k) s = fma(Ls[c][k], Us[k][i], s);
      AB[(int64_t)(kb + c) * BLD + (PB + i - c)] = sg[i] * s;
    }
    return;
  }
  const int r0 = blockIdx.x * PB, rows = min(PB, m - r0);
  if (blockIdx.x == 0) {
    for (int e = tid; e < PB * PB; e += PT) {
      const int r = e >> 6, c = e & 63;
      if (r >= rows) continue;
      const double v = c < r ? LU[e] : (c == r ? 1.0 : 0.0);
      Ybuf[(int64_t)r * PB + c] = v;
      Sp[(int64_t)r * lda + c] = v;
    }
    return;
  }
  load64(Ls, L3, tid, PT);
  load64(Us, LU, tid, PT);
  for (int e = tid; e < PB * PB; e += PT) {
    const int r = e >> 6, c = e & 63;
    Xs[r][c] = r < rows ? Q[(int64_t)(r0 + r) * PB + c] : 0.0;
  }
  if (tid < PB) sg[tid] = sign[tid];
  for (int e = tid; e < 512; e += PT) { s_inv[e] = L3[PB * PB + e]; s_uinv[e] = LU[PB * PB + 512 + e]; }
  __syncthreads();
  rows_trsm_lt(Xs, Ls, s_inv, tid);
  {
    const int r = tid >> 2, kq = tid & 3;
    for (int k = kq; k < PB; k += 4) Xs[r][k] *= -sg[k];
    __syncwarp();
  }
  rows_trsm_u(Xs, Us, s_uinv, tid);
  __syncthreads();
  for (int e = tid; e < PB * PB; e += PT) {
    const int r = e >> 6, c = e & 63;
    if (r >= rows) continue;
    const double v = Xs[r][c];
    Ybuf[(int64_t)(r0 + r) * PB + c] = v;
    Sp[(int64_t)(r0 + r) * lda + c] = v;
  }
}

// ------------------------------------------------------------------ (4) Z = A22 Y on the DMMA pipe, split-K partials
// A: m x m row-major (lda, both triangles valid), Y: m x 64 (pitch 64).  CTA tile 128 x 64, k chunk per blockIdx.y.
constexpr int SY_BM = 128, SY_BK = 16, SY_T = 256, SY_LDA = SY_BM + 4, SY_LDB = PB + 4;

__global__ void __launch_bounds__(SY_T, 2)
sbr_symm_kernel(const double* __restrict__ A, int64_t lda, int m, const double* __restrict__ Y,
                double* __restrict__ Zp, int kchunk) {
  pdl_enter();
  __shared__ double As[SY_BK][SY_LDA];
  __shared__ double Bs[SY_BK][SY_LDB];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int gid = lane >> 2, tig = lane & 3;
  const int wm = (warp & 3) * 32, wn = (warp >> 2) * 32;
  const int m0 = blockIdx.x * SY_BM;
  const int kb = blockIdx.y * kchunk, ke = min(m, kb + kchunk);
  double c[4][4][2];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) { c[i][j][0] = 0.0; c[i][j][1] = 0.0; }
  const int akk = tid & 15, ar = tid >> 4;            // A slab: 16 consecutive k of row ar + 16 i
  const int bn = tid & 63, bk = tid >> 6;             // Y slab: row k = bk + 4 i, column bn
  double ra[8], rb[4];
  auto gload = [&](int k0) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int r = m0 + ar + 16 * i, k = k0 + akk;
      ra[i] = (r < m && k < ke) ? A[(int64_t)r * lda + k] : 0.0;
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int k = k0 + bk + 4 * i;
      rb[i] = k < ke ? Y[(int64_t)k * PB + bn] : 0.0;
    }
  };
  if (kb < ke) gload(kb);
  for (int k0 = kb; k0 < ke; k0 += SY_BK) {
#pragma unroll
    for (int i = 0; i < 8; ++i) As[akk][(ar + 16 * i) ^ ((akk >> 2) & 3)] = ra[i];
#pragma unroll
    for (int i = 0; i < 4; ++i) Bs[bk + 4 * i][bn ^ (((bk + 4 * i) >> 2) & 3)] = rb[i];
    __syncthreads();
    if (k0 + SY_BK < ke) gload(k0 + SY_BK);
#pragma unroll
    for (int ks = 0; ks < SY_BK; ks += 4) {
      double a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        a[i] = As[ks + tig][(wm + 8 * i + gid) ^ ((ks >> 2) & 3)];
        b[i] = Bs[ks + tig][(wn + 8 * i + gid) ^ ((ks >> 2) & 3)];
      }
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j)
          asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                       : "+d"(c[i][j][0]), "+d"(c[i][j][1]) : "d"(a[i]), "d"(b[j]));
    }
    __syncthreads();
  }
  double* Zo = Zp + (int64_t)blockIdx.y * m * PB;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int r = m0 + wm + 8 * i + gid;
    if (r >= m) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = wn + 8 * j + 2 * tig;
      *reinterpret_cast<double2*>(Zo + (int64_t)r * PB + n) = make_double2(c[i][j][0], c[i][j][1]);
    }
  }
}

// ------------------------------------------------------------------ (8) A22 -= XY YX^T, symmetric rank-128 update on the DMMA pipe
// XY = [X | Y], YX = [Y | X] (m x 128, k contiguous).  CTA tile 128 rows x 64 columns, 256 threads, two CTAs per SM so
// that the read-modify-write of one tile overlaps the MMAs of the other (the general kernel runs one 512-thread CTA
// per SM and spent more time in the prologue / epilogue of these K = 128 tiles than in their MMAs).  Only tiles that
// touch the lower triangle are launched (1-D grid over row tile bm, column tile bn <= 2 bm + 1); entries with
// col <= row are written and mirrored, so the matrix stays exactly symmetric.
__global__ void __launch_bounds__(SY_T, 2)
sbr_syr2k_kernel(const double* __restrict__ XY, const double* __restrict__ YX, int m, double* __restrict__ D, int64_t ldd) {
  pdl_enter();
  __shared__ double As[SY_BK][SY_LDA];
  __shared__ double Bs[SY_BK][SY_LDB];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int gid = lane >> 2, tig = lane & 3;
  const int wm = (warp & 3) * 32, wn = (warp >> 2) * 32;
  // blockIdx.x -> (bm, bn): tiles of row bm start at bm (bm + 1)
  int bm = (int)((sqrtf(4.0f * (float)blockIdx.x + 1.0f) - 1.0f) * 0.5f);
  while ((bm + 1) * (bm + 2) <= (int)blockIdx.x) ++bm;
  while (bm * (bm + 1) > (int)blockIdx.x) --bm;
  const int bn = (int)blockIdx.x - bm * (bm + 1);
  const int m0 = bm * SY_BM, n0 = bn * PB;
  double c[4][4][2];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) { c[i][j][0] = 0.0; c[i][j][1] = 0.0; }
  const int akk = tid & 15, ar = tid >> 4;
  double ra[8], rb[4];
  auto gload = [&](int k0) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int r = m0 + ar + 16 * i;
      ra[i] = r < m ? XY[(int64_t)r * (2 * PB) + k0 + akk] : 0.0;
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int r = n0 + ar + 16 * i;
      rb[i] = r < m ? YX[(int64_t)r * (2 * PB) + k0 + akk] : 0.0;
    }
  };
  gload(0);
  for (int k0 = 0; k0 < 2 * PB; k0 += SY_BK) {
#pragma unroll
    for (int i = 0; i < 8; ++i) As[akk][(ar + 16 * i) ^ ((akk >> 2) & 3)] = ra[i];
#pragma unroll
    for (int i = 0; i < 4; ++i) Bs[akk][(ar + 16 * i) ^ ((akk >> 2) & 3)] = rb[i];
    __syncthreads();
    if (k0 + SY_BK < 2 * PB) gload(k0 + SY_BK);
#pragma unroll
    for (int ks = 0; ks < SY_BK; ks += 4) {
      double a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        a[i] = As[ks + tig][(wm + 8 * i + gid) ^ ((ks >> 2) & 3)];
        b[i] = Bs[ks + tig][(wn + 8 * i + gid) ^ ((ks >> 2) & 3)];
      }
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j)
          asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                       : "+d"(c[i][j][0]), "+d"(c[i][j][1]) : "d"(a[i]), "d"(b[j]));
    }
    __syncthreads();
  }
  // epilogue: all old values first (independent loads), then the stores
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int r = m0 + wm + 8 * i + gid;
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int cc = n0 + wn + 8 * j + 2 * tig + e;
        const double old = (r < m && cc <= r) ? D[(int64_t)r * ldd + cc] : 0.0;
        c[i][j][e] = old - c[i][j][e];
      }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int r = m0 + wm + 8 * i + gid;
    if (r >= m) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int cc = n0 + wn + 8 * j + 2 * tig + e;
        if (cc > r) continue;
        D[(int64_t)r * ldd + cc] = c[i][j][e];
        if (cc < r) D[(int64_t)cc * ldd + r] = c[i][j][e];
      }
  }
}

// ------------------------------------------------------------------ (5) Z = sum of the split-K partials, partial Y^T Z
__global__ void __launch_bounds__(PT)
sbr_yz_kernel(const double* __restrict__ Zp, int split, int m, const double* __restrict__ Ybuf,
              double* __restrict__ Zbuf, double* __restrict__ Gpart) {
  pdl_enter();
  extern __shared__ double smem[];
  double (*Ys)[PLD] = reinterpret_cast<double (*)[PLD]>(smem);
  double (*Zs)[PLD] = reinterpret_cast<double (*)[PLD]>(smem + PB * PLD);
  const int tid = threadIdx.x;
  const int r0 = blockIdx.x * PB, rows = min(PB, m - r0);
  {
    double z[16], y[16];
#pragma unroll
    for (int q = 0; q < 16; ++q) {
      const int e = tid + q * PT, r = e >> 6;
      const int64_t off = (int64_t)(r0 + r) * PB + (e & 63);
      z[q] = 0.0;
      y[q] = r < rows ? Ybuf[off] : 0.0;
    }
    for (int s = 0; s < split; ++s) {
      const double* zp = Zp + (int64_t)s * m * PB;
#pragma unroll
      for (int q = 0; q < 16; ++q) {
        const int e = tid + q * PT, r = e >> 6;
        if (r < rows) z[q] += zp[(int64_t)(r0 + r) * PB + (e & 63)];
      }
    }
#pragma unroll
    for (int q = 0; q < 16; ++q) {
      const int e = tid + q * PT, r = e >> 6, c = e & 63;
      if (r < rows) Zbuf[(int64_t)(r0 + r) * PB + c] = z[q];
      Ys[r][c] = y[q];
      Zs[r][c] = z[q];
    }
  }
  __syncthreads();
  double acc[4][4];
  tile_atb(Ys, Zs, rows, tid, acc);
  double* G = Gpart + (int64_t)blockIdx.x * PB * PB;
  const int tx = tid & 15, ty = tid >> 4;
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b) G[(ty + 16 * a) * PB + tx + 16 * b] = acc[a][b];
}

// ------------------------------------------------------------------ (7) X = (Z - 1/2 Y C1) T; operands of the rank-128 update
// XY = [X | Y], YX = [Y | X]  (m x 128 each):  A22 -= XY YX^T = X Y^T + Y X^T
__global__ void __launch_bounds__(PT)
sbr_xbuild_kernel(const double* __restrict__ Zbuf, const double* __restrict__ Ybuf, int m,
                  const double* __restrict__ C1, const double* __restrict__ Tm,
                  double* __restrict__ XY, double* __restrict__ YX) {
  pdl_enter();
  extern __shared__ double smem[];
  double (*Ys)[PLD] = reinterpret_cast<double (*)[PLD]>(smem);
  double (*Zs)[PLD] = reinterpret_cast<double (*)[PLD]>(smem + PB * PLD);
  double (*Cs)[PLD] = reinterpret_cast<double (*)[PLD]>(smem + 2 * PB * PLD);
  const int tid = threadIdx.x;
  const int r0 = blockIdx.x * PB, rows = min(PB, m - r0);
  for (int e = tid; e < PB * PB; e += PT) {
    const int r = e >> 6, c = e & 63;
    const int64_t off = (int64_t)(r0 + r) * PB + c;
    Ys[r][c] = r < rows ? Ybuf[off] : 0.0;
    Zs[r][c] = r < rows ? Zbuf[off] : 0.0;
    Cs[r][c] = C1[e];
  }
  __syncthreads();
  // A1 = Z - 1/2 Y C1 (thread owns 16 elements: row r = tid >> 2, columns (tid & 3) + 4 q)
  const int r = tid >> 2, cq = tid & 3;
  double a1[16];
#pragma unroll
  for (int q = 0; q < 16; ++q) a1[q] = 0.0;
  for (int k = 0; k < PB; ++k) {
    const double y = Ys[r][k];
#pragma unroll
    for (int q = 0; q < 16; ++q) a1[q] = fma(y, Cs[k][cq + 4 * q], a1[q]);
  }
  __syncthreads();
#pragma unroll
  for (int q = 0; q < 16; ++q) Zs[r][cq + 4 * q] -= 0.5 * a1[q];
  for (int e = tid; e < PB * PB; e += PT) Cs[e >> 6][e & 63] = Tm[e];
  __syncthreads();
  // X = A1 T (T upper triangular)
#pragma unroll
  for (int q = 0; q < 16; ++q) a1[q] = 0.0;
  for (int k = 0; k < PB; ++k) {
    const double z = Zs[r][k];
#pragma unroll
    for (int q = 0; q < 16; ++q) a1[q] = fma(z, Cs[k][cq + 4 * q], a1[q]);   // T[k][c] = 0 for c < k
  }
  if (r < rows) {
    double* xy = XY + (int64_t)(r0 + r) * 2 * PB;
    double* yx = YX + (int64_t)(r0 + r) * 2 * PB;
#pragma unroll
    for (int q = 0; q < 16; ++q) {
      const int c = cq + 4 * q;
      const double y = Ys[r][c];
      xy[c] = a1[q];  xy[PB + c] = y;
      yx[c] = y;      yx[PB + c] = a1[q];
    }
  }
}

// ------------------------------------------------------------------ (9) last panel, m < 64 rows: one CTA, plain Householder
// P = S[r0 .., kb .. kb + 63] (m x 64), A22 = S[r0 .., r0 ..] (m x m).  Writes R into the band, A22 back (both
// triangles), Y (m x 64 dense) into the panel's place, T (64 x 64) to Tout.
__global__ void __launch_bounds__(PT)
sbr_tail_kernel(double* __restrict__ S, int64_t lda, int n, int kb, double* __restrict__ AB, double* __restrict__ Tout) {
  extern __shared__ double smem[];
  double (*Ps)[PLD] = reinterpret_cast<double (*)[PLD]>(smem);
  double (*As)[PLD] = reinterpret_cast<double (*)[PLD]>(smem + PB * PLD);
  double (*Ys)[PLD] = reinterpret_cast<double (*)[PLD]>(smem + 2 * PB * PLD);
  double (*Ts)[PLD] = reinterpret_cast<double (*)[PLD]>(smem + 3 * PB * PLD);
  __shared__ double taus[PB], dots[PB], zz[PB];
  __shared__ double s_tau;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int r0 = kb + PB, m = n - r0;
  for (int e = tid; e < PB * PB; e += PT) {
    const int r = e >> 6, c = e & 63;
    Ps[r][c] = r < m ? S[(int64_t)(r0 + r) * lda + kb + c] : 0.0;
    As[r][c] = (r < m && c < m) ? S[(int64_t)(r0 + r) * lda + r0 + c] : 0.0;
    Ys[r][c] = 0.0;
    Ts[r][c] = 0.0;
  }
  if (tid < PB) taus[tid] = 0.0;
  __syncthreads();
  const int nref = min(m - 1, PB);
  for (int c = 0; c < nref; ++c) {
    if (warp == 0) {                                  // reflector from Ps[c .. m - 1][c]
      const int i0 = c + lane, i1 = c + lane + 32;
      double x0 = i0 < m ? Ps[i0][c] : 0.0, x1 = i1 < m ? Ps[i1][c] : 0.0;
      const double alpha = __shfl_sync(0xffffffffu, x0, 0);
      const double ssq = warp_sum((lane == 0 ? 0.0 : x0 * x0) + x1 * x1);
      double tau = 0.0, beta = alpha, sc = 0.0;
      if (ssq != 0.0) {
        const double nrm = sqrt(fma(alpha, alpha, ssq));
        beta = alpha >= 0.0 ? -nrm : nrm;
        tau = (beta - alpha) / beta;
        sc = 1.0 / (alpha - beta);
      }
      if (i0 < m) { Ys[i0][c] = lane == 0 ? 1.0 : x0 * sc; Ps[i0][c] = lane == 0 ? beta : 0.0; }
      if (i1 < m) { Ys[i1][c] = x1 * sc; Ps[i1][c] = 0.0; }
      if (lane == 0) { s_tau = tau; taus[c] = tau; }
    }
    __syncthreads();
    const double tau = s_tau;
    if (tau != 0.0) {
      // left on the remaining panel columns and on A22: thread per column
      if (tid < PB) {
        if (tid > c) {
          double dsum = 0.0;
          for (int i = c; i < m; ++i) dsum = fma(Ys[i][c], Ps[i][tid], dsum);
          dsum *= tau;
          for (int i = c; i < m; ++i) Ps[i][tid] = fma(-dsum, Ys[i][c], Ps[i][tid]);
        }
      } else if (tid < 2 * PB) {
        const int cc = tid - PB;
        if (cc < m) {
          double dsum = 0.0;
          for (int i = c; i < m; ++i) dsum = fma(Ys[i][c], As[i][cc], dsum);
          dsum *= tau;
          for (int i = c; i < m; ++i) As[i][cc] = fma(-dsum, Ys[i][c], As[i][cc]);
        }
      }
      __syncthreads();
      if (tid < m) {                                  // right on A22: thread per row
        double dsum = 0.0;
        for (int i = c; i < m; ++i) dsum = fma(As[tid][i], Ys[i][c], dsum);
        dsum *= tau;
        for (int i = c; i < m; ++i) As[tid][i] = fma(-dsum, Ys[i][c], As[tid][i]);
      }
    }
    __syncthreads();
  }
  // T (dlarft, forward columnwise): T[j][j] = tau_j, T[0:j, j] = -tau_j T[0:j, 0:j] (Y[:, 0:j]^T Y[:, j])
  for (int j = 0; j < nref; ++j) {
    if (tid < j) {
      double s = 0.0;
      for (int i = j; i < m; ++i) s = fma(Ys[i][tid], Ys[i][j], s);
      zz[tid] = s;
    }
    __syncthreads();
    if (tid < j) {
      double s = 0.0;
      for (int k = tid; k < j; ++k) s = fma(Ts[tid][k], zz[k], s);
      dots[tid] = -taus[j] * s;
    }
    __syncthreads();
    if (tid < j) Ts[tid][j] = dots[tid];
    if (tid == j) Ts[j][j] = taus[j];
    __syncthreads();
  }
  for (int e = tid; e < PB * PB; e += PT) {
    const int r = e >> 6, c = e & 63;
    Tout[e] = Ts[r][c];
    if (r < m) {
      S[(int64_t)(r0 + r) * lda + kb + c] = Ys[r][c];
      if (c < m) S[(int64_t)(r0 + r) * lda + r0 + c] = As[r][c];
      if (r <= c) AB[(int64_t)(kb + c) * BLD + (PB + r - c)] = Ps[r][c];
    }
  }
}

// ------------------------------------------------------------------ (10) band array from the reduced matrix
// AB[c][d] = A[c + d][c], d = 0 .. 64: diagonal blocks from S; the sub-diagonal blocks of factored panels were
// written by the panel kernels (R diag(s)); beyond the last panel straight from S.
__global__ void sbr_band_extract_kernel(const double* __restrict__ S, int64_t lda, int n, int npanels,
                                        double* __restrict__ AB) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int c = (int)(idx / (PB + 1)), d = (int)(idx % (PB + 1));
  if (c >= n) return;
  const int r = c + d;
  if (r >= n) return;
  const int k = c / PB;
  if (r >= (k + 1) * PB && k < npanels) return;
  AB[(int64_t)c * BLD + d] = S[(int64_t)r * lda + c];
}

// ------------------------------------------------------------------ (11) back-transformation through stage 1
// rows of Z <- Q1 row, Q1 = W_0 W_1 ... (panel order), W_p = I - Y_p T_p Y_p^T acting on rows r0_p ..: last panel first.
// One CTA per NV vectors (shared memory); per panel two passes over Y (u = Y^T z; z -= Y (T u)).
template <int NV>
__global__ void __launch_bounds__(512)
sbr_apply_q1_kernel(const double* __restrict__ S, int64_t lda, const double* __restrict__ tfac, int n, int npanels,
                    double* __restrict__ Z, int64_t ldz, int kvec) {
  extern __shared__ double smem[];
  double* zs = smem;                                  // NV x n
  double (*Ts)[PLD] = reinterpret_cast<double (*)[PLD]>(smem + (size_t)NV * n);
  double* pa = smem + (size_t)NV * n + PB * PLD;      // NV x 8 x 64 partials
  double* uu = pa + NV * 8 * PB;                      // NV x 64
  const int tid = threadIdx.x;
  const int v0 = blockIdx.x * NV;
  for (int a = 0; a < NV; ++a)
    for (int i = tid; i < n; i += 512) zs[a * n + i] = (v0 + a < kvec) ? Z[(int64_t)(v0 + a) * ldz + i] : 0.0;
  __syncthreads();
  const int c = tid & 63, g = tid >> 6;
  for (int p = npanels - 1; p >= 0; --p) {
    const int r0 = (p + 1) * PB, m = n - r0;
    const double* Y = S + (int64_t)r0 * lda + (int64_t)p * PB;
    for (int e = tid; e < PB * PB; e += 512) Ts[e >> 6][e & 63] = tfac[(int64_t)p * PB * PB + e];
    // u = Y^T z  (16 independent loads per thread in flight)
    double acc[NV];
#pragma unroll
    for (int a = 0; a < NV; ++a) acc[a] = 0.0;
    for (int rb = g; rb < m; rb += 8 * 16) {
      double y[16];
#pragma unroll
      for (int t = 0; t < 16; ++t) {
        const int r = rb + 8 * t;
        y[t] = r < m ? __ldg(Y + (int64_t)r * lda + c) : 0.0;
      }
#pragma unroll
      for (int t = 0; t < 16; ++t) {
        const int r = min(rb + 8 * t, m - 1);            // (y is zero beyond m)
#pragma unroll
        for (int a = 0; a < NV; ++a) acc[a] = fma(y[t], zs[a * n + r0 + r], acc[a]);
      }
    }
#pragma unroll
    for (int a = 0; a < NV; ++a) pa[(a * 8 + g) * PB + c] = acc[a];
    __syncthreads();
    if (tid < NV * PB) {
      const int a = tid >> 6;
      double s = 0.0;
#pragma unroll
      for (int gg = 0; gg < 8; ++gg) s += pa[(a * 8 + gg) * PB + c];
      uu[a * PB + c] = s;
    }
    __syncthreads();
    // u2 = T u  (T upper): partial over columns k = g, g + 8, ...
#pragma unroll
    for (int a = 0; a < NV; ++a) acc[a] = 0.0;
    for (int k = g; k < PB; k += 8) {
      const double t = Ts[c][k];                       // zero for k < c
#pragma unroll
      for (int a = 0; a < NV; ++a) acc[a] = fma(t, uu[a * PB + k], acc[a]);
    }
    __syncthreads();
#pragma unroll
    for (int a = 0; a < NV; ++a) pa[(a * 8 + g) * PB + c] = acc[a];
    __syncthreads();
    if (tid < NV * PB) {
      const int a = tid >> 6;
      double s = 0.0;
#pragma unroll
      for (int gg = 0; gg < 8; ++gg) s += pa[(a * 8 + gg) * PB + c];
      uu[a * PB + c] = s;
    }
    __syncthreads();
    // z -= Y u2: thread (row sub = tid >> 3, segment = tid & 7 -> 8 columns), 64 rows per trip
    {
      const int seg = tid & 7, rs = tid >> 3;
      double u2[NV][8];
#pragma unroll
      for (int a = 0; a < NV; ++a)
#pragma unroll
        for (int q = 0; q < 8; ++q) u2[a][q] = uu[a * PB + 8 * seg + q];
      for (int rb = 0; rb < m; rb += 128) {              // two blocks of 64 rows per trip: 16 loads in flight
        double yv[2][8];
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int r = rb + 64 * h + rs;
          const double* yr = Y + (int64_t)r * lda + 8 * seg;
#pragma unroll
          for (int q = 0; q < 8; ++q) yv[h][q] = r < m ? __ldg(yr + q) : 0.0;
        }
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int r = rb + 64 * h + rs;
#pragma unroll
          for (int a = 0; a < NV; ++a) {
            double s = 0.0;
#pragma unroll
            for (int q = 0; q < 8; ++q) s = fma(yv[h][q], u2[a][q], s);
            s += __shfl_xor_sync(0xffffffffu, s, 1);
            s += __shfl_xor_sync(0xffffffffu, s, 2);
            s += __shfl_xor_sync(0xffffffffu, s, 4);
            if (seg == 0 && r < m) zs[a * n + r0 + r] -= s;
          }
        }
      }
    }
    __syncthreads();
  }
  for (int a = 0; a < NV; ++a)
    if (v0 + a < kvec)
      for (int i = tid; i < n; i += 512) Z[(int64_t)(v0 + a) * ldz + i] = zs[a * n + i];
}

// ------------------------------------------------------------------ host side
static size_t al256(size_t b) { return (b + 255) / 256 * 256; }

static int sbr_npanels(int64_t n) {                  // panels with m = n - (k + 1) 64 >= 2 rows
  int k = 0;
  while (n - (int64_t)(k + 1) * PB >= 2) ++k;
  return k;
}

struct SbrPlan {
  size_t off_ab, off_q, off_y, off_z, off_xy, off_yx, off_zp, off_gpart, off_small, off_ints, off_ll, total;
  int max_split;
};

static SbrPlan sbr_plan(int64_t n) {
  SbrPlan p;
  p.max_split = 16;
  const size_t nn = (size_t)n;
  size_t o = 0;
  p.off_ab = o;    o += al256(nn * BLD * 8);
  p.off_q = o;     o += al256(nn * PB * 8);
  p.off_y = o;     o += al256(nn * PB * 8);
  p.off_z = o;     o += al256(nn * PB * 8);
  p.off_xy = o;    o += al256(nn * 2 * PB * 8);
  p.off_yx = o;    o += al256(nn * 2 * PB * 8);
  p.off_zp = o;    o += al256((size_t)p.max_split * nn * PB * 8);
  p.off_gpart = o; o += al256(((nn + PB - 1) / PB + 1) * PB * PB * 8);
  p.off_small = o; o += al256(16 * PB * PB * 8);
  p.off_ints = o;  o += al256((nn + 64) * 4);
  p.off_ll = o;    o += al256(sb_chase_ll_bytes((int)n) + 256);
  p.total = o;
  return p;
}

// internal high-priority panel stream + two events, one set per (device, caller stream) so that two reductions
// enqueued on two streams (the paired surrogate runs of rule_n) do not serialise on one panel stream.  Created once,
// kept for the life of the process (a handful per device).
struct SbrStreams { cudaStream_t owner = nullptr; bool used = false; cudaStream_t panel = nullptr; cudaEvent_t ev_qr = nullptr, ev_x = nullptr; };
static SbrStreams* sbr_streams(cudaStream_t caller) {
  constexpr int SLOTS = 4;
  static SbrStreams pool[64][SLOTS];
  static int next_slot[64];
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return nullptr;
  SbrStreams* s = nullptr;
  for (int i = 0; i < SLOTS; ++i)
    if (pool[dev][i].used && pool[dev][i].owner == caller) s = &pool[dev][i];
  if (!s) {
    s = &pool[dev][next_slot[dev]];
    next_slot[dev] = (next_slot[dev] + 1) % SLOTS;
    s->owner = caller;
    s->used = true;
  }
  if (!s->panel) {
    int lo = 0, hi = 0;
    cudaDeviceGetStreamPriorityRange(&lo, &hi);
    if (cudaStreamCreateWithPriority(&s->panel, cudaStreamNonBlocking, hi) != cudaSuccess) { s->panel = nullptr; return nullptr; }
    cudaEventCreateWithFlags(&s->ev_qr, cudaEventDisableTiming);
    cudaEventCreateWithFlags(&s->ev_x, cudaEventDisableTiming);
  }
  return s;
}

// XMCA_SYTRD2_PROF=1: CUDA events between the launches of xmca_sytrd2, summed per kernel class and printed to stderr
// (diagnostics: ncu's per-launch times are cold-cache and overstate the small straight-line kernels)
struct SbrProf {
  bool on = false;
  cudaStream_t st = nullptr;
  std::vector<cudaEvent_t> ev;
  std::vector<int> cls;
  void init(cudaStream_t s) {
    const char* e = getenv("XMCA_SYTRD2_PROF");
    on = e && e[0] == '1';
    st = s;
    if (on) mark(-1);
  }
  void mark(int c) {
    if (!on) return;
    cudaEvent_t e;
    cudaEventCreate(&e);
    cudaEventRecord(e, st);
    ev.push_back(e);
    cls.push_back(c);
  }
  void report(int64_t n) {
    if (!on) return;
    static const char* names[] = {"gram", "reduce", "applyfinal", "symm", "yz", "xbuild", "syr2k", "tail", "extract", "chase"};
    double sum[10] = {0}; int cnt[10] = {0};
    cudaEventSynchronize(ev.back());
    for (size_t i = 1; i < ev.size(); ++i) {
      float ms = 0.f;
      cudaEventElapsedTime(&ms, ev[i - 1], ev[i]);
      sum[cls[i]] += ms; cnt[cls[i]]++;
    }
    fprintf(stderr, "xmca_sytrd2 n=%lld:", (long long)n);
    for (int c = 0; c < 10; ++c) if (cnt[c]) fprintf(stderr, " %s %.2f ms/%d", names[c], sum[c], cnt[c]);
    fprintf(stderr, "\n");
    for (auto e : ev) cudaEventDestroy(e);
  }
};

}  // namespace xmca

using namespace xmca;

extern "C" size_t xmca_sytrd2_workspace_bytes(int64_t n) { return n > 0 ? sbr_plan(n).total : 0; }
extern "C" size_t xmca_sytrd2_info_offset(int64_t n) { return n > 0 ? sbr_plan(n).off_ints + sizeof(int) : 0; }
extern "C" size_t xmca_sytrd2_tfac_bytes(int64_t n) {
  const int np = n > 0 ? sbr_npanels(n) : 0;
  return (size_t)(np > 0 ? np : 1) * PB * PB * sizeof(double);
}

extern "C" int xmca_sytrd2(int64_t n, double* d_A, int64_t lda, double* d_d, double* d_e, double* d_tfac,
                           int want_vectors, void* d_workspace, size_t workspace_bytes, void* stream) {
  XMCA_REQUIRE(n >= 1 && n <= 26000, "xmca_sytrd2: n out of range");
  XMCA_REQUIRE(d_A && d_d && d_e && d_tfac && d_workspace, "xmca_sytrd2: null pointer");
  XMCA_REQUIRE(lda >= n, "xmca_sytrd2: lda < n");
  const SbrPlan pl = sbr_plan(n);
  XMCA_REQUIRE(workspace_bytes >= pl.total, "xmca_sytrd2: workspace too small");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  char* ws = reinterpret_cast<char*>(d_workspace);
  double* AB = reinterpret_cast<double*>(ws + pl.off_ab);
  double* Qb = reinterpret_cast<double*>(ws + pl.off_q);
  double* Yb = reinterpret_cast<double*>(ws + pl.off_y);
  double* Zb = reinterpret_cast<double*>(ws + pl.off_z);
  double* XY = reinterpret_cast<double*>(ws + pl.off_xy);
  double* YX = reinterpret_cast<double*>(ws + pl.off_yx);
  double* Zp = reinterpret_cast<double*>(ws + pl.off_zp);
  double* Gpart = reinterpret_cast<double*>(ws + pl.off_gpart);
  double* small = reinterpret_cast<double*>(ws + pl.off_small);
  const int SL = PB * PB + 1024;                      // slot: 64 x 64 matrix + two sets of 8 x 8 diagonal-block inverses
  double* G = small, *L1 = small + SL, *L2 = small + 2 * SL, *L3 = small + 3 * SL, *LU = small + 4 * SL,
        *M12 = small + 5 * SL, *C1 = small + 6 * SL, *sign = small + 7 * SL;
  int* ints = reinterpret_cast<int*>(ws + pl.off_ints);
  int* ticket = ints, *fail = ints + 1, *counters = ints + 8;

  XMCA_CUDA(cudaMemsetAsync(AB, 0, sizeof(double) * (size_t)n * BLD, st));
  XMCA_CUDA(cudaMemsetAsync(ints, 0, 32, st));
  const int np = sbr_npanels(n);
  const size_t sm2 = 2 * PB * PLD * sizeof(double), sm3 = 3 * PB * PLD * sizeof(double), sm4 = 4 * PB * PLD * sizeof(double);
  XMCA_CUDA(cudaFuncSetAttribute(sbr_gram_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm2));
  XMCA_CUDA(cudaFuncSetAttribute(sbr_reduce_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm2));
  XMCA_CUDA(cudaFuncSetAttribute(sbr_yz_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm2));
  XMCA_CUDA(cudaFuncSetAttribute(sbr_applyfinal_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm3));
  XMCA_CUDA(cudaFuncSetAttribute(sbr_xbuild_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm3));
  XMCA_CUDA(cudaFuncSetAttribute(sbr_tail_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm4));
  const int nsm = sm_count();
  const char* pdl_env = getenv("XMCA_SYTRD2_PDL");       // "0": ordinary launches (A/B runs)
  const bool pdl = !(pdl_env && pdl_env[0] == '0');
  SbrProf prof;
  prof.init(st);
  // Look-ahead: the factorisation of panel p + 1 (latency bound: three Gram / Cholesky rounds and the reconstruction)
  // runs on an internal high-priority stream `sp` next to the rank-128 update of the trailing matrix by panel p on the
  // caller's stream.  For that the update is split: the first block column of A22 (= the next panel and the next
  // diagonal block) is updated by a skinny product on `sp`, the rest A22[64:, 64:] by the big one on `st`.
  SbrStreams* ss = prof.on ? nullptr : sbr_streams(st);
  cudaStream_t sp = ss ? ss->panel : st;
  if (ss) {
    XMCA_CUDA(cudaEventRecord(ss->ev_x, st));
    XMCA_CUDA(cudaStreamWaitEvent(sp, ss->ev_x, 0));
  }

  for (int p = 0; p < np; ++p) {
    const int kb = p * PB, r0 = kb + PB, m = (int)n - r0;
    double* Tp = d_tfac + (int64_t)p * PB * PB;
    if (m < PB) {
      if (ss) {                                       // join: the tail runs on the caller's stream
        XMCA_CUDA(cudaEventRecord(ss->ev_qr, sp));
        XMCA_CUDA(cudaStreamWaitEvent(st, ss->ev_qr, 0));
        sp = st;
        ss = nullptr;
      }
      sbr_tail_kernel<<<1, PT, sm4, st>>>(d_A, lda, (int)n, kb, AB, Tp);
      XMCA_LAUNCHED();
      prof.mark(7);
      continue;
    }
    const int nch = (m + PB - 1) / PB;
    double* Pp = d_A + (int64_t)r0 * lda + kb;        // the panel, m x 64, pitch lda
    RedParams R;
    R.part = Gpart; R.npart = nch; R.G = G; R.ticket = ticket; R.m = m; R.Qtop = Qb; R.LU = LU; R.sign = sign;
    R.Tm = Tp; R.C1 = C1; R.fail = fail;
    // ---- panel stream: pass 1 (shifted), 2, 3, reconstruction
    XMCA_CUDA(launch_pdl(pdl, sbr_gram_kernel, dim3(nch), dim3(PT), sm2, sp, Pp, lda, m, nullptr, nullptr, Gpart, nch, nullptr, nullptr, nullptr));
    XMCA_LAUNCHED();
    prof.mark(0);
    R.mode = 0; R.Lout = L1;
    XMCA_CUDA(launch_pdl(pdl, sbr_reduce_kernel, dim3(NRED), dim3(PT), sm2, sp, R));
    XMCA_LAUNCHED();
    prof.mark(1);
    XMCA_CUDA(launch_pdl(pdl, sbr_gram_kernel, dim3(nch), dim3(PT), sm2, sp, Pp, lda, m, L1, Qb, Gpart, nch, nullptr, nullptr, nullptr));
    XMCA_LAUNCHED();
    prof.mark(0);
    R.mode = 1; R.Lout = L2;
    XMCA_CUDA(launch_pdl(pdl, sbr_reduce_kernel, dim3(NRED), dim3(PT), sm2, sp, R));
    XMCA_LAUNCHED();
    prof.mark(1);
    XMCA_CUDA(launch_pdl(pdl, sbr_gram_kernel, dim3(nch + 1), dim3(PT), sm2, sp, Qb, PB, m, L2, Qb, Gpart, nch, L1, L2, M12));
    XMCA_LAUNCHED();
    prof.mark(0);
    R.mode = 2; R.Lout = L3;
    XMCA_CUDA(launch_pdl(pdl, sbr_reduce_kernel, dim3(NRED), dim3(PT), sm2, sp, R));
    XMCA_LAUNCHED();
    prof.mark(1);
    XMCA_CUDA(launch_pdl(pdl, sbr_applyfinal_kernel, dim3(nch + 2), dim3(PT), sm3, sp, Qb, m, L3, LU, sign, Yb, Pp, lda, nch, Tp, M12, AB, kb));
    XMCA_LAUNCHED();
    prof.mark(2);
    if (ss) {
      XMCA_CUDA(cudaEventRecord(ss->ev_qr, sp));
      XMCA_CUDA(cudaStreamWaitEvent(st, ss->ev_qr, 0));
    }
    // ---- caller's stream: Z = A22 Y (A22 complete: the previous big update is ordered before it on `st`)
    const int tiles = (m + SY_BM - 1) / SY_BM;
    // split-K factor: the one that fills whole waves of 2 CTAs per SM best (a nearly empty last wave costs a full one)
    int split = 1;
    {
      const int slots = 2 * nsm, smax = min(pl.max_split, (m + 127) / 128);
      double best = -1.0;
      for (int sq = 1; sq <= smax; ++sq) {
        const int ctas = tiles * sq, waves = (ctas + slots - 1) / slots;
        const double eff = (double)ctas / ((double)waves * slots) - 0.004 * sq;       // (mild preference for fewer partials)
        if (eff > best) { best = eff; split = sq; }
      }
    }
    int kchunk = ((m + split - 1) / split + SY_BK - 1) / SY_BK * SY_BK;
    split = (m + kchunk - 1) / kchunk;
    double* A22 = d_A + (int64_t)r0 * lda + r0;
    XMCA_CUDA(launch_pdl(pdl, sbr_symm_kernel, dim3(tiles, split), dim3(SY_T), 0, st, A22, lda, m, Yb, Zp, kchunk));
    XMCA_LAUNCHED();
    prof.mark(3);
    XMCA_CUDA(launch_pdl(pdl, sbr_yz_kernel, dim3(nch), dim3(PT), sm2, st, Zp, split, m, Yb, Zb, Gpart));
    XMCA_LAUNCHED();
    prof.mark(4);
    R.mode = 3;
    XMCA_CUDA(launch_pdl(pdl, sbr_reduce_kernel, dim3(NRED), dim3(PT), sm2, st, R));
    XMCA_LAUNCHED();
    prof.mark(1);
    XMCA_CUDA(launch_pdl(pdl, sbr_xbuild_kernel, dim3(nch), dim3(PT), sm3, st, Zb, Yb, m, C1, Tp, XY, YX));
    XMCA_LAUNCHED();
    prof.mark(5);
    if (ss) {
      XMCA_CUDA(cudaEventRecord(ss->ev_x, st));
      XMCA_CUDA(cudaStreamWaitEvent(sp, ss->ev_x, 0));
    }
    // ---- A22 -= XY YX^T in two pieces: first block column (next panel + next diagonal block) on the panel stream,
    //      A22[64:, 64:] on the caller's stream (symmetric: lower tiles computed, upper mirrored)
    int rc = xmca_gemm_ex(1, 1, m, PB, 2 * PB, -1.0, XY, XMCA_F64, 2 * PB, YX, XMCA_F64, 2 * PB, A22, XMCA_F64, lda, 1,
                          XMCA_F64, 1, nullptr, 0, 0, sp);
    if (rc != XMCA_OK) return rc;
    if (m > PB) {
      const int mm = m - PB, tm = (mm + SY_BM - 1) / SY_BM;
      XMCA_CUDA(launch_pdl(pdl, sbr_syr2k_kernel, dim3(tm * (tm + 1)), dim3(SY_T), 0, st, XY + (int64_t)PB * 2 * PB, YX + (int64_t)PB * 2 * PB, mm, A22 + (int64_t)PB * lda + PB, lda));
      XMCA_LAUNCHED();
    }
    prof.mark(6);
  }
  if (ss) {                                           // join
    XMCA_CUDA(cudaEventRecord(ss->ev_qr, sp));
    XMCA_CUDA(cudaStreamWaitEvent(st, ss->ev_qr, 0));
  }
  {
    const int64_t tot = n * (PB + 1);
    sbr_band_extract_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, st>>>(d_A, lda, (int)n, np, AB);
    XMCA_LAUNCHED();
    prof.mark(8);
  }
  // stage 2; the reflectors go to the strict upper triangle of A (row j, columns > j), free after stage 1
  if (!(want_vectors & 4)) {                         // (bit 2: stop after stage 1 -- diagnostics, scripts/check_sytrd2.py)
    int rc = sb_chase(AB, (int)n, counters, (want_vectors & 1) ? d_A : nullptr, lda, d_d, d_e, st, nullptr, ws + pl.off_ll);
    if (rc != XMCA_OK) return rc;
    prof.mark(9);
  }
  if (want_vectors & 8) return XMCA_OK;                // asynchronous: the caller reads the flag at xmca_sytrd2_info_offset
  int h_fail = 0;
  XMCA_CUDA(cudaMemcpyAsync(&h_fail, fail, sizeof(int), cudaMemcpyDeviceToHost, st));
  XMCA_CUDA(cudaStreamSynchronize(st));
  prof.report(n);
  if (h_fail)
    return ::xmca::fail(XMCA_NUMERIC, "xmca_sytrd2: panel factorisation broke down (matrix not finite or panel rank deficient)",
                        __FILE__, __LINE__);
  return XMCA_OK;
}

// workspace of xmca_ormtr2: U, U2 (k x 64 each) and the split-K partials of U = Z Y
static int ormtr2_split(int64_t m) {                  // the product has ONE 128 x 128 output tile: K chunks of 64 spread it
  int64_t s = m / 64;                                  // over the SMs (a CTA's time is its number of k steps)
  return (int)(s < 1 ? 1 : (s > 128 ? 128 : s));
}
extern "C" size_t xmca_ormtr2_workspace_bytes(int64_t n, int64_t k) {
  if (n <= 0 || k <= 0) return 0;
  return al256((size_t)k * PB * 8) * 2 + al256(xmca_gemm_workspace_bytes(k, PB, ormtr2_split(n), XMCA_F64)) + 256;
}

extern "C" int xmca_ormtr2(int64_t n, const double* d_A, int64_t lda, const double* d_tfac, int64_t k,
                           double* d_Z, int64_t ldz, void* d_workspace, size_t workspace_bytes, void* stream) {
  XMCA_REQUIRE(n >= 1 && k >= 0 && d_A && d_tfac && d_Z && ldz >= n, "xmca_ormtr2: bad argument");
  if (k == 0) return XMCA_OK;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  int rc = sb_apply_q2(d_A, lda, (int)n, d_Z, ldz, (int)k, st);
  if (rc != XMCA_OK) return rc;
  const int np = sbr_npanels(n);
  if (np == 0) return XMCA_OK;
  if (d_workspace && workspace_bytes >= xmca_ormtr2_workspace_bytes(n, k)) {
    // stage 1 as products over ALL vectors at once (Y is read twice per panel in total, not twice per vector):
    //   U = Z[:, r0:] Y (split-K), U2 = U T^T, Z[:, r0:] -= U2 Y^T        (rows of Z: z <- z (I - Y T^T Y^T))
    char* ws = reinterpret_cast<char*>(d_workspace);
    double* U = reinterpret_cast<double*>(ws);
    double* U2 = reinterpret_cast<double*>(ws + al256((size_t)k * PB * 8));
    void* gws = ws + 2 * al256((size_t)k * PB * 8);
    for (int p = np - 1; p >= 0; --p) {
      const int64_t r0 = (int64_t)(p + 1) * PB, m = n - r0;
      const double* Y = d_A + r0 * lda + (int64_t)p * PB;
      const double* T = d_tfac + (int64_t)p * PB * PB;
      double* Zs = d_Z + r0;
      const int split = ormtr2_split(m);
      rc = xmca_gemm_ex(1, 0, k, PB, m, 1.0, Zs, XMCA_F64, ldz, Y, XMCA_F64, lda, U, XMCA_F64, PB, 0, XMCA_F64, split,
                        gws, xmca_gemm_workspace_bytes(k, PB, split, XMCA_F64), 0, stream);
      if (rc != XMCA_OK) return rc;
      rc = xmca_gemm_ex(1, 1, k, PB, PB, 1.0, U, XMCA_F64, PB, T, XMCA_F64, PB, U2, XMCA_F64, PB, 0, XMCA_F64, 1, nullptr, 0,
                        0, stream);
      if (rc != XMCA_OK) return rc;
      rc = xmca_gemm_ex(1, 1, k, m, PB, -1.0, U2, XMCA_F64, PB, Y, XMCA_F64, lda, Zs, XMCA_F64, ldz, 1, XMCA_F64, 1, nullptr,
                        0, 0, stream);
      if (rc != XMCA_OK) return rc;
    }
    return XMCA_OK;
  }
  const size_t extra = (PB * PLD + 2 * 8 * PB + 2 * PB) * sizeof(double);
  const size_t one = sizeof(double) * (size_t)n;
  if (2 * one + extra <= 200 * 1024 && k > sm_count()) {
    const size_t sm = 2 * one + extra;
    XMCA_CUDA(cudaFuncSetAttribute(sbr_apply_q1_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
    sbr_apply_q1_kernel<2><<<(unsigned)((k + 1) / 2), 512, sm, st>>>(d_A, lda, d_tfac, (int)n, np, d_Z, ldz, (int)k);
  } else {
    const size_t sm = one + extra;
    XMCA_REQUIRE(sm <= 227 * 1024, "xmca_ormtr2: n too large for the shared-memory vector");
    XMCA_CUDA(cudaFuncSetAttribute(sbr_apply_q1_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
    sbr_apply_q1_kernel<1><<<(unsigned)k, 512, sm, st>>>(d_A, lda, d_tfac, (int)n, np, d_Z, ldz, (int)k);
  }
  XMCA_LAUNCHED();
  return XMCA_OK;
}

// diagnostics (scripts/check_sytrd2.py): stage 2 alone on a caller-built band array (n x 128 doubles, see sbtrd.cu)
extern "C" size_t xmca_dbg_band_chase_ll_bytes(int64_t n) { return sb_chase_ll_bytes((int)n); }
extern "C" int xmca_dbg_band_chase(int64_t n, double* d_AB, double* d_d, double* d_e, double* d_V2, int64_t ldv,
                                   int* d_counters, long long* d_prof, void* d_ll, void* stream) {
  return sb_chase(d_AB, (int)n, d_counters, d_V2, ldv, d_d, d_e, reinterpret_cast<cudaStream_t>(stream), d_prof, d_ll);
}
