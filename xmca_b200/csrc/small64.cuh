// 64 x 64 Cholesky in shared memory, blocked by 8 (256 threads) -- shared by the panel factorisations of the two-stage
// tridiagonalisation (sbr.cu) and the diagonal blocks of xmca_cholesky (cholesky.cu).
#pragma once
#include "common.cuh"

namespace xmca {

constexpr int PB = 64;              // block size
constexpr int PLD = PB + 1;         // shared-memory pitch

// ---- 64 x 64 factorisations / triangular solves, blocked by 8 (shared memory, 256 threads).
// Measured dead ends: rows in registers with the 64-step recurrence fully unrolled (~100 KB of straight-line code per
// solve, runs at instruction-fetch speed: 45-65 us per factorisation), element-wise rolled loops in shared memory
// (every step re-reads the matrix: bound by shared-memory bandwidth / latency, 50 us).  Here every 8 x 8 diagonal
// block is factorised and inverted in REGISTERS by every thread redundantly (no broadcast, no barrier inside the
// 8-step chain; ~500 instructions that are re-executed eight times), the rest is 8-wide rank updates.
// Dinv[cb * 64 + ii * 8 + jj]: inverse of diagonal block cb.
constexpr int NB8 = 8;

// Cholesky A = L L^T (lower part in place; the upper part is not touched), Dinv: inverses of the diagonal blocks of L.
// fail: set to 2 when a pivot is not above min_pivot / not finite (it is replaced by 1 and the factorisation goes on);
// bad_col (optional, shared memory, zero on entry): 1 + the first failing column.
__device__ __forceinline__ void chol64_blocked(double (*A)[PLD], double* Dinv, int tid, int* fail,
                                               double min_pivot = 0.0, int* bad_col = nullptr) {
  const int ti = tid >> 4, tk = tid & 15;
  for (int cb = 0; cb < NB8; ++cb) {
    const int j0 = 8 * cb;
    double a[8][8], li[8][8];
#pragma unroll
    for (int ii = 0; ii < 8; ++ii)
#pragma unroll
      for (int jj = 0; jj <= ii; ++jj) a[ii][jj] = A[j0 + ii][j0 + jj];
    bool bad = false;
    int first_bad = 0;
#pragma unroll
    for (int jj = 0; jj < 8; ++jj) {
      double d = a[jj][jj];
      if (!(d > min_pivot) || !isfinite(d)) { if (!bad) first_bad = j0 + jj; d = 1.0; bad = true; }
      const double rinv = fast_rsqrt(d);
      a[jj][jj] = d * rinv;
      li[jj][jj] = rinv;
#pragma unroll
      for (int ii = jj + 1; ii < 8; ++ii) a[ii][jj] *= rinv;
#pragma unroll
      for (int kk = jj + 1; kk < 8; ++kk)
#pragma unroll
        for (int ii = kk; ii < 8; ++ii) a[ii][kk] = fma(-a[ii][jj], a[kk][jj], a[ii][kk]);
    }
    if (bad && tid == 0) {
      atomicExch(fail, 2);
      if (bad_col && *bad_col == 0) *bad_col = first_bad + 1;
    }
    // inverse of the 8 x 8 lower block: column jj of Li by forward substitution
#pragma unroll
    for (int jj = 0; jj < 8; ++jj)
#pragma unroll
      for (int ii = jj + 1; ii < 8; ++ii) {
        double sacc = 0.0;
#pragma unroll
        for (int k = jj; k < ii; ++k) sacc = fma(a[ii][k], li[k][jj], sacc);
        li[ii][jj] = -sacc * li[ii][ii];
      }
    if (tid == 0) {
#pragma unroll
      for (int ii = 0; ii < 8; ++ii)
#pragma unroll
        for (int jj = 0; jj < 8; ++jj) Dinv[cb * 64 + ii * 8 + jj] = jj <= ii ? li[ii][jj] : 0.0;
    }
    // panel: L[i][j0 .. j0 + 7] = A[i][j0 ..] Lkk^-T for the rows below the block (thread = row)
    if (tid < PB && tid >= j0 + 8) {
      double x[8], l[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) x[k] = A[tid][j0 + k];
#pragma unroll
      for (int jj = 0; jj < 8; ++jj) {
        double sacc = 0.0;
#pragma unroll
        for (int k = 0; k <= jj; ++k) sacc = fma(x[k], li[jj][k], sacc);
        l[jj] = sacc;
      }
#pragma unroll
      for (int jj = 0; jj < 8; ++jj) A[tid][j0 + jj] = l[jj];
    }
    __syncthreads();
    if (tid == 0) {                                  // (after the barrier: every thread has read the original block)
#pragma unroll
      for (int ii = 0; ii < 8; ++ii)
#pragma unroll
        for (int jj = 0; jj <= ii; ++jj) A[j0 + ii][j0 + jj] = a[ii][jj];
    }
    // trailing update: A[i][k] -= sum_jj L[i][j0 + jj] L[k][j0 + jj], rows ti + 16 a, columns tk + 16 b
    if (cb < NB8 - 1) {
      double lr[4][8], lc[4][8];
#pragma unroll
      for (int q = 0; q < 4; ++q)
#pragma unroll
        for (int jj = 0; jj < 8; ++jj) {            // (rows of the diagonal block are being written back by thread 0: not read)
          lr[q][jj] = ti + 16 * q >= j0 + 8 ? A[ti + 16 * q][j0 + jj] : 0.0;
          lc[q][jj] = tk + 16 * q >= j0 + 8 ? A[tk + 16 * q][j0 + jj] : 0.0;
        }
#pragma unroll
      for (int qa = 0; qa < 4; ++qa)
#pragma unroll
        for (int qb = 0; qb < 4; ++qb) {
          const int i = ti + 16 * qa, k = tk + 16 * qb;
          if (k >= j0 + 8 && i >= k) {
            double sacc = 0.0;
#pragma unroll
            for (int jj = 0; jj < 8; ++jj) sacc = fma(lr[qa][jj], lc[qb][jj], sacc);
            A[i][k] -= sacc;
          }
        }
    }
    __syncthreads();
  }
}

}  // namespace xmca
