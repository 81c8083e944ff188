// DRAFT -- not part of libxmca_b200.so (__graft_entry__.build() compiles csrc/*.cu only), never run on hardware.
// Stage 2 of the two-stage tridiagonalisation planned for round 2: symmetric BAND matrix
// (bandwidth b = 64) -> tridiagonal by bulge chasing.  The algorithm, the storage and the concurrency rule are
// the ones verified in numpy in scripts/proto/two_stage_sytrd.py (stage2_band_storage) and
// scripts/proto/bulge_chase_schedule.py (LAG = 2); this file is their transcription, to be debugged on a GPU.
//
// Storage: AB[j * LD + d] = A[j + d][j], d = 0 .. 2b (LD = 2b + 1): every column keeps its sub-diagonal part
// contiguous, with room for the bulge.  Task (j, k) = k-th chase step of sweep j works on the index range
// I = [r0, r1), r0 = j + 1 + k b, with the reflector taken from column col (= j for k = 0, r0 - b otherwise):
//   left   A[I, lo:r0]  <- H A[I, lo:r0]     diag  A[I, I] <- H A[I, I] H     below  A[r1:hi, I] <- A[r1:hi, I] H
// One persistent CTA per sweep (sweeps handed out in order by an atomic counter); task k of sweep j starts once
// sweep j - 1 has published k + 3 finished tasks (or all of its tasks).
// Known to-do before it can run: the three staging blocks (100 KB) must move to dynamic shared memory with the
// opt-in attribute; consecutive tasks of a sweep should keep the block "below" in shared memory as the next
// task's "left" block instead of storing and re-loading it; grid = co-resident CTAs only (cooperative launch).
#include <cuda_runtime.h>
#include <stdint.h>
#include <math.h>

namespace xmca_draft {

constexpr int SB = 64;                 // bandwidth
constexpr int LD = 2 * SB + 1;         // column stride of the band array
constexpr int SB_THREADS = 256;

__device__ __forceinline__ int tasks_of_sweep(int n, int j) {
  const int rem = n - 2 - j;           // rows r0 = j + 1 + k b < n - 1
  return rem <= 0 ? 0 : (rem + SB - 1) / SB;
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Dense window in shared memory, column-major with stride WS: column c of the window = matrix column lo + c,
// row r of the window = matrix row r0 - SB + r ... kept simple here: the three blocks are staged separately.
__global__ void __launch_bounds__(SB_THREADS)
sbtrd_chase_kernel(double* __restrict__ AB, int n, int* __restrict__ next_sweep, int* __restrict__ progress,
                   double* __restrict__ d_out, double* __restrict__ e_out) {
  __shared__ double Lb[SB][SB + 1];    // left block   [row i of I][column c - lo]
  __shared__ double Db[SB][SB + 1];    // diagonal block (full symmetric copy)
  __shared__ double Bb[SB][SB + 1];    // block below  [row r - r1][column k of I]
  __shared__ double v[SB], pv[SB], wv[SB];
  __shared__ double s_tau, s_red[SB_THREADS / 32];
  __shared__ int s_sweep;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

  for (;;) {
    if (tid == 0) s_sweep = atomicAdd(next_sweep, 1);
    __syncthreads();
    const int j = s_sweep;
    __syncthreads();
    if (j >= n - 2) break;
    const int nt = tasks_of_sweep(n, j);
    const int nt_prev = j > 0 ? tasks_of_sweep(n, j - 1) : 0;
    int col = j, r0 = j + 1;
    for (int k = 0; k < nt; ++k) {
      // ---- wait for the previous sweep to be far enough ahead (acquire)
      if (j > 0 && tid == 0) {
        const int need = min(k + 3, nt_prev);
        int seen;
        do {
          asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(seen) : "l"(progress + (j - 1)) : "memory");
        } while (seen < need);
      }
      __syncthreads();
      const int r1 = min(r0 + SB, n), m = r1 - r0;
      const int lo = max(col, r0 - SB), hi = min(n, r1 + SB);
      const int nl = r0 - lo, nbw = hi - r1;                      // columns of the left block, rows of the block below
      // ---- stage the three blocks (each matrix column is one contiguous run of the band array)
      for (int e = tid; e < nl * m; e += SB_THREADS) {
        const int c = e / m, i = e - c * m;
        Lb[i][c] = AB[(int64_t)(lo + c) * LD + (r0 + i - lo - c)];
      }
      for (int e = tid; e < m * m; e += SB_THREADS) {
        const int kk = e / m, i = e - kk * m;                     // column kk of I, row i
        if (i >= kk) { const double a = AB[(int64_t)(r0 + kk) * LD + (i - kk)]; Db[i][kk] = a; Db[kk][i] = a; }
      }
      for (int e = tid; e < m * nbw; e += SB_THREADS) {
        const int kk = e / nbw, r = e - kk * nbw;
        Bb[r][kk] = AB[(int64_t)(r0 + kk) * LD + (r1 + r - r0 - kk)];
      }
      __syncthreads();
      // ---- reflector from the first column of the left block (dlarfg), warp 0
      if (warp == 0) {
        const int c0 = col - lo;
        const double x0 = Lb[0][c0];
        double ss = 0.0;
        for (int i = 1 + lane; i < m; i += 32) ss = fma(Lb[i][c0], Lb[i][c0], ss);
        ss = warp_sum(ss);
        double tau = 0.0, beta = x0, scale = 0.0;
        if (ss > 0.0) {
          beta = -copysign(sqrt(fma(x0, x0, ss)), x0);
          tau = (beta - x0) / beta;
          scale = 1.0 / (x0 - beta);
        }
        for (int i = lane; i < m; i += 32) v[i] = (i == 0) ? 1.0 : Lb[i][c0] * scale;
        __syncwarp();
        for (int i = lane; i < m; i += 32) Lb[i][c0] = (i == 0) ? beta : 0.0;       // the eliminated column
        if (lane == 0) s_tau = tau;
      }
      __syncthreads();
      const double tau = s_tau;
      if (tau != 0.0) {
        const int c0 = col - lo;
        // ---- left block: every other column  x <- x - tau v (v . x)      (one warp per column)
        for (int c = warp; c < nl; c += SB_THREADS / 32) {
          if (c == c0) continue;
          double dt = 0.0;
          for (int i = lane; i < m; i += 32) dt = fma(v[i], Lb[i][c], dt);
          dt = tau * warp_sum(dt);
          for (int i = lane; i < m; i += 32) Lb[i][c] = fma(-dt, v[i], Lb[i][c]);
        }
        // ---- block below: every row  y <- y - tau (y . v) v                 (one warp per row)
        for (int r = warp; r < nbw; r += SB_THREADS / 32) {
          double dt = 0.0;
          for (int kk = lane; kk < m; kk += 32) dt = fma(Bb[r][kk], v[kk], dt);
          dt = tau * warp_sum(dt);
          for (int kk = lane; kk < m; kk += 32) Bb[r][kk] = fma(-dt, v[kk], Bb[r][kk]);
        }
        // ---- diagonal block: p = tau D v, w = p - (tau/2)(v . p) v, D <- D - v w^T - w v^T
        for (int i = warp; i < m; i += SB_THREADS / 32) {
          double dt = 0.0;
          for (int kk = lane; kk < m; kk += 32) dt = fma(Db[i][kk], v[kk], dt);
          dt = warp_sum(dt);
          if (lane == 0) pv[i] = tau * dt;
        }
        __syncthreads();
        {
          double dt = 0.0;
          for (int i = tid; i < m; i += SB_THREADS) dt = fma(v[i], pv[i], dt);
          dt = warp_sum(dt);
          if (lane == 0) s_red[warp] = dt;
          __syncthreads();
          double vp = 0.0;
          for (int w = 0; w < SB_THREADS / 32; ++w) vp += s_red[w];
          for (int i = tid; i < m; i += SB_THREADS) wv[i] = fma(-0.5 * tau * vp, v[i], pv[i]);
        }
        __syncthreads();
        for (int e = tid; e < m * m; e += SB_THREADS) {
          const int i = e / m, kk = e - i * m;
          Db[i][kk] -= v[i] * wv[kk] + wv[i] * v[kk];
        }
      }
      __syncthreads();
      // ---- write the blocks back
      for (int e = tid; e < nl * m; e += SB_THREADS) {
        const int c = e / m, i = e - c * m;
        AB[(int64_t)(lo + c) * LD + (r0 + i - lo - c)] = Lb[i][c];
      }
      for (int e = tid; e < m * m; e += SB_THREADS) {
        const int kk = e / m, i = e - kk * m;
        if (i >= kk) AB[(int64_t)(r0 + kk) * LD + (i - kk)] = Db[i][kk];
      }
      for (int e = tid; e < m * nbw; e += SB_THREADS) {
        const int kk = e / nbw, r = e - kk * nbw;
        AB[(int64_t)(r0 + kk) * LD + (r1 + r - r0 - kk)] = Bb[r][kk];
      }
      // ---- publish (release): this sweep has finished k + 1 tasks
      __threadfence();
      __syncthreads();
      if (tid == 0) asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(progress + j), "r"(k + 1) : "memory");
      col = r0; r0 = r1;
    }
  }
  // (d, e) are read off the band array by the caller once every sweep has finished: d[j] = AB[j * LD],
  // e[j] = AB[j * LD + 1]
  (void)d_out; (void)e_out;
}

}  // namespace xmca_draft
