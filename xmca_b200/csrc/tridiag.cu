// Symmetric eigen-solver by Householder tridiagonalisation (fp64):
//   xmca_sytrd : S = Q T Q^T, blocked (LAPACK dsytrd/dlatrd scheme), one persistent
//                cooperative kernel per panel of 64 columns + one rank-128 update
//   xmca_stebz : all eigenvalues of T by bisection on the Sturm count
//   xmca_stein : selected eigenvectors of T by inverse iteration (pivoted LU of
//                T - lambda I), modified Gram-Schmidt inside eigenvalue clusters
//   xmca_ormtr : back-transformation  x = Q z
//
// This is the engine's full-spectrum route for large problems: it replaces the
// np.linalg.svd calls of array.py:479 (x2) and :570 by the eigen-decomposition
// of ONE T x T symmetric matrix (see engine.py), at 4/3 n^3 flops instead of the
// ~150 n^3 of a converged Jacobi SVD.  The dominant cost is one streaming pass
// over the trailing matrix per column (y = A v): n^3 * 8 / 3 bytes in total,
// HBM-bound -- that pass is the roofline kernel of solve().
#include "common.cuh"
#include <cooperative_groups.h>
#include <math.h>
#include <stdlib.h>

namespace cg = cooperative_groups;

namespace xmca {

constexpr int TD_NB = 64;                 // panel width
constexpr int TD_THREADS = 1024;          // one CTA per SM
constexpr int TD_WARPS = TD_THREADS / 32;
constexpr int TD_K2 = 2 * TD_NB;          // row length of the [V | W] panel buffers
constexpr int TD_PART = TD_K2 + 8;        // per-CTA reduction slots: [0] ssq, [1..128] p, [130] w.v
constexpr int TD_MAXF = 16;               // max column split of one row in the symv
constexpr int TD_TS = 64;                 // tile size of the tile-major trailing matrix (= panel width)
constexpr int TD_DEFAULT_VARIANT = 7;     // see XMCA_SYTRD_VARIANT in xmca_sytrd
constexpr double TD_KEEP_MB = 88.0;       // trailing matrices up to this size are kept L2-resident (plain loads)
constexpr int TD_TILE_MIN = 4096;         // tile-major one-triangle passes while the trailing size exceeds this

struct SytrdParams {
  double* A; int64_t lda; int n;
  int j0, nb;
  double* VW;          // n x 128 row-major: [V | W]
  double* WV;          // n x 128 row-major: [W | V]
  double* u;           // n
  double* wraw;        // slots x n: row partials of A v (slots = column split, or tile columns when tiled)
  double* cpart;       // NT x n: mirrored (column) partials of A v, tiled mode
  double* tiles;       // tile-major lower triangle of the trailing matrix (tiled mode), 64 x 64 tiles
  double* wpre;        // n
  double* part;        // grid x TD_PART
  double* d; double* e; double* tau;
  unsigned long long* clk;   // [8] per-phase clock totals of CTA 0 (XMCA_SYTRD_TRACE)
  unsigned int* bar;         // arrival counter of this launch's grid barrier (zeroed by the host)
  int slot_t;                // tiled mode: partial sums in ONE transposed array wraw[r * NT + K] (coalesced reads)
  int bar_ra;                // grid barrier by red.release / ld.acquire instead of fence + atomic + fence
  int keep_l2;               // trailing matrix fits the L2: plain loads (stay resident) instead of evict-first
  int64_t bs_a, bs_v, bs_ws; // batched launch: distance of problem 1 behind problem 0 (doubles, doubles, BYTES)
};

// streaming load of the trailing matrix: evict-first while it is larger than the L2 (it is read once per column
// and must not push the panel buffers / partial vectors out), plain once it fits
__device__ __forceinline__ double ld_trail(const double* p, int keep) { return keep ? *p : __ldcs(p); }

__device__ __forceinline__ double block_sum_1024(double v, double* red) {
  // all threads get the sum; red: 32 doubles of shared memory
  v = warp_sum(v);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  double t = red[threadIdx.x & 31];
  t = warp_sum(t);
  return t;
}

// sum over CTAs of part[g * TD_PART + slot]; every thread gets the result
__device__ __forceinline__ double grid_slot_sum(const double* part, int slot, int G, double* red) {
  double s = 0.0;
  for (int g = threadIdx.x; g < G; g += TD_THREADS) s += part[(int64_t)g * TD_PART + slot];
  return block_sum_1024(s, red);
}

// Grid-wide barrier for the co-resident CTAs of a cooperative launch: one monotonically increasing
// arrival counter per launch, release/acquire through __threadfence; about half the latency of
// cooperative_groups' grid.sync(), which is paid three times per column.
__device__ __forceinline__ void grid_barrier(unsigned int* counter, unsigned int& epoch, int release_acquire, int nctas) {
  __syncthreads();
  if (threadIdx.x == 0) {
    epoch += (unsigned)nctas;
    if (release_acquire) {
      // the CTA barrier above orders the other threads' writes before this release (cumulativity)
      asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(counter) : "memory");
      unsigned int seen;
      do {
        asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(counter) : "memory");
      } while (seen < epoch);
    } else {
      __threadfence();
      atomicAdd(counter, 1u);
      while (*((volatile unsigned int*)counter) < epoch) { }
      __threadfence();
    }
  }
  __syncthreads();
}

// tile (I, J), J <= I, of the tile-major lower triangle: 64 x 64 doubles, row-major, contiguous
__device__ __forceinline__ int64_t tile_off(int I, int J) { return ((int64_t)I * (I + 1) / 2 + J) * (TD_TS * TD_TS); }

// sums of 8 per-lane values over the 32 lanes of a warp with 9 shuffles: afterwards every lane l holds the
// total of x[(l >> 2) & 7]
__device__ __forceinline__ double warp_reduce8(double (&x)[8], int lane) {
  double y[4], z[2];
  const bool b4 = lane & 16, b3 = lane & 8, b2 = lane & 4;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const double keep = b4 ? x[k + 4] : x[k], send = b4 ? x[k] : x[k + 4];
    y[k] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
  }
#pragma unroll
  for (int k = 0; k < 2; ++k) {
    const double keep = b3 ? y[k + 2] : y[k], send = b3 ? y[k] : y[k + 2];
    z[k] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
  }
  const double keep = b2 ? z[1] : z[0], send = b2 ? z[0] : z[1];
  double r = keep + __shfl_xor_sync(0xffffffffu, send, 4);
  r += __shfl_xor_sync(0xffffffffu, r, 2);
  r += __shfl_xor_sync(0xffffffffu, r, 1);
  return r;
}

// TILED: the trailing matrix lives in P.tiles (tile-major, lower triangle, diagonal tiles full) and y = A v
// reads every tile ONCE -- each tile feeds the row sums of its block row and, mirrored, the column sums of
// its block column -- half the HBM bytes of the row-major pass, in contiguous 32 KB pieces.  Partial sums go
// to per-tile-row / per-tile-column slots (deterministic).  P.A then only receives the reflectors.
// TWO: two grid barriers per column instead of three -- v^T A v is accumulated during the streaming pass, so
// alpha = -tau^2/2 (v^T A v - 2 (V^T v).(W^T v)) is known right after the second barrier and w is stored at once;
// every CTA recomputes w for the first active row (the only entry the next column needs from another CTA).
template <bool TILED, bool TWO, bool BATCH>
__global__ void __launch_bounds__(TD_THREADS, 1) sytrd_panel_kernel(SytrdParams P) {
  cg::grid_group grid = cg::this_grid();
  extern __shared__ double vs[];                 // current Householder vector (n doubles)
  __shared__ double red[32];
  __shared__ double Vc[TD_NB], Wc[TD_NB];        // row c of V and W
  __shared__ double pv[TD_K2];                   // p1 = V^T v (first 64), p2 = W^T v (last 64)
  __shared__ double psum[16][TD_K2];
  __shared__ double s_wfirst;                    // TWO: w of the first active row of the previous column

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  // BATCH: two independent problems of the same size in one launch, half of the grid each (own barrier counter);
  // while one group streams its trailing matrix the other is in its latency-bound phases.  Problem 1 lives at
  // fixed strides behind problem 0 (matrix, d / e / tau, workspace).
  const int Gdim = BATCH ? (int)(gridDim.x >> 1) : (int)gridDim.x;
  const int grp = BATCH ? (int)(blockIdx.x >= (unsigned)Gdim) : 0;
  const int bx = (int)blockIdx.x - grp * Gdim;
  const int64_t ows = BATCH ? grp * P.bs_ws : 0, oa = BATCH ? grp * P.bs_a : 0, ov = BATCH ? grp * P.bs_v : 0;
  auto wsp = [&](double* q) { return reinterpret_cast<double*>(reinterpret_cast<char*>(q) + ows); };
  double* const qVW = wsp(P.VW); double* const qWV = wsp(P.WV); double* const qu = wsp(P.u);
  double* const qwraw = wsp(P.wraw); double* const qcpart = wsp(P.cpart); double* const qtiles = wsp(P.tiles);
  double* const qwpre = wsp(P.wpre); double* const qpart = wsp(P.part);
  double* const qA = P.A + oa; double* const qd = P.d + ov; double* const qe = P.e + ov; double* const qtau = P.tau + ov;
  unsigned long long* const qclk = P.clk ? reinterpret_cast<unsigned long long*>(reinterpret_cast<char*>(P.clk) + ows) : nullptr;
  unsigned int* const qbar = reinterpret_cast<unsigned int*>(reinterpret_cast<char*>(P.bar) + ows);
  const int G = Gdim;
  const int NW = G * TD_WARPS, gw = bx * TD_WARPS + warp;
  const int n = P.n;
  const int64_t lda = P.lda;
  double alpha2_prev = 0.0;
  unsigned int epoch = 0;
  long long tk[8] = {0, 0, 0, 0, 0, 0, 0, 0};   // [5] reflector set-up, [6] streaming pass, both also inside [2]

  for (int i = 0; i < P.nb; ++i) {
    long long c0 = clock64();
    const int c = P.j0 + i;
    const int n1 = n - c - 1;
    // ------------------------------------------------------------ phase A: column c
    if (tid < i) {
      if (tid == i - 1) { Vc[tid] = 1.0; Wc[tid] = TWO ? s_wfirst : qwpre[c] + alpha2_prev; }
      else { Vc[tid] = qVW[(int64_t)c * TD_K2 + tid]; Wc[tid] = qVW[(int64_t)c * TD_K2 + TD_NB + tid]; }
    }
    __syncthreads();
    double ssq = 0.0;
    {
      // two rows per trip: all loads of both rows are in flight before the first reduction (the phase is
      // pure latency: ~1.7 rows per warp)
      int r = c + ((gw - c % NW) + NW) % NW;      // first row >= c with r % NW == gw
      for (; r < n; r += 2 * NW) {
        const int rb = r + NW;
        const bool hb = rb < n;
        double acr0 = 0.0, acr1 = 0.0;
        if (lane == 0) {
          acr0 = TILED ? qtiles[tile_off(r / TD_TS, c / TD_TS) + (int64_t)(r % TD_TS) * TD_TS + c % TD_TS]
                       : qA[(int64_t)c * lda + r];
          if (hb) acr1 = TILED ? qtiles[tile_off(rb / TD_TS, c / TD_TS) + (int64_t)(rb % TD_TS) * TD_TS + c % TD_TS]
                               : qA[(int64_t)c * lda + rb];
        }
        double acc0 = 0.0, acc1 = 0.0;
        const double* vw0 = qVW + (int64_t)r * TD_K2;
        const double* vw1 = qVW + (int64_t)(hb ? rb : r) * TD_K2;
        for (int t = lane; t < i; t += 32) {
          const double wc = Wc[t], vc = Vc[t];
          acc0 = fma(vw0[t], wc, fma(vw0[TD_NB + t], vc, acc0));
          acc1 = fma(vw1[t], wc, fma(vw1[TD_NB + t], vc, acc1));
        }
        acc0 = warp_sum(acc0); acc1 = warp_sum(acc1);
        if (lane == 0) {
          const double u0 = acr0 - acc0;
          qu[r] = u0;
          if (r == c) qd[c] = u0;
          if (r >= c + 2) ssq = fma(u0, u0, ssq);
          if (hb) {
            const double u1 = acr1 - acc1;
            qu[rb] = u1;
            if (rb >= c + 2) ssq = fma(u1, u1, ssq);
          }
        }
      }
    }
    if (n1 == 0) break;                           // last column: only its diagonal entry
    ssq = block_sum_1024(ssq, red);
    if (tid == 0) qpart[(int64_t)bx * TD_PART] = ssq;
    { long long c1 = clock64(); tk[0] += c1 - c0; c0 = c1; }
    grid_barrier(qbar, epoch, P.bar_ra, G);                   // #1
    { long long c1 = clock64(); tk[1] += c1 - c0; c0 = c1; }

    // ------------------------------------------------------------ phase B: reflector, A v
    // (the loads of u are issued before the norm reduction, whose latency they then share)
    double ureg[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const int j = tid + TD_THREADS * k;
      ureg[k] = (j < n1) ? qu[c + 1 + j] : 0.0;
    }
    const double alpha = qu[c + 1];
    const double xnorm2 = grid_slot_sum(qpart, 0, G, red);
    double beta, tau, scale;
    if (xnorm2 == 0.0) { beta = alpha; tau = 0.0; scale = 0.0; }
    else {
      beta = -copysign(sqrt(fma(alpha, alpha, xnorm2)), alpha);
      tau = (beta - alpha) / beta;
      scale = 1.0 / (alpha - beta);
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const int j = tid + TD_THREADS * k;
      if (j < n1) vs[j] = (j == 0) ? 1.0 : ureg[k] * scale;
    }
    for (int j = tid + 8 * TD_THREADS; j < n1; j += TD_THREADS) vs[j] = qu[c + 1 + j] * scale;
    if (bx == 0 && tid == 0) { qe[c] = beta; qtau[c] = tau; }
    __syncthreads();
    for (int j = bx * TD_THREADS + tid; j < n1; j += G * TD_THREADS)
      qA[(int64_t)c * lda + c + 1 + j] = vs[j];                                       // reflector storage
    long long cb = clock64();
    tk[5] += cb - c0;

    int F = 1;
    double q = 0.0;                               // TWO: this thread's share of v^T (A v)
    const int NTs = (n + TD_TS - 1) / TD_TS;
    if (TILED) {
      const int NT = (n + TD_TS - 1) / TD_TS, IB = (c + 1) / TD_TS, m = NT - IB;
      const int items = m * (m + 1) / 2;
      // v by GLOBAL index, zero outside the active range (this also masks the retired rows / columns of block IB)
      auto vget = [&](int g) -> double { return (g > c && g < n) ? vs[g - c - 1] : 0.0; };
      for (int item = gw; item < items; item += NW) {
        int Ir = (int)((sqrt(8.0 * (double)item + 1.0) - 1.0) * 0.5);
        while (Ir * (Ir + 1) / 2 > item) --Ir;
        while ((Ir + 1) * (Ir + 2) / 2 <= item) ++Ir;
        const int I = IB + Ir, J = IB + (item - Ir * (Ir + 1) / 2);
        const double* tp = qtiles + tile_off(I, J) + lane;
        const bool diag = (I == J);
        const double vc0 = vget(TD_TS * J + lane), vc1 = vget(TD_TS * J + 32 + lane);
        double ca0 = 0.0, ca1 = 0.0;
#pragma unroll 1
        for (int rg = 0; rg < TD_TS / 8; ++rg) {
          double a0[8], a1[8];
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            a0[k] = ld_trail(tp + (rg * 8 + k) * TD_TS, P.keep_l2);
            a1[k] = ld_trail(tp + (rg * 8 + k) * TD_TS + 32, P.keep_l2);
          }
          double x8[8];
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            x8[k] = fma(a0[k], vc0, a1[k] * vc1);
            if (!diag) {
              const double vr = vget(TD_TS * I + rg * 8 + k);
              ca0 = fma(a0[k], vr, ca0);
              ca1 = fma(a1[k], vr, ca1);
            }
          }
          const double rsum = warp_reduce8(x8, lane);
          const int gr = TD_TS * I + rg * 8 + ((lane >> 2) & 7);
          if (!(lane & 3) && gr > c && gr < n) {
            if (P.slot_t) qwraw[(int64_t)gr * NTs + J] = rsum; else qwraw[(int64_t)J * n + gr] = rsum;
            if (TWO) q = fma(vs[gr - c - 1], rsum, q);
          }
        }
        if (!diag) {
          const int g0 = TD_TS * J + lane, g1 = g0 + 32;
          if (g0 > c && g0 < n) { if (P.slot_t) qwraw[(int64_t)g0 * NTs + I] = ca0; else qcpart[(int64_t)I * n + g0] = ca0; }
          if (g1 > c && g1 < n) { if (P.slot_t) qwraw[(int64_t)g1 * NTs + I] = ca1; else qcpart[(int64_t)I * n + g1] = ca1; }
          if (TWO) q = fma(ca0, vc0, fma(ca1, vc1, q));
        }
      }
    } else {
    // column split so that every warp of the grid gets >= ~4 row pieces of >= 256 elements
    while (F < TD_MAXF && (int64_t)n1 * F < 4LL * NW && n1 / (2 * F) >= 256) F *= 2;
    const int len = ((n1 + F - 1) / F + 31) & ~31;
    {
      const int64_t items = (int64_t)n1 * F;
      for (int64_t item = gw; item < items; item += NW) {
        const int rr = (int)(item / F), qq = (int)(item - (int64_t)rr * F);
        const int s0 = qq * len, s1 = min(n1, s0 + len);
        const double* row = qA + (int64_t)(c + 1 + rr) * lda + (c + 1);
        double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0, a4 = 0.0, a5 = 0.0, a6 = 0.0, a7 = 0.0;
        int s = s0 + lane;
        for (; s + 224 < s1; s += 256) {
          // streaming loads (evict-first): the trailing matrix is read once per column and must not
          // push the panel buffers / partial vectors out of the caches
          const int kp = P.keep_l2;
          const double x0 = ld_trail(row + s, kp), x1 = ld_trail(row + s + 32, kp), x2 = ld_trail(row + s + 64, kp), x3 = ld_trail(row + s + 96, kp);
          const double x4 = ld_trail(row + s + 128, kp), x5 = ld_trail(row + s + 160, kp), x6 = ld_trail(row + s + 192, kp), x7 = ld_trail(row + s + 224, kp);
          a0 = fma(x0, vs[s], a0);       a1 = fma(x1, vs[s + 32], a1);
          a2 = fma(x2, vs[s + 64], a2);  a3 = fma(x3, vs[s + 96], a3);
          a4 = fma(x4, vs[s + 128], a4); a5 = fma(x5, vs[s + 160], a5);
          a6 = fma(x6, vs[s + 192], a6); a7 = fma(x7, vs[s + 224], a7);
        }
        for (; s < s1; s += 32) a0 = fma(row[s], vs[s], a0);
        double sum = warp_sum(((a0 + a1) + (a2 + a3)) + ((a4 + a5) + (a6 + a7)));
        if (lane == 0) {
          qwraw[(int64_t)qq * n + c + 1 + rr] = sum;
          if (TWO) q = fma(sum, vs[rr], q);
        }
      }
    }
    }
    { const long long cb2 = clock64(); tk[6] += cb2 - cb; }
    if (i > 0 || TWO) {
      double pa[4] = {0.0, 0.0, 0.0, 0.0};
      int r = c + 1 + ((gw - (c + 1) % NW) + NW) % NW;
      for (; r < n; r += 2 * NW) {                 // two rows per trip (8 independent loads)
        const int rb = r + NW;
        const bool hb = rb < n;
        const double vr0 = vs[r - c - 1], vr1 = hb ? vs[rb - c - 1] : 0.0;
        const double* vw0 = qVW + (int64_t)r * TD_K2;
        const double* vw1 = qVW + (int64_t)(hb ? rb : r) * TD_K2;
        double x0[4], x1[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const int idx = lane + 32 * k;
          const bool on = (idx & (TD_NB - 1)) < i;
          x0[k] = on ? vw0[idx] : 0.0;
          x1[k] = on ? vw1[idx] : 0.0;
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) pa[k] = fma(x0[k], vr0, fma(x1[k], vr1, pa[k]));
      }
      if (TWO) {
        // v^T A v rides in entry 127 of the p vector (p2[63], always zero: a panel has at most 63 earlier columns)
        q = warp_sum(q);
        if (lane == 31) pa[3] = q;
      }
      // CTA reduction of the 32 per-warp partial vectors, 16 warps at a time
      for (int pass = 0; pass < 2; ++pass) {
        __syncthreads();
        if ((warp >> 4) == pass) {
#pragma unroll
          for (int k = 0; k < 4; ++k) psum[warp & 15][lane + 32 * k] = pa[k];
        }
        __syncthreads();
        if (tid < TD_K2) {
          double s0 = 0.0, s1 = 0.0;
#pragma unroll
          for (int w = 0; w < 8; ++w) { s0 += psum[w][tid]; s1 += psum[w + 8][tid]; }
          pv[tid] = (pass == 0 ? 0.0 : pv[tid]) + (s0 + s1);
        }
      }
      __syncthreads();
      if (tid < TD_K2) qpart[(int64_t)bx * TD_PART + 1 + tid] = pv[tid];
    }
    { long long c1 = clock64(); tk[2] += c1 - c0; c0 = c1; }
    grid_barrier(qbar, epoch, P.bar_ra, G);                   // #2
    { long long c1 = clock64(); tk[1] += c1 - c0; c0 = c1; }

    // ------------------------------------------------------------ phase C: w (before the alpha correction)
    if (i > 0 || TWO) {
      const int idx = tid & (TD_K2 - 1), gq = tid >> 7;        // 8 groups of CTAs
      double s = 0.0;
      if ((idx & (TD_NB - 1)) < i || (TWO && idx == TD_K2 - 1)) {   // (entries >= i of p1 / p2 are zero)
        for (int g0 = gq; g0 < G; g0 += 8 * 20) {              // 20 independent loads in flight
          double tmp[20];
#pragma unroll
          for (int k = 0; k < 20; ++k) {
            const int g = g0 + 8 * k;
            tmp[k] = (g < G) ? qpart[(int64_t)g * TD_PART + 1 + idx] : 0.0;
          }
#pragma unroll
          for (int k = 0; k < 20; ++k) s += tmp[k];
        }
      }
      psum[gq][idx] = s;
      __syncthreads();
      if (tid < TD_K2) {
        double t = 0.0;
#pragma unroll
        for (int w = 0; w < 8; ++w) t += psum[w][tid];
        pv[tid] = t;
      }
      __syncthreads();
    }
    // w before the alpha correction for row r: this lane's share (to be summed over the warp, times tau)
    auto row_part = [&](int r) -> double {
      double acc = 0.0;
      const double* vw = qVW + (int64_t)r * TD_K2;
      for (int t = lane; t < i; t += 32) acc = fma(vw[t], pv[TD_NB + t], fma(vw[TD_NB + t], pv[t], acc));
      double wr = 0.0;
      if (TILED) {
        const int IB = (c + 1) / TD_TS, I = r / TD_TS;
        if (P.slot_t) {
          for (int K = IB + lane; K < NTs; K += 32) wr += qwraw[(int64_t)r * NTs + K];         // one contiguous run
        } else {
          for (int J = IB + lane; J <= I; J += 32) wr += qwraw[(int64_t)J * n + r];           // tiles (I, J <= I)
          for (int I2 = I + 1 + lane; I2 < NTs; I2 += 32) wr += qcpart[(int64_t)I2 * n + r];  // mirrored: tiles (I2 > I, I)
        }
      } else if (lane < F) wr = qwraw[(int64_t)lane * n + r];
      return wr - acc;
    };
    double alpha2;
    if (TWO) {
      double pp = 0.0;
      for (int t = lane; t < i; t += 32) pp = fma(pv[t], pv[TD_NB + t], pp);
      pp = warp_sum(pp);
      const double vAv = pv[TD_K2 - 1];
      alpha2 = -0.5 * tau * tau * (vAv - 2.0 * pp);
      // two rows per trip (loads of both in flight before the reductions); warp 0 of every CTA also takes
      // the first active row, whose w the next column needs from shared memory
      int r = c + 1 + ((gw - (c + 1) % NW) + NW) % NW;
      bool extra = (warp == 0);
      while (r < n || extra) {
        int ra, rb = -1;
        bool ea = false, eb = false;
        if (r < n) { ra = r; r += NW; } else { ra = c + 1; ea = true; extra = false; }
        if (r < n) { rb = r; r += NW; } else if (extra) { rb = c + 1; eb = true; extra = false; }
        double pa0 = row_part(ra), pa1 = (rb >= 0) ? row_part(rb) : 0.0;
        pa0 = tau * warp_sum(pa0); pa1 = tau * warp_sum(pa1);
        if (lane == 0) {
          const double v0 = vs[ra - c - 1], w0 = fma(alpha2, v0, pa0);
          if (ea) s_wfirst = w0;
          else {
            qVW[(int64_t)ra * TD_K2 + i] = v0;        qVW[(int64_t)ra * TD_K2 + TD_NB + i] = w0;
            qWV[(int64_t)ra * TD_K2 + i] = w0;        qWV[(int64_t)ra * TD_K2 + TD_NB + i] = v0;
          }
          if (rb >= 0) {
            const double v1 = vs[rb - c - 1], w1 = fma(alpha2, v1, pa1);
            if (eb) s_wfirst = w1;
            else {
              qVW[(int64_t)rb * TD_K2 + i] = v1;      qVW[(int64_t)rb * TD_K2 + TD_NB + i] = w1;
              qWV[(int64_t)rb * TD_K2 + i] = w1;      qWV[(int64_t)rb * TD_K2 + TD_NB + i] = v1;
            }
          }
        }
      }
      { long long c1 = clock64(); tk[3] += c1 - c0; c0 = c1; }
    } else {
    double dotacc = 0.0;
    {
      int r = c + 1 + ((gw - (c + 1) % NW) + NW) % NW;
      for (; r < n; r += 2 * NW) {
        const int rb = r + NW;
        const bool hb = rb < n;
        double p0 = row_part(r), p1 = hb ? row_part(rb) : 0.0;
        p0 = tau * warp_sum(p0); p1 = tau * warp_sum(p1);
        if (lane == 0) {
          qwpre[r] = p0;
          dotacc = fma(p0, vs[r - c - 1], dotacc);
          if (hb) { qwpre[rb] = p1; dotacc = fma(p1, vs[rb - c - 1], dotacc); }
        }
      }
    }
    dotacc = block_sum_1024(dotacc, red);
    if (tid == 0) qpart[(int64_t)bx * TD_PART + 130] = dotacc;
    { long long c1 = clock64(); tk[3] += c1 - c0; c0 = c1; }
    grid_barrier(qbar, epoch, P.bar_ra, G);                   // #3
    { long long c1 = clock64(); tk[1] += c1 - c0; c0 = c1; }

    // ------------------------------------------------------------ phase D: finish w, store panel column i
    const double dot = grid_slot_sum(qpart, 130, G, red);
    alpha2 = -0.5 * tau * dot;
    {
      int r = c + 1 + ((gw - (c + 1) % NW) + NW) % NW;
      for (; r < n; r += NW) {
        if (lane == 0) {
          const double v = vs[r - c - 1];
          const double w = fma(alpha2, v, qwpre[r]);
          qVW[(int64_t)r * TD_K2 + i] = v;          qVW[(int64_t)r * TD_K2 + TD_NB + i] = w;
          qWV[(int64_t)r * TD_K2 + i] = w;          qWV[(int64_t)r * TD_K2 + TD_NB + i] = v;
        }
      }
    }
    }
    alpha2_prev = alpha2;
    __syncwarp();
    __syncthreads();
    { long long c1 = clock64(); tk[4] += c1 - c0; c0 = c1; }
  }
  if (bx == 0 && tid == 0 && qclk)
    for (int q = 0; q < 8; ++q) qclk[q] += (unsigned long long)tk[q];
}

// ------------------------------------------------------------------ tile-major helpers
// row-major symmetric S (both triangles) -> tile-major lower triangle (diagonal tiles full, zero padded)
__global__ void to_tiles_kernel(const double* __restrict__ A, int64_t lda, int n, int NT, double* __restrict__ tiles) {
  const int t = blockIdx.x;                       // linear tile index: I (I + 1) / 2 + J
  int I = (int)((sqrt(8.0 * (double)t + 1.0) - 1.0) * 0.5);
  while (I * (I + 1) / 2 > t) --I;
  while ((I + 1) * (I + 2) / 2 <= t) ++I;
  const int J = t - I * (I + 1) / 2;
  double* tp = tiles + (int64_t)t * (TD_TS * TD_TS);
  for (int e = threadIdx.x; e < TD_TS * TD_TS; e += blockDim.x) {
    const int r = TD_TS * I + (e >> 6), c = TD_TS * J + (e & 63);
    tp[e] = (r < n && c < n) ? A[(int64_t)r * lda + c] : 0.0;
  }
}

// tile-major -> row-major for the trailing block rows / columns >= 64 * I0 (both triangles)
__global__ void from_tiles_kernel(const double* __restrict__ tiles, int I0, int NT, int n, double* __restrict__ A,
                                  int64_t lda) {
  __shared__ double buf[TD_TS][TD_TS + 1];
  const int m = NT - I0, t = blockIdx.x;          // tiles (I0 + Ir, I0 + Jr), Jr <= Ir
  int Ir = (int)((sqrt(8.0 * (double)t + 1.0) - 1.0) * 0.5);
  while (Ir * (Ir + 1) / 2 > t) --Ir;
  while ((Ir + 1) * (Ir + 2) / 2 <= t) ++Ir;
  (void)m;
  const int I = I0 + Ir, J = I0 + (t - Ir * (Ir + 1) / 2);
  const double* tp = tiles + tile_off(I, J);
  for (int e = threadIdx.x; e < TD_TS * TD_TS; e += blockDim.x) {
    const int rr = e >> 6, cc = e & 63;
    const double v = tp[e];
    buf[rr][cc] = v;
    const int r = TD_TS * I + rr, c = TD_TS * J + cc;
    if (r < n && c < n) A[(int64_t)r * lda + c] = v;
  }
  if (I == J) return;
  __syncthreads();
  for (int e = threadIdx.x; e < TD_TS * TD_TS; e += blockDim.x) {      // mirrored tile, coalesced through smem
    const int cc = e >> 6, rr = e & 63;
    const int r = TD_TS * I + rr, c = TD_TS * J + cc;
    if (r < n && c < n) A[(int64_t)c * lda + r] = buf[rr][cc];
  }
}

// Trailing update of the tile-major matrix after a panel: tile(I, J) -= VW_I WV_J^T (= V W^T + W V^T) for
// all tiles I0 <= J <= I.  One CTA of 4 warps per tile: 64 x 64 x 128 on the DMMA pipe (32 x 32 warp tiles),
// operand slabs of 16 k in XOR-swizzled shared memory, the tile itself is one contiguous 32 KB read-modify-write.
constexpr int TU_LD = TD_TS + 4;
__global__ void __launch_bounds__(128)
syr2k_tiles_kernel(const double* __restrict__ VW, const double* __restrict__ WV, int I0, int n, double* __restrict__ tiles) {
  __shared__ double As[16][TU_LD];
  __shared__ double Bs[16][TU_LD];
  const int t = blockIdx.x;
  int Ir = (int)((sqrt(8.0 * (double)t + 1.0) - 1.0) * 0.5);
  while (Ir * (Ir + 1) / 2 > t) --Ir;
  while ((Ir + 1) * (Ir + 2) / 2 <= t) ++Ir;
  const int I = I0 + Ir, J = I0 + (t - Ir * (Ir + 1) / 2);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int gid = lane >> 2, tig = lane & 3;
  const int wm = (warp & 1) * 32, wn = (warp >> 1) * 32;
  double c[4][4][2];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) { c[i][j][0] = 0.0; c[i][j][1] = 0.0; }
  const int kk = tid & 15, mrow = tid >> 4;       // loader: 16 consecutive k of rows mrow + 8 i
  for (int k0 = 0; k0 < TD_K2; k0 += 16) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int rl = mrow + 8 * i;
      const int ra = TD_TS * I + rl, rb = TD_TS * J + rl;
      As[kk][rl ^ ((kk >> 2) & 3)] = ra < n ? VW[(int64_t)ra * TD_K2 + k0 + kk] : 0.0;
      Bs[kk][rl ^ ((kk >> 2) & 3)] = rb < n ? WV[(int64_t)rb * TD_K2 + k0 + kk] : 0.0;
    }
    __syncthreads();
#pragma unroll
    for (int ks = 0; ks < 16; ks += 4) {
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
  double* tp = tiles + tile_off(I, J);
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      double2* q = reinterpret_cast<double2*>(tp + (wm + 8 * i + gid) * TD_TS + wn + 8 * j + 2 * tig);
      double2 d = *q;
      d.x -= c[i][j][0]; d.y -= c[i][j][1];
      *q = d;
    }
}

// ------------------------------------------------------------------ bisection
// SB_LANES lanes per eigenvalue index k (0 = LARGEST, descending output): every level evaluates the Sturm count at
// SB_LANES interior points of the current bracket.  Multi-section shortens the dependent chain (levels) at the price
// of more evaluations in total (9-section: 20 levels x 8 = 160 per eigenvalue, 5-section: 27 x 4 = 108, bisection: 63):
// with the product-form recurrence the kernel is bound by fp64 issue, so FOUR lanes win (n = 8192: 8 lanes 5.9 ms,
// 4 lanes 4.5 ms); the quotient form (latency bound) was tuned with eight.
constexpr int SB_LANES = 4;

__global__ void __launch_bounds__(128)
stebz_kernel(int n, const double* __restrict__ d, const double* __restrict__ e2,
             double gl, double gu, double pivmin, double abstol, double* __restrict__ w) {
  const int gid = blockIdx.x * blockDim.x + threadIdx.x;
  const int k = gid / SB_LANES, sub = gid % SB_LANES;
  const bool live = k < n;
  const int want = n - 1 - (live ? k : n - 1);          // index in ascending order
  const unsigned lane = threadIdx.x & 31, grp = lane & ~(SB_LANES - 1);
  double lo = gl, hi = gu;
  for (int it = 0; it < 80; ++it) {
    const double h = (hi - lo) * (1.0 / (SB_LANES + 1));
    const double x = lo + h * (double)(sub + 1);
    // Sturm count: number of eigenvalues < x
    int cnt = 0;
    double q = d[0] - x;
    cnt += (q < 0.0);
    for (int j = 1; j < n; ++j) {
      if (fabs(q) < pivmin) q = -pivmin;
      q = d[j] - x - e2[j - 1] * fast_rcp(q);
      cnt += (q < 0.0);
    }
    // first sub-point whose count exceeds `want` bounds the eigenvalue from above
    const unsigned above = __ballot_sync(0xffffffffu, cnt > want);
    const unsigned mine = (above >> grp) & ((1u << SB_LANES) - 1u);
    const int first = mine ? (__ffs(mine) - 1) : SB_LANES;       // SB_LANES: above every sub-point
    const double nlo = (first == 0) ? lo : lo + h * (double)first;
    const double nhi = (first == SB_LANES) ? hi : lo + h * (double)(first + 1);
    const bool stalled = !(nlo > lo || nhi < hi) || !(nhi - nlo > 0.0);
    lo = nlo; hi = nhi;
    // converged when the bracket no longer shrinks in floating point (uniform per group: all
    // lanes of a group hold the same lo / hi); other groups of the warp may continue
    const bool done = stalled || (hi - lo) <= fmax(abstol, 2.220446049250313e-16 * fmax(fabs(lo), fabs(hi)));
    if (__all_sync(0xffffffffu, done)) break;
  }
  if (live && sub == 0) w[k] = 0.5 * (lo + hi);
}

// The same search with the Sturm sequence in its PRODUCT form  p_j = (d_j - x) p_{j-1} - e_{j-1}^2 p_{j-2}  (count =
// sign changes): the dependent chain per step is one DFMA instead of a Newton-refined reciprocal plus two more
// operations (~85 clocks), and the recurrence is all there is to this kernel.  The matrix is scaled by s = 1 / ||T||
// (es = e^2 s^2 precomputed, d s - x s folded into one FMA), so a step grows the pair by at most 3x; every 8 steps both
// values are renormalised by the larger binary exponent (overflow impossible, underflow only for a value ~1e-300 below
// its partner, which then no longer matters).  A value below 1e-290 of its predecessor is replaced by -1e-290 p_{j-1}
// (LAPACK's pivmin rule in product form).  Brackets / tolerances are the scaled ones; the result is scaled back.
__global__ void __launch_bounds__(128)
stebz_prod_kernel(int n, const double2* __restrict__ pk, double d0s, double tn,
                  double gl, double gu, double abstol, double* __restrict__ w) {
  // pk[j] = (d_j s, e_{j-1}^2 s^2) for j = 1 .. n - 1 (one 16-byte warp-uniform load per step); d0s = d_0 s
  const int gid = blockIdx.x * blockDim.x + threadIdx.x;
  const int k = gid / SB_LANES, sub = gid % SB_LANES;
  const bool live = k < n;
  const int want = n - 1 - (live ? k : n - 1);          // index in ascending order
  const unsigned lane = threadIdx.x & 31, grp = lane & ~(SB_LANES - 1);
  constexpr double TINY = 1e-290;
  double lo = gl, hi = gu;                              // scaled units
  for (int it = 0; it < 80; ++it) {
    const double h = (hi - lo) * (1.0 / (SB_LANES + 1));
    const double x = lo + h * (double)(sub + 1);
    unsigned cnt = 0;
    double p0 = 1.0, p1 = d0s - x;
    cnt += (unsigned)__double2hiint(p1) >> 31;
    // one step: three fp64 operations (the kernel is bound by fp64 issue, not by the chain).  The pivmin rule is
    // only DETECTED here, on the exponent fields with integer instructions; a chunk in which it fired (rare) is
    // redone from its saved state with the rule applied (`careful`).
    unsigned bad = 0;
    auto step = [&](const double2 v) {
      const double t = v.x - x;
      const double pn = fma(t, p1, -v.y * p0);
      const unsigned hn = (unsigned)__double2hiint(pn), h1 = (unsigned)__double2hiint(p1);
      bad |= (unsigned)((hn & 0x7ff00000u) + (963u << 20) < (h1 & 0x7ff00000u));      // |pn| < ~1e-290 |p1|
      cnt += (hn ^ h1) >> 31;
      p0 = p1;
      p1 = pn;
    };
    auto careful = [&](const double2 v) {
      const double t = v.x - x;
      double pn = fma(t, p1, -v.y * p0);
      if (fabs(pn) < TINY * fabs(p1)) pn = -TINY * p1;
      cnt += ((unsigned)__double2hiint(pn) ^ (unsigned)__double2hiint(p1)) >> 31;
      p0 = p1;
      p1 = pn;
    };
    auto renorm = [&]() {                                 // by the larger binary exponent of the pair
      const int e1 = (__double2hiint(p1) >> 20) & 0x7ff, e0 = (__double2hiint(p0) >> 20) & 0x7ff;
      const int em = max(max(e1, e0), 1);
      const double sc = __hiloint2double((2046 - em) << 20, 0);        // 2^(1023 - em): em in [1, 2045]
      p1 *= sc;
      p0 *= sc;
    };
    // chunks of 8 steps in two register buffers (A, B): the operands of the next chunk are loaded before the 8
    // dependent steps of the current one
    int j = 1;
    double2 va[8], vb[8];
    auto load = [&](double2 (&v)[8], int j0) {
#pragma unroll
      for (int u = 0; u < 8; ++u) v[u] = __ldg(pk + j0 + u);
    };
    auto run = [&](const double2 (&v)[8]) {
      const double s0 = p0, s1 = p1;
      const unsigned c0 = cnt;
#pragma unroll
      for (int u = 0; u < 8; ++u) step(v[u]);
      if (bad) {
        p0 = s0; p1 = s1; cnt = c0; bad = 0;
#pragma unroll
        for (int u = 0; u < 8; ++u) careful(v[u]);       // (static indices: the buffers must stay in registers)
      }
      renorm();
    };
    if (j + 8 <= n) load(va, j);
    while (j + 8 <= n) {
      if (j + 16 <= n) load(vb, j + 8);
      run(va);
      j += 8;
      if (j + 8 > n) break;
      if (j + 16 <= n) load(va, j + 8);
      run(vb);
      j += 8;
    }
    if (j < n) {
      for (; j < n; ++j) careful(__ldg(pk + j));
      renorm();
    }
    // first sub-point whose count exceeds `want` bounds the eigenvalue from above
    const unsigned above = __ballot_sync(0xffffffffu, (int)cnt > want);
    const unsigned mine = (above >> grp) & ((1u << SB_LANES) - 1u);
    const int first = mine ? (__ffs(mine) - 1) : SB_LANES;       // SB_LANES: above every sub-point
    const double nlo = (first == 0) ? lo : lo + h * (double)first;
    const double nhi = (first == SB_LANES) ? hi : lo + h * (double)(first + 1);
    const bool stalled = !(nlo > lo || nhi < hi) || !(nhi - nlo > 0.0);
    lo = nlo; hi = nhi;
    const bool done = stalled || (hi - lo) <= fmax(abstol, 2.220446049250313e-16 * fmax(fabs(lo), fabs(hi)));
    if (__all_sync(0xffffffffu, done)) break;
  }
  if (live && sub == 0) w[k] = 0.5 * (lo + hi) * tn;
}

// pk[j] = (d_j s, e_{j-1}^2 s^2), j = 1 .. n - 1 (pk[0] unused)
__global__ void stebz_pack_kernel(int n, const double* __restrict__ d, const double* __restrict__ e, double s,
                                  double2* __restrict__ pk) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= 1 && j < n) pk[j] = make_double2(d[j] * s, e[j - 1] * e[j - 1] * s * s);
}

__global__ void square_kernel(int n, const double* __restrict__ e, double* __restrict__ e2) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j < n) e2[j] = e[j] * e[j];
}

__global__ void tridiag_bounds_kernel(int n, const double* __restrict__ d, const double* __restrict__ e,
                                      double* __restrict__ out) {
  // out[0] = Gershgorin lower, out[1] = upper, out[2] = max e^2, out[3] = d[0]   (single CTA)
  __shared__ double red[3][32];
  double lo = INFINITY, hi = -INFINITY, e2 = 0.0;
  for (int j = threadIdx.x; j < n; j += blockDim.x) {
    const double a = (j > 0) ? fabs(e[j - 1]) : 0.0, b = (j < n - 1) ? fabs(e[j]) : 0.0;
    lo = fmin(lo, d[j] - a - b);
    hi = fmax(hi, d[j] + a + b);
    e2 = fmax(e2, b * b);
  }
  for (int o = 16; o > 0; o >>= 1) {
    lo = fmin(lo, __shfl_xor_sync(0xffffffffu, lo, o));
    hi = fmax(hi, __shfl_xor_sync(0xffffffffu, hi, o));
    e2 = fmax(e2, __shfl_xor_sync(0xffffffffu, e2, o));
  }
  if ((threadIdx.x & 31) == 0) { red[0][threadIdx.x >> 5] = lo; red[1][threadIdx.x >> 5] = hi; red[2][threadIdx.x >> 5] = e2; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int i = 1; i < (int)(blockDim.x >> 5); ++i) {
      lo = fmin(lo, red[0][i]); hi = fmax(hi, red[1][i]); e2 = fmax(e2, red[2][i]);
    }
    out[0] = lo; out[1] = hi; out[2] = e2; out[3] = d[0];
  }
}

// ---------------------------------------------------------- inverse iteration
// One CTA per cluster of eigenvalues (clusters processed member after member, MGS
// against the earlier members).  Thread 0 runs the serial tridiagonal LU / solves.
constexpr int ST_THREADS = 256;

__device__ __forceinline__ double block_sum_256(double v, double* red) {
  v = warp_sum(v);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  double t = (threadIdx.x & 31) < (ST_THREADS / 32) ? red[threadIdx.x & 31] : 0.0;
  return warp_sum(t);
}

__device__ __forceinline__ double hash_unit(uint32_t a, uint32_t b) {
  uint32_t x = a * 0x9E3779B1u ^ (b + 0x7F4A7C15u) * 0x85EBCA77u;
  x ^= x >> 15; x *= 0x2C1B3C6Du; x ^= x >> 12; x *= 0x297A2D39u; x ^= x >> 15;
  return ((double)x + 0.5) * (1.0 / 4294967296.0) - 0.5;
}

constexpr int ST_CH = 512;                 // chunk of the tridiagonal staged in shared memory

// The LU factorisation and the two triangular solves are first-order recurrences: thread 0 runs
// them on chunks that all threads stage through shared memory (coalesced global traffic, no
// global-memory latency inside the dependent chain, pivots stored as reciprocals).
__global__ void __launch_bounds__(ST_THREADS)
stein_kernel(int n, const double* __restrict__ d, const double* __restrict__ e,
             const double* __restrict__ lam, const int* __restrict__ cluster_start, int n_clusters,
             double tnorm, int iters, double* __restrict__ Z, int64_t ldz, double* __restrict__ work) {
  __shared__ double red[32];
  __shared__ double buf[6][ST_CH + 2];
  const int cl = blockIdx.x;
  if (cl >= n_clusters) return;
  const int k0 = cluster_start[cl], k1 = cluster_start[cl + 1];
  const int tid = threadIdx.x;
  double* u0i = work + (int64_t)cl * 5 * n;   // reciprocal pivots
  double* u1 = u0i + n;
  double* u2 = u1 + n;
  double* lm = u2 + n;                        // multipliers
  double* sw = lm + n;                        // 0 / 1 row-swap flags
  const double tiny = 2.220446049250313e-16 * tnorm + 1e-300;
  const int n_it = (k1 - k0 > 1) ? iters + 1 : iters;

  double prev_used = 0.0;
  for (int k = k0; k < k1; ++k) {
    // separate numerically identical eigenvalues a little (dstein), keeps the solves distinct;
    // eigenvalues are descending inside a cluster
    double lk = lam[k];
    if (k > k0) {
      const double pert = 10.0 * 2.220446049250313e-16 * fabs(lk) + tiny;
      if (prev_used - lk < pert) lk = prev_used - pert;
    }
    prev_used = lk;
    double* x = Z + (int64_t)k * ldz;

    // ---- pivoted LU of T - lk I (rows j = 0 .. n-2 eliminate; row n-1 closes)
    double ca = d[0] - lk, cb = (n > 1) ? e[0] : 0.0;          // carried by thread 0
    for (int base = 0; base < n - 1; base += ST_CH) {
      const int len = min(ST_CH, n - 1 - base);
      for (int t = tid; t <= len; t += ST_THREADS) {
        const int j = base + t;
        buf[0][t] = (j + 1 < n) ? d[j + 1] - lk : 0.0;         // next diagonal
        const double ej = (j < n - 1) ? e[j] : 0.0;
        buf[1][t] = ej;                                        // e[j]   (t + 1 -> e[j + 1])
        buf[2][t] = (ej != 0.0) ? 1.0 / ej : 0.0;
      }
      __syncthreads();
      if (tid == 0) {
        for (int t = 0; t < len; ++t) {
          const double cj = buf[1][t], an = buf[0][t], bn = buf[1][t + 1];
          if (fabs(ca) >= fabs(cj)) {
            double piv = ca;
            if (fabs(piv) < tiny) piv = copysign(tiny, piv);
            const double ip = fast_rcp(piv);
            const double m = cj * ip;
            buf[3][t] = m; buf[4][t] = 0.0;                    // lm, sw
            buf[0][t] = ip; buf[5][t] = cb;                    // u0i, u1   (u2 = 0)
            buf[2][t] = 0.0;
            ca = fma(-m, cb, an); cb = bn;
          } else {
            const double m = ca * buf[2][t];
            buf[3][t] = m; buf[4][t] = 1.0;
            buf[0][t] = buf[2][t]; buf[5][t] = an;
            buf[2][t] = bn;                                    // u2
            ca = fma(-m, an, cb); cb = -m * bn;
          }
        }
      }
      __syncthreads();
      for (int t = tid; t < len; t += ST_THREADS) {
        const int j = base + t;
        u0i[j] = buf[0][t]; u1[j] = buf[5][t]; u2[j] = buf[2][t]; lm[j] = buf[3][t]; sw[j] = buf[4][t];
      }
      __syncthreads();
    }
    if (tid == 0) {
      if (fabs(ca) < tiny) ca = copysign(tiny, ca);
      u0i[n - 1] = 1.0 / ca; u1[n - 1] = 0.0; u2[n - 1] = 0.0;
    }
    for (int j = tid; j < n; j += ST_THREADS) x[j] = hash_unit((uint32_t)j, (uint32_t)k);
    __threadfence_block();
    __syncthreads();

    for (int it = 0; it < n_it; ++it) {
      if (it > 0) {               // first pass: the random vector stands for L^-1 P b (dstein)
        double cur = x[0];
        for (int base = 0; base < n - 1; base += ST_CH) {
          const int len = min(ST_CH, n - 1 - base);
          for (int t = tid; t < len; t += ST_THREADS) {
            buf[0][t] = x[base + t + 1]; buf[1][t] = lm[base + t]; buf[2][t] = sw[base + t];
          }
          __syncthreads();
          if (tid == 0) {
            for (int t = 0; t < len; ++t) {
              double yj = cur, yn = buf[0][t];
              if (buf[2][t] != 0.0) { const double tmp = yj; yj = yn; yn = tmp; }
              buf[3][t] = yj;
              cur = fma(-buf[1][t], yj, yn);
            }
            buf[4][0] = cur;
          }
          __syncthreads();
          cur = buf[4][0];
          for (int t = tid; t < len; t += ST_THREADS) x[base + t] = buf[3][t];
          __syncthreads();
        }
        if (tid == 0) x[n - 1] = cur;
        __threadfence_block();
        __syncthreads();
      }
      {
        double xp1 = 0.0, xp2 = 0.0;
        for (int top = n; top > 0; top -= ST_CH) {
          const int base = max(0, top - ST_CH), len = top - base;
          for (int t = tid; t < len; t += ST_THREADS) {
            buf[0][t] = x[base + t]; buf[1][t] = u0i[base + t]; buf[2][t] = u1[base + t]; buf[3][t] = u2[base + t];
          }
          __syncthreads();
          if (tid == 0) {
            for (int t = len - 1; t >= 0; --t) {
              const double v = fma(-buf[3][t], xp2, fma(-buf[2][t], xp1, buf[0][t])) * buf[1][t];
              buf[0][t] = v; xp2 = xp1; xp1 = v;
            }
            buf[4][0] = xp1; buf[4][1] = xp2;
          }
          __syncthreads();
          xp1 = buf[4][0]; xp2 = buf[4][1];
          for (int t = tid; t < len; t += ST_THREADS) x[base + t] = buf[0][t];
          __syncthreads();
        }
      }
      __threadfence_block();
      __syncthreads();
      // scale down first (the solve amplifies by up to 1/eps), then MGS inside the cluster
      double mx = 0.0;
      for (int j = tid; j < n; j += ST_THREADS) mx = fmax(mx, fabs(x[j]));
      mx = warp_max(mx);
      __syncthreads();
      if ((tid & 31) == 0) red[tid >> 5] = mx;
      __syncthreads();
      mx = 0.0;
      for (int w = 0; w < ST_THREADS / 32; ++w) mx = fmax(mx, red[w]);
      const double inv = mx > 0.0 ? 1.0 / mx : 1.0;
      for (int j = tid; j < n; j += ST_THREADS) x[j] *= inv;
      __syncthreads();
      for (int p = k0; p < k; ++p) {
        const double* z = Z + (int64_t)p * ldz;
        double dt = 0.0;
        for (int j = tid; j < n; j += ST_THREADS) dt = fma(z[j], x[j], dt);
        dt = block_sum_256(dt, red);
        for (int j = tid; j < n; j += ST_THREADS) x[j] = fma(-dt, z[j], x[j]);
        __syncthreads();
      }
      double ss = 0.0;
      for (int j = tid; j < n; j += ST_THREADS) ss = fma(x[j], x[j], ss);
      ss = block_sum_256(ss, red);
      const double rn = ss > 0.0 ? rsqrt(ss) : 0.0;
      for (int j = tid; j < n; j += ST_THREADS) x[j] *= rn;
      __threadfence_block();
      __syncthreads();
    }
  }
}

// ------------------------------------------------------- back-transformation
// Z[k, :] <- Q Z[k, :],  Q = H(0) H(1) ... H(n-2), reflector c stored in A[c, c+1:n].
// One CTA per vector, vector kept in shared memory.
constexpr int OR_THREADS = 512;
constexpr int OR_REG = 16;                 // reflector elements per thread kept in registers (n <= 8192), the
                                           // rest (longer reflectors) is read from global memory directly

// The n - 1 reflectors are strictly sequential; what a step pays for is latency: the L2 read of the
// reflector row, one block reduction, one barrier.  The NEXT reflector row is therefore prefetched
// into registers while the current one is reduced and applied.
__global__ void __launch_bounds__(OR_THREADS)
ormtr_kernel(int n, const double* __restrict__ A, int64_t lda, const double* __restrict__ tau,
             double* __restrict__ Z, int64_t ldz) {
  extern __shared__ double zs[];
  __shared__ double red[2][OR_THREADS / 32];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  double* z = Z + (int64_t)blockIdx.x * ldz;
  for (int j = tid; j < n; j += OR_THREADS) zs[j] = z[j];
  __syncthreads();
  int par = 0;
  double vn[OR_REG];                         // prefetched head of reflector c (elements tid + 512 q)
  {
    const int c = n - 2;
    const double* v = A + (int64_t)c * lda + (c + 1);
#pragma unroll
    for (int q = 0; q < OR_REG; ++q) { const int j = tid + OR_THREADS * q; vn[q] = (c >= 0 && j < n - c - 1) ? v[j] : 0.0; }
  }
  for (int c = n - 2; c >= 0; --c) {
    const double* v = A + (int64_t)c * lda + (c + 1);
    const int n1 = n - c - 1;
    double vc[OR_REG];
#pragma unroll
    for (int q = 0; q < OR_REG; ++q) vc[q] = vn[q];
    if (c > 0) {                             // prefetch reflector c - 1
      const double* v2 = A + (int64_t)(c - 1) * lda + c;
#pragma unroll
      for (int q = 0; q < OR_REG; ++q) { const int j = tid + OR_THREADS * q; vn[q] = (j < n1 + 1) ? v2[j] : 0.0; }
    }
    const double tc = tau[c];
    if (tc == 0.0) continue;                 // uniform
    double dt = 0.0;
#pragma unroll
    for (int q = 0; q < OR_REG; ++q) { const int j = tid + OR_THREADS * q; if (j < n1) dt = fma(vc[q], zs[c + 1 + j], dt); }
    for (int j = tid + OR_THREADS * OR_REG; j < n1; j += OR_THREADS) dt = fma(v[j], zs[c + 1 + j], dt);
    dt = warp_sum(dt);
    if (lane == 0) red[par][warp] = dt;
    __syncthreads();
    double tot = 0.0;
#pragma unroll
    for (int w = 0; w < OR_THREADS / 32; ++w) tot += red[par][w];
    const double f = -tc * tot;
#pragma unroll
    for (int q = 0; q < OR_REG; ++q) { const int j = tid + OR_THREADS * q; if (j < n1) zs[c + 1 + j] = fma(f, vc[q], zs[c + 1 + j]); }
    for (int j = tid + OR_THREADS * OR_REG; j < n1; j += OR_THREADS) zs[c + 1 + j] = fma(f, v[j], zs[c + 1 + j]);
    par ^= 1;
    __syncthreads();
  }
  for (int j = tid; j < n; j += OR_THREADS) z[j] = zs[j];
}

static int sytrd_grid(size_t smem, int* grid_out) {
  int occ = 0;
  XMCA_CUDA(cudaFuncSetAttribute(sytrd_panel_kernel<false, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  XMCA_CUDA(cudaFuncSetAttribute(sytrd_panel_kernel<true, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  XMCA_CUDA(cudaFuncSetAttribute(sytrd_panel_kernel<false, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  XMCA_CUDA(cudaFuncSetAttribute(sytrd_panel_kernel<true, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  XMCA_CUDA(cudaFuncSetAttribute(sytrd_panel_kernel<false, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  XMCA_CUDA(cudaFuncSetAttribute(sytrd_panel_kernel<true, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  XMCA_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, sytrd_panel_kernel<true, true, true>, TD_THREADS, smem));
  if (occ < 1) return fail(XMCA_CUDA_ERROR, "sytrd panel kernel does not fit on an SM", __FILE__, __LINE__);
  *grid_out = sm_count();
  return XMCA_OK;
}

static size_t al256(size_t b) { return (b + 255) / 256 * 256; }

}  // namespace xmca

using namespace xmca;

extern "C" int64_t xmca_sytrd_max_n(void) { return 26000; }

// tile-major mode is used for the leading panels of large problems (trailing size > TD_TILE_MIN)
static int64_t sytrd_tile_min() {
  static const char* e = getenv("XMCA_SYTRD_TILE_MIN");      // (tuning knob)
  static const int64_t v = e ? atoll(e) / TD_TS * TD_TS : TD_TILE_MIN;
  return v < 2 * TD_TS ? 2 * TD_TS : v;
}
static bool sytrd_tiled(int64_t n) {
  static const bool off = getenv("XMCA_SYTRD_NO_TILES") != nullptr;
  return !off && n >= sytrd_tile_min() + 8 * TD_TS;
}
static int64_t sytrd_nt(int64_t n) { return (n + TD_TS - 1) / TD_TS; }
static size_t sytrd_wraw_slots(int64_t n) { return sytrd_tiled(n) ? (size_t)(sytrd_nt(n) > TD_MAXF ? sytrd_nt(n) : TD_MAXF) : TD_MAXF; }

extern "C" size_t xmca_sytrd_workspace_bytes(int64_t n) {
  size_t b = 0;
  b += 2 * al256((size_t)n * TD_K2 * 8);          // VW, WV
  b += al256((size_t)n * 8) * 2;                  // u, wpre
  b += al256((size_t)n * sytrd_wraw_slots(n) * 8);   // row partials
  if (sytrd_tiled(n)) {
    b += al256((size_t)n * sytrd_nt(n) * 8);                                               // mirrored partials
    b += al256((size_t)(sytrd_nt(n) * (sytrd_nt(n) + 1) / 2) * TD_TS * TD_TS * 8);         // tile-major lower triangle
  }
  b += al256((size_t)(148 * 2) * TD_PART * 8);    // partials
  b += 256;                                       // phase clocks
  b += al256((size_t)(n / TD_NB + 2) * 4);       // one barrier counter per panel launch
  return b;
}

// batch = 1 or 2 problems of the same size; problem 1 at fixed strides behind problem 0
static int sytrd_impl(int64_t n, int batch, double* d_A, int64_t lda, int64_t stride_a, double* d_d, double* d_e,
                      double* d_tau, int64_t stride_v, void* d_workspace, size_t workspace_bytes, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  const size_t smem = (size_t)n * 8;
  int grid = 0;
  int rc = sytrd_grid(smem, &grid);
  if (rc != XMCA_OK) return rc;
  XMCA_REQUIRE(grid <= 148 * 2, "xmca_sytrd: grid larger than the workspace plan");
  if (batch == 2) grid &= ~1;                      // two equal groups of CTAs
  const size_t ws_one = xmca_sytrd_workspace_bytes(n);

  char* ws = reinterpret_cast<char*>(d_workspace);
  SytrdParams P;
  P.A = d_A; P.lda = lda; P.n = (int)n;
  P.bs_a = stride_a; P.bs_v = stride_v; P.bs_ws = (int64_t)ws_one;
  size_t o = 0;
  P.VW = reinterpret_cast<double*>(ws + o); o += al256((size_t)n * TD_K2 * 8);
  P.WV = reinterpret_cast<double*>(ws + o); o += al256((size_t)n * TD_K2 * 8);
  P.u = reinterpret_cast<double*>(ws + o); o += al256((size_t)n * 8);
  P.wpre = reinterpret_cast<double*>(ws + o); o += al256((size_t)n * 8);
  P.wraw = reinterpret_cast<double*>(ws + o); o += al256((size_t)n * sytrd_wraw_slots(n) * 8);
  P.cpart = nullptr; P.tiles = nullptr;
  const bool tiled_mode = sytrd_tiled(n);
  const int NT = (int)sytrd_nt(n);
  if (tiled_mode) {
    P.cpart = reinterpret_cast<double*>(ws + o); o += al256((size_t)n * NT * 8);
    P.tiles = reinterpret_cast<double*>(ws + o); o += al256((size_t)(NT * (NT + 1) / 2) * TD_TS * TD_TS * 8);
  }
  P.part = reinterpret_cast<double*>(ws + o); o += al256((size_t)(148 * 2) * TD_PART * 8);
  P.clk = reinterpret_cast<unsigned long long*>(ws + o); o += 256;
  unsigned int* bars = reinterpret_cast<unsigned int*>(ws + o);
  const size_t off_clk = (size_t)(reinterpret_cast<char*>(P.clk) - ws);
  for (int b = 0; b < batch; ++b) {
    XMCA_CUDA(cudaMemsetAsync(ws + b * ws_one + off_clk, 0, 256 + (size_t)(n / TD_NB + 2) * 4, st));   // clocks + barrier counters
    XMCA_CUDA(cudaMemsetAsync(d_tau + b * stride_v, 0, (size_t)n * 8, st));
  }
  P.d = d_d; P.e = d_e; P.tau = d_tau;
  auto ws_of = [&](double* q, int b) { return reinterpret_cast<double*>(reinterpret_cast<char*>(q) + (size_t)b * ws_one); };

  // XMCA_SYTRD_VARIANT (bit mask, for A/B measurements): 1 = two barriers per column, 2 = transposed slot array,
  // 4 = release/acquire grid barrier
  const char* var_env = getenv("XMCA_SYTRD_VARIANT");
  const int variant = (var_env ? atoi(var_env) : TD_DEFAULT_VARIANT) | (batch == 2 ? 1 : 0);
  const bool two = variant & 1;
  P.slot_t = (variant & 2) ? 1 : 0;
  P.bar_ra = (variant & 4) ? 1 : 0;
  bool in_tiles = false;
  if (tiled_mode) {
    for (int b = 0; b < batch; ++b) {
      to_tiles_kernel<<<(unsigned)(NT * (NT + 1) / 2), 256, 0, st>>>(d_A + b * stride_a, lda, (int)n, NT, ws_of(P.tiles, b));
      XMCA_LAUNCHED();
    }
    in_tiles = true;
  }
  for (int64_t j0 = 0; j0 < n; j0 += TD_NB) {
    const int nb = (int)((n - j0 < TD_NB) ? (n - j0) : TD_NB);
    if (in_tiles && n - j0 <= sytrd_tile_min()) {
      // the trailing matrix now fits the L2: back to the row-major layout (both triangles) for the rest
      const int I0 = (int)(j0 / TD_TS), m = NT - I0;
      for (int b = 0; b < batch; ++b) {
        from_tiles_kernel<<<(unsigned)(m * (m + 1) / 2), 256, 0, st>>>(ws_of(P.tiles, b), I0, NT, (int)n, d_A + b * stride_a, lda);
        XMCA_LAUNCHED();
      }
      in_tiles = false;
    }
    P.j0 = (int)j0; P.nb = nb;
    P.bar = bars + j0 / TD_NB;
    {
      const double m = (double)(n - j0), bytes = (in_tiles ? m * m * 4.0 : m * m * 8.0) * batch;
      static const char* ke = getenv("XMCA_SYTRD_KEEP_MB");  // (tuning knob)
      P.keep_l2 = bytes <= (ke ? atof(ke) : TD_KEEP_MB) * 1e6;
    }
    // (only the last panel can be short, and it has no trailing block to update)
    void* args[] = {&P};
    const void* fn;
    if (batch == 2) fn = in_tiles ? (const void*)sytrd_panel_kernel<true, true, true> : (const void*)sytrd_panel_kernel<false, true, true>;
    else fn = two ? (in_tiles ? (const void*)sytrd_panel_kernel<true, true, false> : (const void*)sytrd_panel_kernel<false, true, false>)
                  : (in_tiles ? (const void*)sytrd_panel_kernel<true, false, false> : (const void*)sytrd_panel_kernel<false, false, false>);
    XMCA_CUDA(cudaLaunchCooperativeKernel(fn, dim3(grid), dim3(TD_THREADS), args, smem, st));
    XMCA_LAUNCHED();
    const int64_t r0 = j0 + nb;
    if (r0 < n) {
      for (int b = 0; b < batch; ++b) {
        if (in_tiles) {
          const int I0 = (int)(r0 / TD_TS), m = NT - I0;
          syr2k_tiles_kernel<<<(unsigned)(m * (m + 1) / 2), 128, 0, st>>>(ws_of(P.VW, b), ws_of(P.WV, b), I0, (int)n, ws_of(P.tiles, b));
          XMCA_LAUNCHED();
        } else {
          // trailing update  A22 -= V W^T + W V^T  =  [V | W] [W | V]^T   (dsyr2k, full square kept)
          const int64_t m = n - r0;
          rc = xmca_gemm_ex(1, 1, m, m, TD_K2, -1.0, ws_of(P.VW, b) + r0 * TD_K2, XMCA_F64, TD_K2, ws_of(P.WV, b) + r0 * TD_K2,
                            XMCA_F64, TD_K2, d_A + b * stride_a + r0 * lda + r0, XMCA_F64, lda, 1, XMCA_F64, 1, nullptr, 0,
                            XMCA_GEMM_SYMMETRIC, stream);
          if (rc != XMCA_OK) return rc;
        }
      }
    }
  }
  if (getenv("XMCA_SYTRD_TRACE")) {
    unsigned long long h[8];
    XMCA_CUDA(cudaMemcpyAsync(h, P.clk, sizeof h, cudaMemcpyDeviceToHost, st));
    XMCA_CUDA(cudaStreamSynchronize(st));
    fprintf(stderr, "[xmca sytrd] n=%lld batch=%d clocks (CTA 0): column update %.3e | grid syncs %.3e | reflector+symv+p %.3e | w %.3e | store %.3e || inside symv: set-up %.3e, streaming %.3e\n",
            (long long)n, batch, (double)h[0], (double)h[1], (double)h[2], (double)h[3], (double)h[4], (double)h[5], (double)h[6]);
  }
  (void)workspace_bytes;
  return XMCA_OK;
}

extern "C" int xmca_sytrd(int64_t n, double* d_A, int64_t lda, double* d_d, double* d_e, double* d_tau,
                          void* d_workspace, size_t workspace_bytes, void* stream) {
  XMCA_REQUIRE(n >= 1 && d_A && d_d && d_e && d_tau && d_workspace, "xmca_sytrd: bad argument");
  XMCA_REQUIRE(lda >= n, "xmca_sytrd: lda < n");
  XMCA_REQUIRE(n <= xmca_sytrd_max_n(), "xmca_sytrd: n too large for the shared-memory Householder vector");
  XMCA_REQUIRE(workspace_bytes >= xmca_sytrd_workspace_bytes(n), "xmca_sytrd: workspace too small");
  return sytrd_impl(n, 1, d_A, lda, 0, d_d, d_e, d_tau, 0, d_workspace, workspace_bytes, stream);
}

extern "C" int xmca_sytrd_batched(int64_t n, int batch, double* d_A, int64_t lda, int64_t stride_a, double* d_d,
                                  double* d_e, double* d_tau, int64_t stride_v, void* d_workspace,
                                  size_t workspace_bytes, void* stream) {
  XMCA_REQUIRE(n >= 1 && d_A && d_d && d_e && d_tau && d_workspace, "xmca_sytrd_batched: bad argument");
  XMCA_REQUIRE(batch == 1 || batch == 2, "xmca_sytrd_batched: batch must be 1 or 2");
  XMCA_REQUIRE(lda >= n, "xmca_sytrd_batched: lda < n");
  XMCA_REQUIRE(batch == 1 || (stride_a >= n * lda && stride_v >= n), "xmca_sytrd_batched: strides too small");
  XMCA_REQUIRE(n <= xmca_sytrd_max_n(), "xmca_sytrd_batched: n too large for the shared-memory Householder vector");
  XMCA_REQUIRE(workspace_bytes >= (size_t)batch * xmca_sytrd_workspace_bytes(n), "xmca_sytrd_batched: workspace too small");
  return sytrd_impl(n, batch, d_A, lda, stride_a, d_d, d_e, d_tau, stride_v, d_workspace, workspace_bytes, stream);
}

extern "C" int xmca_stebz(int64_t n, const double* d_d, const double* d_e, double* d_w, double* d_scratch,
                          void* stream) {
  XMCA_REQUIRE(n >= 1 && d_d && d_e && d_w && d_scratch, "xmca_stebz: bad argument");   // d_scratch: 2 n + 8 doubles
  cudaStream_t st = (cudaStream_t)stream;
  tridiag_bounds_kernel<<<1, 1024, 0, st>>>((int)n, d_d, d_e, d_scratch);
  XMCA_LAUNCHED();
  double h[4];
  XMCA_CUDA(cudaMemcpyAsync(h, d_scratch, sizeof h, cudaMemcpyDeviceToHost, st));
  XMCA_CUDA(cudaStreamSynchronize(st));
  if (!isfinite(h[0]) || !isfinite(h[1]))
    return fail(XMCA_NUMERIC, "xmca_stebz: non-finite tridiagonal (SVD failed. NaN entries may be the problem.)",
                __FILE__, __LINE__);
  const double tn = fmax(fabs(h[0]), fabs(h[1]));
  const double gl = h[0] - 2.0 * 2.220446049250313e-16 * tn * (double)n - 1e-300;
  const double gu = h[1] + 2.0 * 2.220446049250313e-16 * tn * (double)n + 1e-300;
  const double pivmin = 2.2250738585072014e-308 * fmax(1.0, h[2]);
  double* d_e2 = d_scratch + 8;
  const int64_t threads = n * SB_LANES;
  const char* mode = getenv("XMCA_STEBZ");               // "quot": the quotient-form recurrence (A/B runs)
  if (mode && mode[0] == 'q' || !(tn > 0.0)) {
    square_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>((int)n - 1, d_e, d_e2);
    XMCA_LAUNCHED();
    stebz_kernel<<<(unsigned)((threads + 127) / 128), 128, 0, st>>>((int)n, d_d, d_e2, gl, gu, pivmin, 1e-18 * tn, d_w);
    XMCA_LAUNCHED();
    return XMCA_OK;
  }
  const double s = 1.0 / tn;
  double2* d_pk = reinterpret_cast<double2*>(d_scratch + 8);          // 16-byte aligned: d_scratch comes from an allocator
  XMCA_REQUIRE(((uintptr_t)d_pk & 15) == 0, "xmca_stebz: d_scratch must be 16-byte aligned");
  stebz_pack_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>((int)n, d_d, d_e, s, d_pk);
  XMCA_LAUNCHED();
  stebz_prod_kernel<<<(unsigned)((threads + 127) / 128), 128, 0, st>>>((int)n, d_pk, h[3] * s, tn, gl * s, gu * s, 1e-18, d_w);
  XMCA_LAUNCHED();
  return XMCA_OK;
}

extern "C" size_t xmca_stein_workspace_bytes(int64_t n, int64_t n_clusters) {
  return (size_t)n_clusters * 5 * (size_t)n * 8;
}

extern "C" int xmca_stein(int64_t n, const double* d_d, const double* d_e, int64_t k, const double* d_lambda,
                          const int* d_cluster_start, int64_t n_clusters, double tnorm, int iterations,
                          double* d_Z, int64_t ldz, void* d_workspace, size_t workspace_bytes, void* stream) {
  XMCA_REQUIRE(n >= 1 && k >= 1 && d_d && d_e && d_lambda && d_cluster_start && d_Z && d_workspace,
               "xmca_stein: bad argument");
  XMCA_REQUIRE(n_clusters >= 1 && n_clusters <= k && ldz >= n, "xmca_stein: bad cluster table / ldz");
  XMCA_REQUIRE(workspace_bytes >= xmca_stein_workspace_bytes(n, n_clusters), "xmca_stein: workspace too small");
  if (iterations <= 0) iterations = 2;   // singletons; clusters run one more
  stein_kernel<<<(unsigned)n_clusters, ST_THREADS, 0, (cudaStream_t)stream>>>(
      (int)n, d_d, d_e, d_lambda, d_cluster_start, (int)n_clusters, tnorm, iterations, d_Z, ldz,
      reinterpret_cast<double*>(d_workspace));
  XMCA_LAUNCHED();
  return XMCA_OK;
}

extern "C" int xmca_ormtr(int64_t n, const double* d_A, int64_t lda, const double* d_tau, int64_t k,
                          double* d_Z, int64_t ldz, void* stream) {
  XMCA_REQUIRE(n >= 1 && k >= 1 && d_A && d_tau && d_Z && lda >= n && ldz >= n, "xmca_ormtr: bad argument");
  XMCA_REQUIRE(n <= xmca_sytrd_max_n(), "xmca_ormtr: n too large");
  XMCA_CUDA(cudaFuncSetAttribute(ormtr_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 26000 * 8));
  ormtr_kernel<<<(unsigned)k, OR_THREADS, (size_t)n * 8, (cudaStream_t)stream>>>((int)n, d_A, lda, d_tau, d_Z, ldz);
  XMCA_LAUNCHED();
  return XMCA_OK;
}
