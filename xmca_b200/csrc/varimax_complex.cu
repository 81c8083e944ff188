// Fused Kaiser-normalised Varimax fixed point for COMPLEX loadings (complex MCA,
// xmca/tools/rotation.py:15-78 evaluated with complex dtype): same structure as
// varimax.cu -- one persistent cooperative kernel, per iteration one streaming pass
//     b = a R,   T1 += a^H (b |b|^2),   c += |b|^2
// a deterministic cross-CTA reduction, and on every CTA redundantly
//     T = T1 - (gamma/n) (G R) diag(c),  G = A^H A,   R = polar(T) = U V^H,  d = sum(s)
// with the polar factor from a warm-started one-sided COMPLEX Jacobi SVD in shared memory
// (phase-align the pair, then a real Givens rotation).  Loadings are planar (re, im) in
// storage precision; all p x p state is fp64 (the reference promotes to complex128).
// p <= 32.
#include "common.cuh"
#include <cooperative_groups.h>
#include <math.h>

namespace cg = cooperative_groups;

namespace xmca {

constexpr int CP = 32;                 // padded number of rotated modes
constexpr int CPP = CP + 1;            // stride of the column-major work matrices
constexpr int CT = 32;                 // rows per tile
constexpr int CTHREADS = 512;
constexpr int CSLOT = 2 * CP * CP + CP;   // doubles per partial: T1 re, T1 im, c
constexpr int CQ = 2 * CP + 4;            // pitch of the real images used by the MMA streaming pass (conflict-free fragments)

struct VarimaxCParams {
  const void* Lr; const void* Li; int64_t ldl; int64_t n; int p;
  double gamma; int max_iter; double tol;
  void* Anr; void* Ani;      // n x p normalised loadings (storage dtype, ld = p)
  double* h;                 // n
  double* partial;           // [grid][CSLOT]
  double* reduced;           // [CSLOT]
  double* Br; double* Bi; int64_t ldb; double* Rr; double* Ri; double* out;
};

// complex Z(i,j) = sum_k X(i,k) Y(k,j) [Y conjugated if conj_y], all p x p, element access through strides
// (as small_matmul in varimax.cu); thread -> column j = tid & 31 and rows i = (tid >> 5) + 16 q, q < 2
__device__ __forceinline__ void cmatmul(const double* Xr, const double* Xi, int xs_i, int xs_k,
                                        const double* Yr, const double* Yi, int ys_k, int ys_j, bool conj_y,
                                        double* Zr, double* Zi, int zs_i, int zs_j, int p) {
  const int j = threadIdx.x & 31, ig = threadIdx.x >> 5;
  double ar[2] = {0.0, 0.0}, ai[2] = {0.0, 0.0};
  if (j < p) {
    for (int k = 0; k < p; ++k) {
      const double yr = Yr[k * ys_k + j * ys_j];
      const double yi = conj_y ? -Yi[k * ys_k + j * ys_j] : Yi[k * ys_k + j * ys_j];
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        const double xr = Xr[(ig + 16 * q) * xs_i + k * xs_k], xi = Xi[(ig + 16 * q) * xs_i + k * xs_k];
        ar[q] = fma(xr, yr, fma(-xi, yi, ar[q]));
        ai[q] = fma(xr, yi, fma(xi, yr, ai[q]));
      }
    }
  }
#pragma unroll
  for (int q = 0; q < 2; ++q) {
    const int i = ig + 16 * q;
    const bool live = i < p && j < p;
    Zr[i * zs_i + j * zs_j] = live ? ar[q] : 0.0;
    Zi[i * zs_i + j * zs_j] = live ? ai[q] : 0.0;
  }
}

// One-sided complex Jacobi on the columns of X (p x p, column-major stride CPP, planar), accumulating V.
// One pair per warp (pe / 2 <= 16 warps), lane <-> row.
__device__ int polar_jacobi_complex(double* Xr, double* Xi, double* Vr, double* Vi, int pe,
                                    const unsigned char* rr, double* s_max, double stop2) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  const int npairs = pe >> 1;
  int sweeps = 0;
  while (sweeps < 40) {
    double cmax2 = 0.0;
    for (int step = 0; step < pe - 1; ++step) {
      if (warp < npairs) {
        const int cp = rr[step * pe + 2 * warp], cq = rr[step * pe + 2 * warp + 1];
        const double pr = Xr[cp * CPP + lane], pi = Xi[cp * CPP + lane];
        const double qr = Xr[cq * CPP + lane], qi = Xi[cq * CPP + lane];
        double al = pr * pr + pi * pi, be = qr * qr + qi * qi;
        double gr = pr * qr + pi * qi, gi = pr * qi - pi * qr;          // conj(x_p) x_q
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          al += __shfl_xor_sync(0xffffffffu, al, o);
          be += __shfl_xor_sync(0xffffffffu, be, o);
          gr += __shfl_xor_sync(0xffffffffu, gr, o);
          gi += __shfl_xor_sync(0xffffffffu, gi, o);
        }
        const double ab = al * be, g2 = gr * gr + gi * gi;
        if (g2 > 1e-30 * ab && g2 > 1e-300) {
          cmax2 = fmax(cmax2, g2 * fast_rcp(ab));
          const double ig = fast_rsqrt(g2);                 // 1 / |gamma|
          const double wr = gr * ig, wi = gi * ig;            // omega = gamma / |gamma|
          const float zf = (float)((be - al) * 0.5 * ig);
          const float tf = copysignf(1.0f, zf) / (fabsf(zf) + sqrtf(fmaf(zf, zf, 1.0f)));
          const double t = (double)tf;
          const double c = fast_rsqrt(fma(t, t, 1.0)), sn = c * t;
          // y = x_q conj(omega);  x_p' = c x_p - s y;  x_q' = s x_p + c y
          const double yr = qr * wr + qi * wi, yi = qi * wr - qr * wi;
          Xr[cp * CPP + lane] = c * pr - sn * yr; Xi[cp * CPP + lane] = c * pi - sn * yi;
          Xr[cq * CPP + lane] = sn * pr + c * yr; Xi[cq * CPP + lane] = sn * pi + c * yi;
          const double vpr = Vr[cp * CPP + lane], vpi = Vi[cp * CPP + lane];
          const double vqr = Vr[cq * CPP + lane], vqi = Vi[cq * CPP + lane];
          const double zr = vqr * wr + vqi * wi, zi = vqi * wr - vqr * wi;
          Vr[cp * CPP + lane] = c * vpr - sn * zr; Vi[cp * CPP + lane] = c * vpi - sn * zi;
          Vr[cq * CPP + lane] = sn * vpr + c * zr; Vi[cq * CPP + lane] = sn * vpi + c * zi;
        }
      }
      __syncthreads();
    }
    ++sweeps;
    if (lane == 0) s_max[warp] = cmax2;
    __syncthreads();
    double m = 0.0;
    for (int w = 0; w < nwarps; ++w) m = fmax(m, s_max[w]);
    __syncthreads();
    if (m <= stop2) break;              // largest squared cosine seen BEFORE this sweep's rotations
  }
  return sweeps;
}

template <typename TS>
__global__ void __launch_bounds__(CTHREADS, 1) varimax_complex_kernel(VarimaxCParams P) {
  cg::grid_group grid = cg::this_grid();
  extern __shared__ double sm[];
  double* Rr = sm;                      double* Ri = Rr + CP * CP;      // rotation, row-major
  double* Gr = Ri + CP * CP;            double* Gi = Gr + CP * CP;      // A^H A
  double* Wr = Gi + CP * CP;            double* Wi = Wr + CP * CP;      // scratch
  double* Vr = Wi + CP * CP;            double* Vi = Vr + CP * CPP;     // right singular vectors (column-major)
  double* Xr = Vi + CP * CPP;           double* Xi = Xr + CP * CPP;     // SVD work matrix, then U
  double* Ar = Xi + CP * CPP;           double* Ai = Ar + CT * CP;      // tile of a
  double* Br = Ai + CT * CP;            double* Bi = Br + CT * CP;      // tile of b |b|^2
  double* Tr = Ar;                      double* Ti = Ar + CP * CPP;     // T^T (stride CPP), aliases the tiles
  double* cs = Ar + 4 * CT * CP;        // [32]
  unsigned char* rr = reinterpret_cast<unsigned char*>(cs + CP);
  // streaming pass on the fp64 tensor-core path (iteration phase): real 64 x 64 image of the rotation, tiles at pitch CQ
  double* Rq = reinterpret_cast<double*>(rr + CP * CP);     // [64][CQ]: rows k < 32: [Rr | Ri], rows 32 + k: [-Ri | Rr]
  double* At = Rq + 2 * CP * CQ;                            // [CT][CQ]: [a_re (32) | a_im (32)]
  double* Bt = At + CT * CQ;                                // [CT][CQ]: [g_re | g_im],  g = b |b|^2
  __shared__ double s_max[CTHREADS / 32];

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int p = P.p, pe = (p + 1) & ~1;
  const int64_t n = P.n;
  const TS* Lr = reinterpret_cast<const TS*>(P.Lr);
  const TS* Li = reinterpret_cast<const TS*>(P.Li);
  TS* Anr = reinterpret_cast<TS*>(P.Anr);
  TS* Ani = reinterpret_cast<TS*>(P.Ani);
  const int64_t ntiles = (n + CT - 1) / CT;

  if (tid == 0) {
    unsigned char idx[CP];
    for (int i = 0; i < pe; ++i) idx[i] = (unsigned char)i;
    for (int r = 0; r < pe - 1; ++r) {
      for (int i = 0; i < pe / 2; ++i) { rr[r * pe + 2 * i] = idx[i]; rr[r * pe + 2 * i + 1] = idx[pe - 1 - i]; }
      unsigned char last = idx[pe - 1];
      for (int i = pe - 1; i > 1; --i) idx[i] = idx[i - 1];
      idx[1] = last;
    }
  }
  for (int e = tid; e < CP * CP; e += CTHREADS) {
    const int i = e >> 5, j = e & 31;
    Rr[e] = (i == j && i < p) ? 1.0 : 0.0; Ri[e] = 0.0;
    Vr[i * CPP + j] = (i == j) ? 1.0 : 0.0; Vi[i * CPP + j] = 0.0;
  }

  // accumulator mapping: thread owns T1(i, j) and T1(i + 16, j), i = warp, j = lane
  double t1r[2], t1i[2], csq;
  const int jcol = lane, rg = warp;           // b-phase: column jcol, rows 2 rg, 2 rg + 1

  // ---------------- phase 0: h, An = L / h, G = An^H An ----------------
  t1r[0] = t1r[1] = t1i[0] = t1i[1] = 0.0;
  for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int64_t r0 = tile * CT;
    __syncthreads();
    for (int e = tid; e < CT * CP; e += CTHREADS) {
      const int r = e >> 5, c = e & 31;
      const int64_t row = r0 + r;
      const bool ok = row < n && c < p;
      Ar[e] = ok ? (double)Lr[row * P.ldl + c] : 0.0;
      Ai[e] = ok ? (double)Li[row * P.ldl + c] : 0.0;
    }
    __syncthreads();
    for (int r = warp; r < CT; r += CTHREADS / 32) {      // row norms in storage precision (rotation.py:46-48)
      const TS a = (TS)Ar[r * CP + lane], b = (TS)Ai[r * CP + lane];
      const double ss = warp_sum((double)(a * a) + (double)(b * b));
      const TS hh = (TS)sqrt((TS)ss);
      const TS inv = (TS)1 / hh;
      const int64_t row = r0 + r;
      if (row < n) {
        if (lane == 0) P.h[row] = (double)hh;
        const TS nr = inv * a, ni = inv * b;
        if (lane < p) { Anr[row * p + lane] = nr; Ani[row * p + lane] = ni; }
        Br[r * CP + lane] = (double)nr; Bi[r * CP + lane] = (double)ni;
      } else { Br[r * CP + lane] = 0.0; Bi[r * CP + lane] = 0.0; }
    }
    __syncthreads();
#pragma unroll 4
    for (int r = 0; r < CT; ++r) {
      const double gr = Br[r * CP + lane], gi = Bi[r * CP + lane];
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        const double ar = Br[r * CP + warp + 16 * q], ai = Bi[r * CP + warp + 16 * q];   // conj(a_i) a_j
        t1r[q] = fma(ar, gr, fma(ai, gi, t1r[q]));
        t1i[q] = fma(ar, gi, fma(-ai, gr, t1i[q]));
      }
    }
  }
  {
    double* slot = P.partial + (int64_t)blockIdx.x * CSLOT;
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      slot[(warp + 16 * q) * CP + lane] = t1r[q];
      slot[CP * CP + (warp + 16 * q) * CP + lane] = t1i[q];
    }
    if (tid < CP) slot[2 * CP * CP + tid] = 0.0;
  }
  __threadfence();
  grid.sync();
  for (int e = blockIdx.x * CTHREADS + tid; e < CSLOT; e += gridDim.x * CTHREADS) {
    double s = 0.0;
    for (int k = 0; k < (int)gridDim.x; ++k) s += P.partial[(int64_t)k * CSLOT + e];
    P.reduced[e] = s;
  }
  __threadfence();
  grid.sync();
  for (int e = tid; e < CP * CP; e += CTHREADS) { Gr[e] = P.reduced[e]; Gi[e] = P.reduced[CP * CP + e]; }
  __syncthreads();

  // ---------------- fixed-point iteration ----------------
  double d = 0.0;
  int it = 0, converged = 0, svd_sweeps = 0;
  for (it = 1; it <= P.max_iter; ++it) {
    const double d_old = d;
    // ---- streaming pass on mma.sync.m8n8k4.f64 through the REAL image of the complex products:
    //   [b_re | b_im] = [a_re | a_im] [[Rr, Ri], [-Ri, Rr]]          (32 x 64 tile, K = 2 x 32 with the zero k-steps skipped)
    //   T1_re += a_re^T g_re + a_im^T g_im,   T1_im += a_re^T g_im - a_im^T g_re      (contraction over the tile's rows)
    // warp w: rows 8 (w & 3) and the real + imaginary 8-column block (w >> 2) of b (so that |b|^2 is thread local),
    // then the 8 x 8 block (i-block w & 3, j-block w >> 2) of T1.  The next tile is fetched into registers meanwhile.
    {
      const int gid = lane >> 2, tig = lane & 3;
      const int rb = warp & 3, cb = warp >> 2;
      const int kq = (p + 3) >> 2;
      for (int e = tid; e < 2 * CP * 2 * CP; e += CTHREADS) {          // real image of R (R changes every iteration)
        const int k = e >> 6, c = e & 63;
        const int kr = k & 31, cr = c & 31;
        double v;
        if (k < CP) v = c < CP ? Rr[kr * CP + cr] : Ri[kr * CP + cr];
        else v = c < CP ? -Ri[kr * CP + cr] : Rr[kr * CP + cr];
        Rq[k * CQ + c] = v;
      }
      double tr[2] = {0.0, 0.0}, tp[2] = {0.0, 0.0}, tn[2] = {0.0, 0.0}, cq[2] = {0.0, 0.0};
      const int lr = tid >> 5, lc = tid & 31;                           // loader: rows lr, lr + 16; column lc; re and im
      double pre[4];
      auto fetch = [&](int64_t tile) {
#pragma unroll
        for (int q = 0; q < 2; ++q) {
          const int64_t row = tile * CT + lr + 16 * q;
          const bool ok = row < n && lc < p;
          pre[2 * q] = ok ? (double)Anr[row * p + lc] : 0.0;
          pre[2 * q + 1] = ok ? (double)Ani[row * p + lc] : 0.0;
        }
      };
      if ((int64_t)blockIdx.x < ntiles) fetch(blockIdx.x);
      for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        __syncthreads();
#pragma unroll
        for (int q = 0; q < 2; ++q) {
          At[(lr + 16 * q) * CQ + lc] = pre[2 * q];
          At[(lr + 16 * q) * CQ + CP + lc] = pre[2 * q + 1];
        }
        __syncthreads();
        if (tile + gridDim.x < ntiles) fetch(tile + gridDim.x);
        {
          double cr[2] = {0.0, 0.0}, ci[2] = {0.0, 0.0};
          const double* ap = At + (8 * rb + gid) * CQ + tig;
          const double* rp = Rq + tig * CQ + 8 * cb + gid;
#pragma unroll
          for (int half = 0; half < 2; ++half)
            for (int kk = 0; kk < kq; ++kk) {
              const int k0 = CP * half + 4 * kk;
              const double a = ap[k0], b_r = rp[k0 * CQ], b_i = rp[k0 * CQ + CP];
              asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                           : "+d"(cr[0]), "+d"(cr[1]) : "d"(a), "d"(b_r));
              asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                           : "+d"(ci[0]), "+d"(ci[1]) : "d"(a), "d"(b_i));
            }
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            const double m2 = cr[e] * cr[e] + ci[e] * ci[e];
            cq[e] += m2;
            double* g = Bt + (8 * rb + gid) * CQ + 8 * cb + 2 * tig + e;
            g[0] = cr[e] * m2;
            g[CP] = ci[e] * m2;
          }
        }
        __syncthreads();
        {
          const double* ap = At + tig * CQ + 8 * rb + gid;              // a[row 4 ks + tig][column i0 + gid] (re; + CP: im)
          const double* gp = Bt + tig * CQ + 8 * cb + gid;              // g[row 4 ks + tig][column j0 + gid]
#pragma unroll
          for (int ks = 0; ks < CT / 4; ++ks) {
            const double a_r = ap[4 * ks * CQ], a_i = ap[4 * ks * CQ + CP];
            const double g_r = gp[4 * ks * CQ], g_i = gp[4 * ks * CQ + CP];
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                         : "+d"(tr[0]), "+d"(tr[1]) : "d"(a_r), "d"(g_r));
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                         : "+d"(tr[0]), "+d"(tr[1]) : "d"(a_i), "d"(g_i));
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                         : "+d"(tp[0]), "+d"(tp[1]) : "d"(a_r), "d"(g_i));
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                         : "+d"(tn[0]), "+d"(tn[1]) : "d"(a_i), "d"(g_r));
          }
        }
      }
      // per-CTA partial: T1 block of this warp, column sums of |b|^2 (summed over the four row-block warps)
      __syncthreads();
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        double v = cq[e];
        v += __shfl_xor_sync(0xffffffffu, v, 4);
        v += __shfl_xor_sync(0xffffffffu, v, 8);
        v += __shfl_xor_sync(0xffffffffu, v, 16);
        if (gid == 0) At[warp * 8 + 2 * tig + e] = v;
      }
      __syncthreads();
      double* slot = P.partial + (int64_t)blockIdx.x * CSLOT;
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int i = 8 * rb + gid, j = 8 * cb + 2 * tig + e;
        slot[i * CP + j] = tr[e];
        slot[CP * CP + i * CP + j] = tp[e] - tn[e];
      }
      if (tid < CP) {
        const int g4 = (tid >> 3) * 4, wi = tid & 7;                    // column tid = 8 cb + wi: warps 4 cb .. 4 cb + 3
        slot[2 * CP * CP + tid] = (At[g4 * 8 + wi] + At[(g4 + 1) * 8 + wi]) + (At[(g4 + 2) * 8 + wi] + At[(g4 + 3) * 8 + wi]);
      }
    }
    __threadfence();
    grid.sync();
    for (int e = blockIdx.x * CTHREADS + tid; e < CSLOT; e += gridDim.x * CTHREADS) {
      double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
      int k = 0;
      for (; k + 3 < (int)gridDim.x; k += 4) {
        s0 += P.partial[(int64_t)k * CSLOT + e];       s1 += P.partial[(int64_t)(k + 1) * CSLOT + e];
        s2 += P.partial[(int64_t)(k + 2) * CSLOT + e]; s3 += P.partial[(int64_t)(k + 3) * CSLOT + e];
      }
      for (; k < (int)gridDim.x; ++k) s0 += P.partial[(int64_t)k * CSLOT + e];
      P.reduced[e] = (s0 + s1) + (s2 + s3);
    }
    __threadfence();
    grid.sync();

    // ---- phase 2 (redundant on every CTA) ----
    if (tid < CP) cs[tid] = P.reduced[2 * CP * CP + tid];
    cmatmul(Gr, Gi, CP, 1, Rr, Ri, CP, 1, false, Wr, Wi, CP, 1, p);            // W = G R
    __syncthreads();
    const double gn = P.gamma / (double)n;
    for (int e = tid; e < CP * CP; e += CTHREADS) {
      const int i = e >> 5, k = e & 31;                                          // T^T[k][i] = T[i][k]
      const bool live = i < p && k < p;
      Tr[k * CPP + i] = live ? P.reduced[e] - gn * Wr[e] * cs[k] : 0.0;
      Ti[k * CPP + i] = live ? P.reduced[CP * CP + e] - gn * Wi[e] * cs[k] : 0.0;
    }
    __syncthreads();
    // X = T V, computed as X^T(j,i) = sum_k V^T(j,k) T^T(k,i)  (both factors plain, no conjugation)
    cmatmul(Vr, Vi, CPP, 1, Tr, Ti, CPP, 1, false, Xr, Xi, CPP, 1, p);
    __syncthreads();
    // inexact polar factor inside the iteration, polished once converged (see varimax.cu)
    svd_sweeps += polar_jacobi_complex(Xr, Xi, Vr, Vi, pe, rr, s_max, 1e-6);
    __syncthreads();
    auto finish_polar = [&]() {
      for (int j = warp; j < CP; j += CTHREADS / 32) {                            // sigma_j, U = X / sigma
        const double xr = Xr[j * CPP + lane], xi = Xi[j * CPP + lane];
        const double nn = sqrt(warp_sum(xr * xr + xi * xi));
        const bool live = (j < p) && nn > 0.0;
        if (lane == 0) cs[j] = (j < p) ? nn : 0.0;
        Xr[j * CPP + lane] = live ? xr / nn : 0.0;
        Xi[j * CPP + lane] = live ? xi / nn : 0.0;
      }
      __syncthreads();
      // R = U V^H : R(i,l) = sum_j U(i,j) conj(V(l,j)); U(i,j) = X[j*CPP+i], V(l,j) = V[j*CPP+l]
      cmatmul(Xr, Xi, 1, CPP, Vr, Vi, CPP, 1, true, Rr, Ri, CP, 1, p);
      double dd = 0.0;
      for (int j = 0; j < p; ++j) dd += cs[j];
      __syncthreads();
      return dd;
    };
    d = finish_polar();
    if (fabs(d - d_old) / d < P.tol) {
      cmatmul(Vr, Vi, CPP, 1, Tr, Ti, CPP, 1, false, Xr, Xi, CPP, 1, p);          // X = T V again
      __syncthreads();
      svd_sweeps += polar_jacobi_complex(Xr, Xi, Vr, Vi, pe, rr, s_max, 1e-22);
      __syncthreads();
      d = finish_polar();
      converged = 1;
      break;
    }
  }
  if (it > P.max_iter) it = P.max_iter;

  // ---------------- final: B = (h * An) R ----------------
  for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int64_t r0 = tile * CT;
    __syncthreads();
    for (int e = tid; e < CT * CP; e += CTHREADS) {
      const int r = e >> 5, c = e & 31;
      const int64_t row = r0 + r;
      const bool ok = row < n && c < p;
      Ar[e] = ok ? (double)((TS)P.h[row] * Anr[row * p + c]) : 0.0;
      Ai[e] = ok ? (double)((TS)P.h[row] * Ani[row * p + c]) : 0.0;
    }
    __syncthreads();
    if (jcol < p) {
      double br[2] = {0.0, 0.0}, bi[2] = {0.0, 0.0};
      for (int k = 0; k < p; ++k) {
        const double rkr = Rr[k * CP + jcol], rki = Ri[k * CP + jcol];
#pragma unroll
        for (int q = 0; q < 2; ++q) {
          const double ar = Ar[(rg * 2 + q) * CP + k], ai = Ai[(rg * 2 + q) * CP + k];
          br[q] = fma(ar, rkr, fma(-ai, rki, br[q]));
          bi[q] = fma(ar, rki, fma(ai, rkr, bi[q]));
        }
      }
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        const int64_t row = r0 + rg * 2 + q;
        if (row < n) { P.Br[row * P.ldb + jcol] = br[q]; P.Bi[row * P.ldb + jcol] = bi[q]; }
      }
    }
  }
  if (blockIdx.x == 0) {
    for (int e = tid; e < p * p; e += CTHREADS) {
      P.Rr[e] = Rr[(e / p) * CP + (e % p)];
      P.Ri[e] = Ri[(e / p) * CP + (e % p)];
    }
    if (tid == 0) { P.out[0] = (double)it; P.out[1] = (double)converged; P.out[2] = d; P.out[3] = (double)svd_sweeps; }
  }
}

static size_t varimaxc_smem_bytes() {
  return (size_t)(6 * CP * CP + 4 * CP * CPP + 4 * CT * CP + CP) * sizeof(double) + (size_t)CP * CP +
         (size_t)(2 * CP + 2 * CT) * CQ * sizeof(double);
}

constexpr size_t VC_MAX_GRID = 4 * 148 + 64;

}  // namespace xmca

using namespace xmca;

extern "C" size_t xmca_varimax_complex_workspace_bytes(int64_t n, int p) {
  size_t b = 0;
  b += 2 * (((size_t)n * p * 8 + 255) / 256 * 256);   // An re, im
  b += ((size_t)n * 8 + 255) / 256 * 256;             // h
  b += VC_MAX_GRID * CSLOT * 8;                       // partials
  b += (size_t)CSLOT * 8 + 256;                       // reduced
  return b;
}

extern "C" int xmca_varimax_complex(const void* d_Lr, const void* d_Li, int l_dtype, int64_t n, int p, int64_t ldl,
                                    double gamma, int max_iter, double tol,
                                    double* d_Br, double* d_Bi, int64_t ldb, double* d_Rr, double* d_Ri,
                                    int* iterations_out, double* d_out,
                                    void* d_workspace, size_t workspace_bytes, void* stream) {
  XMCA_REQUIRE(d_Lr && d_Li && d_Br && d_Bi && d_Rr && d_Ri && d_out && d_workspace, "xmca_varimax_complex: null argument");
  XMCA_REQUIRE(dtype_ok(l_dtype), "xmca_varimax_complex: bad dtype");
  XMCA_REQUIRE(n > 0 && p >= 2 && p <= CP, "xmca_varimax_complex: need 2 <= p <= 32");
  XMCA_REQUIRE(ldl >= p && ldb >= p, "xmca_varimax_complex: leading dimension too small");
  XMCA_REQUIRE(workspace_bytes >= xmca_varimax_complex_workspace_bytes(n, p), "xmca_varimax_complex: workspace too small");
  cudaStream_t st = (cudaStream_t)stream;
  const void* fn = (l_dtype == XMCA_F64) ? (const void*)varimax_complex_kernel<double> : (const void*)varimax_complex_kernel<float>;
  int occ = 0;
  XMCA_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)varimaxc_smem_bytes()));
  XMCA_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, fn, CTHREADS, varimaxc_smem_bytes()));
  if (occ < 1) return fail(XMCA_CUDA_ERROR, "complex varimax kernel does not fit on an SM", __FILE__, __LINE__);
  int64_t g = (int64_t)occ * sm_count();
  const int64_t tiles = (n + CT - 1) / CT;
  if (g > tiles) g = tiles;
  XMCA_REQUIRE((size_t)g <= VC_MAX_GRID, "xmca_varimax_complex: grid larger than workspace plan");

  char* ws = reinterpret_cast<char*>(d_workspace);
  VarimaxCParams P;
  P.Lr = d_Lr; P.Li = d_Li; P.ldl = ldl; P.n = n; P.p = p;
  P.gamma = gamma; P.max_iter = max_iter; P.tol = tol;
  size_t o = 0;
  const size_t an = ((size_t)n * p * 8 + 255) / 256 * 256;
  P.Anr = ws + o; o += an;
  P.Ani = ws + o; o += an;
  P.h = reinterpret_cast<double*>(ws + o); o += ((size_t)n * 8 + 255) / 256 * 256;
  P.partial = reinterpret_cast<double*>(ws + o); o += VC_MAX_GRID * CSLOT * 8;
  P.reduced = reinterpret_cast<double*>(ws + o);
  P.Br = d_Br; P.Bi = d_Bi; P.ldb = ldb; P.Rr = d_Rr; P.Ri = d_Ri; P.out = d_out;

  void* args[] = {&P};
  XMCA_CUDA(cudaLaunchCooperativeKernel(fn, dim3((unsigned)g), dim3(CTHREADS), args, varimaxc_smem_bytes(), st));
  XMCA_LAUNCHED();
  double h_out[4];
  XMCA_CUDA(cudaMemcpyAsync(h_out, d_out, sizeof h_out, cudaMemcpyDeviceToHost, st));
  XMCA_CUDA(cudaStreamSynchronize(st));
  if (iterations_out) *iterations_out = (int)h_out[0];
  if (h_out[1] == 0.0)
    return fail(XMCA_NOT_CONVERGED, "Rotation process did not converge.", __FILE__, __LINE__);
  return XMCA_OK;
}
