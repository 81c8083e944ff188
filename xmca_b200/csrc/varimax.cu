// Fused Kaiser-normalised Varimax fixed point (real case), one persistent
// cooperative kernel for the whole iteration of xmca/tools/rotation.py:15-78.
//
// Per iteration (rotation.py:52-64):
//   phase 1  every CTA streams its row tiles of the normalised loadings A once:
//            b = a R,  T1 += a^T (b*b*b),  c += b*b        (fp64 accumulation)
//   sync, distributed reduction of the per-CTA partials, sync
//   phase 2  EVERY CTA redundantly (bit-identically) forms
//            T = T1 - (gamma/n) (G R) diag(c),   G = A^T A  (identity of SURVEY 8d)
//            and its polar factor R = U V^T by a warm-started one-sided Jacobi
//            SVD in shared memory (one warp per column pair, shuffle
//            reductions); d = sum(s); stop when |d - d_old| / d < tol.
// Two grid-wide barriers per iteration, no host round trip, all p x p state fp64.
// Algorithmic HBM bytes per iteration: n * p * sizeof(storage) (one read of A).
#include "common.cuh"
#include <cooperative_groups.h>
#include <math.h>
#include <stdlib.h>

namespace cg = cooperative_groups;

namespace xmca {

constexpr int VP = 64;          // padded number of rotated modes (p <= 64)
constexpr int VT = 32;          // rows per tile
constexpr int VTHREADS = 512;
constexpr int VSLOT = VP * VP + VP;   // doubles per partial: T1 (64x64) + c (64)
constexpr int VPP = VP + 4;           // stride of EVERY p x p work matrix: fp64 MMA fragment loads (address
                                      // gid * VPP + tig or tig * VPP + gid) are bank-conflict free for stride = 4 mod 8
constexpr int RS = VPP;               // row stride of the rotation R and of the streamed tiles

struct VarimaxParams {
  const void* L; int ldt; int64_t n; int p; int64_t ldl;
  double gamma; int max_iter; double tol;
  void* An;            // n x p normalised loadings, same dtype as L, ld = p
  double* h;           // n row norms
  double* partial;     // [2*grid][VSLOT]
  double* reduced;     // [VSLOT]
  double* B; int64_t ldb; double* R; double* out;   // out: [0]=iterations [1]=converged [2]=d [3]=svd sweeps total
  unsigned int* bar;   // arrival counter of the lightweight grid barrier (zeroed before the launch)
  int jacobi_oe;       // in-loop sweeps: 0 round-robin, 1 odd-even ordering with register-resident columns (a warp per
                       // slot), 2 the same with half a warp per slot (default)
};

// Z[i][j] = sum_k X(i,k) Y(k,j) for i, j, k < p (padded entries -> 0), 64 x 64 output on the fp64 tensor-core path
// (mma.sync.m8n8k4): 16 warps, warp w owns the 16 x 16 block (w & 3, w >> 2): per k step of 4 two A and two B
// fragments (one double per lane each) feed four MMAs.  Element access: X(i,k) = X[i * xs_i + k * xs_k],
// Y(k,j) = Y[k * ys_k + j * ys_j] (transposed operands are expressed through strides; every matrix has the leading
// stride VPP, so the fragment loads are conflict free either way); Z is written at Z[i * zs_i + j * zs_j].
// At least one operand must be zero for k in [p, 4 ceil(p / 4)) -- all work matrices are zero padded.
__device__ __forceinline__ void dmma(double (&c)[2], double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c[0]), "+d"(c[1]) : "d"(a), "d"(b));
}
__device__ __forceinline__ void small_matmul(const double* X, int xs_i, int xs_k, const double* Y, int ys_k, int ys_j,
                                             double* Z, int zs_i, int zs_j, int p) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, gid = lane >> 2, tig = lane & 3;
  const int i0 = 16 * (w & 3), j0 = 16 * (w >> 2);
  double acc[2][2][2];
#pragma unroll
  for (int a = 0; a < 2; ++a)
#pragma unroll
    for (int b = 0; b < 2; ++b) { acc[a][b][0] = 0.0; acc[a][b][1] = 0.0; }
  const double* xp = X + (i0 + gid) * xs_i + tig * xs_k;
  const double* yp = Y + tig * ys_k + (j0 + gid) * ys_j;
  const int ksteps = (p + 3) >> 2;
#pragma unroll 4
  for (int kk = 0; kk < ksteps; ++kk) {
    const double a0 = xp[4 * kk * xs_k], a1 = xp[4 * kk * xs_k + 8 * xs_i];
    const double b0 = yp[4 * kk * ys_k], b1 = yp[4 * kk * ys_k + 8 * ys_j];
    dmma(acc[0][0], a0, b0);
    dmma(acc[0][1], a0, b1);
    dmma(acc[1][0], a1, b0);
    dmma(acc[1][1], a1, b1);
  }
#pragma unroll
  for (int a = 0; a < 2; ++a)
#pragma unroll
    for (int b = 0; b < 2; ++b)
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int i = i0 + 8 * a + gid, j = j0 + 8 * b + 2 * tig + e;
        Z[i * zs_i + j * zs_j] = (i < p && j < p) ? acc[a][b][e] : 0.0;
      }
}

// max over the CTA (every thread gets it); red: one double per warp
__device__ __forceinline__ double block_max(double v, double* red) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  double m = red[0];
  for (int w = 1; w < (int)(blockDim.x >> 5); ++w) m = fmax(m, red[w]);
  __syncthreads();
  return m;
}
__device__ __forceinline__ double block_sum(double v, double* red) {
  v = warp_sum(v);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  double m = 0.0;
  for (int w = 0; w < (int)(blockDim.x >> 5); ++w) m += red[w];      // fixed order: identical on every CTA
  __syncthreads();
  return m;
}

// Polar factor of a matrix X with NEARLY orthogonal columns (all cosines <= ~1e-3), without rotations:
//   polar(X) = X G^{-1/2},  G = X^T X = D + E  (D = diag(s_j^2), E small)
// and G^{-1/2} to second order in E through the divided differences of f(x) = x^{-1/2} (Daleckii-Krein):
//   Z_ij = [i = j] / s_i  -  Et_ij / (s_i s_j)  +  sum_k Et_ik Et_kj (s_i + s_j + s_k) / (s_k s_i s_j (s_i + s_j)),
//   Et_ij = E_ij / (s_i + s_j)
// (f[a, b] = -1 / (ra rb (ra + rb)),  f[a, b, c] = (ra + rb + rc) / (ra rb rc (ra + rb)(rb + rc)(ra + rc)), r = sqrt).
// The error is third order in the cosines: <= ~1e-8 here, where one more Jacobi sweep would cost 5x as much.
// In: Et (row-major, stride VPP, zero diagonal / padding), s[64].  Out: Z (row-major, stride VPP); returns
// sum_ij Z_ij G_ij = trace(polar(X)^T X) = the sum of the singular values.
__device__ __noinline__ double gram_inv_sqrt2(const double* Et, const double* s, double* Z, int p, double* red) {
  // f = Et diag(1 / s) Et and h = Et Et on the fp64 tensor-core path (same warp tiling as small_matmul)
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, gid = lane >> 2, tig = lane & 3;
  const int i0 = 16 * (w & 3), j0 = 16 * (w >> 2);
  double f[2][2][2], h[2][2][2];
#pragma unroll
  for (int a = 0; a < 2; ++a)
#pragma unroll
    for (int b = 0; b < 2; ++b) { f[a][b][0] = f[a][b][1] = 0.0; h[a][b][0] = h[a][b][1] = 0.0; }
  const double* xp = Et + (i0 + gid) * VPP + tig;
  const double* yp = Et + tig * VPP + j0 + gid;
  const int ksteps = (p + 3) >> 2;
#pragma unroll 2
  for (int kk = 0; kk < ksteps; ++kk) {
    const double sk = s[4 * kk + tig];
    const double rk = sk > 0.0 ? 1.0 / sk : 0.0;
    const double a0 = xp[4 * kk], a1 = xp[4 * kk + 8 * VPP];
    const double b0 = yp[4 * kk * VPP], b1 = yp[4 * kk * VPP + 8];
    const double a0r = a0 * rk, a1r = a1 * rk;
    dmma(h[0][0], a0, b0); dmma(h[0][1], a0, b1); dmma(h[1][0], a1, b0); dmma(h[1][1], a1, b1);
    dmma(f[0][0], a0r, b0); dmma(f[0][1], a0r, b1); dmma(f[1][0], a1r, b0); dmma(f[1][1], a1r, b1);
  }
  double dd = 0.0;
#pragma unroll
  for (int a = 0; a < 2; ++a)
#pragma unroll
    for (int b = 0; b < 2; ++b)
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int i = i0 + 8 * a + gid, j = j0 + 8 * b + 2 * tig + e;
        const double si = s[i], sj = s[j];
        double z = 0.0;
        if (i < p && j < p && si > 0.0 && sj > 0.0) {
          const double et = Et[i * VPP + j], ssum = si + sj, rij = 1.0 / (si * sj);
          z = ((i == j) ? 1.0 / si : 0.0) - et * rij + (ssum * f[a][b][e] + h[a][b][e]) * rij / ssum;
          dd = fma(z, (i == j) ? si * si : et * ssum, dd);
        }
        Z[i * VPP + j] = z;
      }
  return block_sum(dd, red);
}

// One-sided Jacobi on the columns of X (p x p), accumulating V.  Both are held COLUMN-major
// with stride VPP (X[c * VPP + r]) so that a warp reads a column without bank conflicts.
// pe = p rounded up to even (column p is a zero column when p is odd).
// rr: round-robin table [(pe-1)][pe].  Every warp rotates TWO column pairs per step with
// interleaved instruction streams (the step is latency bound: three 64-bit shuffle
// reductions, a division and two square roots), so one step serves all pe/2 <= 32 pairs of
// the round.  A sweep whose largest cosine (before rotating) is <= 1e-6 leaves cosines of
// ~1e-12: no confirming sweep is run.  Returns the number of sweeps used.
// skip2: pairs whose squared cosine is below it are left alone (the in-loop caller finishes with an expansion that
// is accurate for cosines up to ~1e-3, so only the few larger ones need a rotation).
__device__ int polar_jacobi(double* X, double* V, int pe, const unsigned char* rr, double* s_max, double stop2,
                            int max_sweeps = 40, double skip2 = 1e-30) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  const int npairs = pe >> 1;
  const bool fast = stop2 > 1e-13;           // the full-accuracy polish keeps fp64 reductions
  int sweeps = 0;
  while (sweeps < max_sweeps) {
    double cmax2 = 0.0;
    for (int step = 0; step < pe - 1; ++step) {
      int cp[2], cq[2];
      bool on[2];
      double xp0[2], xp1[2], xq0[2], xq1[2], al[2], be[2], ga[2];
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        const int pr = warp + u * nwarps;
        on[u] = pr < npairs;
        cp[u] = on[u] ? rr[step * pe + 2 * pr] : 0;
        cq[u] = on[u] ? rr[step * pe + 2 * pr + 1] : 1;
        // (an idle pair slot must not read columns that another warp rotates in this step)
        xp0[u] = on[u] ? X[cp[u] * VPP + lane] : 0.0; xp1[u] = on[u] ? X[cp[u] * VPP + lane + 32] : 0.0;
        xq0[u] = on[u] ? X[cq[u] * VPP + lane] : 0.0; xq1[u] = on[u] ? X[cq[u] * VPP + lane + 32] : 0.0;
        al[u] = xp0[u] * xp0[u] + xp1[u] * xp1[u];
        be[u] = xq0[u] * xq0[u] + xq1[u] * xq1[u];
        ga[u] = xp0[u] * xq0[u] + xp1[u] * xq1[u];
      }
      if (fast) {
        // in-loop sweeps (stop at cosine 1e-3): the three sums only steer the rotation angle, so they
        // are reduced in fp32 -- half the shuffle traffic on the crossbar the column loads/stores share.
        // Scaled by 1/al so that products of norms cannot overflow the fp32 range.
        // The six sums of a warp (two pairs x {al, be, ga}) go through ONE transposed butterfly: 4 + 2 + 1
        // exchange shuffles halve the value count while summing lanes, two plain steps finish, and six
        // broadcasts hand every lane all sums: 15 shuffles instead of 30.
        float x8[8] = {(float)al[0], (float)be[0], (float)ga[0], 0.0f, (float)al[1], (float)be[1], (float)ga[1], 0.0f};
        float y4[4], z2[2];
        const bool b4 = lane & 16, b3 = lane & 8, b2 = lane & 4;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const float keep = b4 ? x8[k + 4] : x8[k], send = b4 ? x8[k] : x8[k + 4];
          y4[k] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
        }
#pragma unroll
        for (int k = 0; k < 2; ++k) {
          const float keep = b3 ? y4[k + 2] : y4[k], send = b3 ? y4[k] : y4[k + 2];
          z2[k] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
        }
        float r1 = (b2 ? z2[1] : z2[0]) + __shfl_xor_sync(0xffffffffu, b2 ? z2[0] : z2[1], 4);
        r1 += __shfl_xor_sync(0xffffffffu, r1, 2);
        r1 += __shfl_xor_sync(0xffffffffu, r1, 1);
        // lane l now holds the total of value index v = (l >> 2) & 7 with bit order b2 + 2 b3 + 4 b4
        // (v = k_low + 2 k_mid + 4 k_high as selected above): value v sits in lanes 4 v .. 4 v + 3
        al[0] = (double)__shfl_sync(0xffffffffu, r1, 0);  be[0] = (double)__shfl_sync(0xffffffffu, r1, 4);
        ga[0] = (double)__shfl_sync(0xffffffffu, r1, 8);
        al[1] = (double)__shfl_sync(0xffffffffu, r1, 16); be[1] = (double)__shfl_sync(0xffffffffu, r1, 20);
        ga[1] = (double)__shfl_sync(0xffffffffu, r1, 24);
      } else {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
          for (int u = 0; u < 2; ++u) {
            al[u] += __shfl_xor_sync(0xffffffffu, al[u], o);
            be[u] += __shfl_xor_sync(0xffffffffu, be[u], o);
            ga[u] += __shfl_xor_sync(0xffffffffu, ga[u], o);
          }
        }
      }
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        if (!on[u]) continue;
        // |cos| = |ga| / sqrt(al be), compared and recorded through its square (no sqrt / division)
        const double ab = al[u] * be[u], g2 = ga[u] * ga[u];
        if (g2 > skip2 * ab && fabs(ga[u]) > 1e-300) {
          cmax2 = fmax(cmax2, g2 * fast_rcp(ab));
          // The rotation angle only steers convergence: zeta and t in fp32 (hardware MUFU ops);
          // c = 1/sqrt(1 + t^2), s = c t in fp64, so the rotation is orthogonal to rounding.
          const float zf = (float)((be[u] - al[u]) * fast_rcp(2.0 * ga[u]));
          const float tf = copysignf(1.0f, zf) / (fabsf(zf) + sqrtf(fmaf(zf, zf, 1.0f)));
          const double t = (double)tf;
          const double c = fast_rsqrt(fma(t, t, 1.0)), sn = c * t;
          X[cp[u] * VPP + lane] = c * xp0[u] - sn * xq0[u];        X[cq[u] * VPP + lane] = sn * xp0[u] + c * xq0[u];
          X[cp[u] * VPP + lane + 32] = c * xp1[u] - sn * xq1[u]; X[cq[u] * VPP + lane + 32] = sn * xp1[u] + c * xq1[u];
          const double vp0 = V[cp[u] * VPP + lane], vp1 = V[cp[u] * VPP + lane + 32];
          const double vq0 = V[cq[u] * VPP + lane], vq1 = V[cq[u] * VPP + lane + 32];
          V[cp[u] * VPP + lane] = c * vp0 - sn * vq0;        V[cq[u] * VPP + lane] = sn * vp0 + c * vq0;
          V[cp[u] * VPP + lane + 32] = c * vp1 - sn * vq1; V[cq[u] * VPP + lane + 32] = sn * vp1 + c * vq1;
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

// One sweep of the same one-sided Jacobi in the ODD-EVEN TRANSPOSITION ordering with the odd-position columns held
// in REGISTERS: slot s (warp w serves slots w and w + 16) keeps position 2 s + 1 ("co") for the whole sweep and works
// on it together with the even-position column to its left (even steps, position 2 s) or to its right (odd steps,
// position 2 s + 2), which lives in shared memory ("ce": one column read and one written per slot and step).  After
// every rotation the two columns trade places, so each column meets every other exactly once in pe steps and the
// order ends up reversed (undone at the end).  Per step a slot moves 8 column halves through shared memory; the
// round-robin version above moves 32 (both columns of X and V read and written) and was bound by exactly that
// traffic (~1000 clocks per step for 16 warps; measured 98 k clocks per sweep at p = 50).
// Same rotation, same skip / convergence bookkeeping, fp32 reductions for the rotation angle (in-loop use only).
__device__ int polar_jacobi_oe(double* X, double* V, int pe, double* s_max, double skip2, int* nrot = nullptr) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  const int nslots = pe >> 1;
  double cmax2 = 0.0;
  double xo0[2], xo1[2], vo0[2], vo1[2];                 // co of the two slots: rows lane, lane + 32 of X and V
  bool on[2];
  int slot[2];
#pragma unroll
  for (int u = 0; u < 2; ++u) {
    slot[u] = warp + u * nwarps;
    on[u] = slot[u] < nslots;
    const int co = on[u] ? 2 * slot[u] + 1 : 0;
    xo0[u] = on[u] ? X[co * VPP + lane] : 0.0; xo1[u] = on[u] ? X[co * VPP + lane + 32] : 0.0;
    vo0[u] = on[u] ? V[co * VPP + lane] : 0.0; vo1[u] = on[u] ? V[co * VPP + lane + 32] : 0.0;
  }
  for (int step = 0; step < pe; ++step) {
    const int odd = step & 1;
    int ce[2];
    bool act[2];
    double xe0[2], xe1[2], al[2], be[2], ga[2];
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      ce[u] = 2 * slot[u] + 2 * odd;                     // even step: position 2 s, odd step: position 2 s + 2
      act[u] = on[u] && ce[u] < pe;
      xe0[u] = act[u] ? X[ce[u] * VPP + lane] : 0.0; xe1[u] = act[u] ? X[ce[u] * VPP + lane + 32] : 0.0;
      al[u] = xo0[u] * xo0[u] + xo1[u] * xo1[u];
      be[u] = xe0[u] * xe0[u] + xe1[u] * xe1[u];
      ga[u] = xo0[u] * xe0[u] + xo1[u] * xe1[u];
    }
    {
      float x8[8] = {(float)al[0], (float)be[0], (float)ga[0], 0.0f, (float)al[1], (float)be[1], (float)ga[1], 0.0f};
      float y4[4], z2[2];
      const bool b4 = lane & 16, b3 = lane & 8, b2 = lane & 4;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float keep = b4 ? x8[k + 4] : x8[k], send = b4 ? x8[k] : x8[k + 4];
        y4[k] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
      }
#pragma unroll
      for (int k = 0; k < 2; ++k) {
        const float keep = b3 ? y4[k + 2] : y4[k], send = b3 ? y4[k] : y4[k + 2];
        z2[k] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
      }
      float r1 = (b2 ? z2[1] : z2[0]) + __shfl_xor_sync(0xffffffffu, b2 ? z2[0] : z2[1], 4);
      r1 += __shfl_xor_sync(0xffffffffu, r1, 2);
      r1 += __shfl_xor_sync(0xffffffffu, r1, 1);
      al[0] = (double)__shfl_sync(0xffffffffu, r1, 0);  be[0] = (double)__shfl_sync(0xffffffffu, r1, 4);
      ga[0] = (double)__shfl_sync(0xffffffffu, r1, 8);
      al[1] = (double)__shfl_sync(0xffffffffu, r1, 16); be[1] = (double)__shfl_sync(0xffffffffu, r1, 20);
      ga[1] = (double)__shfl_sync(0xffffffffu, r1, 24);
    }
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      if (!act[u]) continue;
      double c = 1.0, sn = 0.0;
      const double ab = al[u] * be[u], g2 = ga[u] * ga[u];
      if (g2 > skip2 * ab && fabs(ga[u]) > 1e-300) {
        cmax2 = fmax(cmax2, g2 * fast_rcp(ab));
        if (nrot && lane == 0) atomicAdd(nrot, 1);
        const float zf = (float)((be[u] - al[u]) * fast_rcp(2.0 * ga[u]));
        const float tf = copysignf(1.0f, zf) / (fabsf(zf) + sqrtf(fmaf(zf, zf, 1.0f)));
        const double t = (double)tf;
        c = fast_rsqrt(fma(t, t, 1.0));
        sn = c * t;
      }
      // rotate (p, q) = (co, ce) and trade places: the "q" result stays in the registers, the "p" result goes to ce
      const double ve0 = V[ce[u] * VPP + lane], ve1 = V[ce[u] * VPP + lane + 32];
      const double np0 = c * xo0[u] - sn * xe0[u], np1 = c * xo1[u] - sn * xe1[u];
      const double nq0 = sn * xo0[u] + c * xe0[u], nq1 = sn * xo1[u] + c * xe1[u];
      const double wp0 = c * vo0[u] - sn * ve0, wp1 = c * vo1[u] - sn * ve1;
      const double wq0 = sn * vo0[u] + c * ve0, wq1 = sn * vo1[u] + c * ve1;
      xo0[u] = nq0; xo1[u] = nq1; vo0[u] = wq0; vo1[u] = wq1;
      X[ce[u] * VPP + lane] = np0; X[ce[u] * VPP + lane + 32] = np1;
      V[ce[u] * VPP + lane] = wp0; V[ce[u] * VPP + lane + 32] = wp1;
    }
    __syncthreads();
  }
#pragma unroll
  for (int u = 0; u < 2; ++u) {
    if (!on[u]) continue;
    const int co = 2 * slot[u] + 1;
    X[co * VPP + lane] = xo0[u]; X[co * VPP + lane + 32] = xo1[u];
    V[co * VPP + lane] = vo0[u]; V[co * VPP + lane + 32] = vo1[u];
  }
  __syncthreads();
  // the sweep reversed the order of the columns: put them back (column q <-> column pe - 1 - q, X and V alike)
  for (int e = threadIdx.x; e < (pe >> 1) * VP; e += blockDim.x) {
    const int q = e >> 6, r = e & 63, q2 = pe - 1 - q;
    const double a = X[q * VPP + r], b = X[q2 * VPP + r];
    X[q * VPP + r] = b; X[q2 * VPP + r] = a;
    const double va = V[q * VPP + r], vb = V[q2 * VPP + r];
    V[q * VPP + r] = vb; V[q2 * VPP + r] = va;
  }
  if (lane == 0) s_max[warp] = cmax2;
  __syncthreads();
  return 1;
}

// The odd-even sweep with a FRACTION of a warp per slot (LPS = 16 or 8 lanes, 64 / LPS rows per lane): the step is
// bound by the number of warp instructions issued (two slots per warp used to run as two interleaved instruction
// streams), so serving 32 / LPS slots with ONE stream divides it.  Slot s = (32 / LPS) warp + lane / LPS; same
// rotations, same bookkeeping as polar_jacobi_oe.  Measured on the config-2 rotation (clocks per iteration of the
// whole polar phase): a warp per slot 144 k, half a warp 118 k (ships), a quarter 132 k (too few warps left).
template <int LPS>
__device__ int polar_jacobi_oe_part(double* X, double* V, int pe, double* s_max, double skip2, int* nrot = nullptr) {
  constexpr int RPL = VP / LPS, H1 = LPS / 2, H2 = LPS / 4;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, hl = lane & (LPS - 1), base = lane & ~(LPS - 1);
  const int nslots = pe >> 1, slot = (32 / LPS) * warp + lane / LPS;
  const bool on = slot < nslots;
  double cmax2 = 0.0;
  double xo[RPL], vo[RPL];
  {
    const int co = on ? 2 * slot + 1 : 0;
#pragma unroll
    for (int r = 0; r < RPL; ++r) {
      xo[r] = on ? X[co * VPP + hl + LPS * r] : 0.0;
      vo[r] = on ? V[co * VPP + hl + LPS * r] : 0.0;
    }
  }
  const bool bh = lane & H1, bq = lane & H2;
  for (int step = 0; step < pe; ++step) {
    const int ce = 2 * slot + 2 * (step & 1);            // even step: position 2 s, odd step: position 2 s + 2
    const bool act = on && ce < pe;
    double xe[RPL], al = 0.0, be = 0.0, ga = 0.0;
#pragma unroll
    for (int r = 0; r < RPL; ++r) {
      xe[r] = act ? X[ce * VPP + hl + LPS * r] : 0.0;
      al = fma(xo[r], xo[r], al);
      be = fma(xe[r], xe[r], be);
      ga = fma(xo[r], xe[r], ga);
    }
    {
      // three sums over the LPS lanes of the slot in fp32 (they only steer the angle): transposed butterfly
      const float x0 = (float)al, x1 = (float)be, x2 = (float)ga;
      const float y0 = (bh ? x2 : x0) + __shfl_xor_sync(0xffffffffu, bh ? x0 : x2, H1);
      const float y1 = (bh ? 0.0f : x1) + __shfl_xor_sync(0xffffffffu, bh ? x1 : 0.0f, H1);
      float z = (bq ? y1 : y0) + __shfl_xor_sync(0xffffffffu, bq ? y0 : y1, H2);
#pragma unroll
      for (int o = H2 / 2; o > 0; o >>= 1) z += __shfl_xor_sync(0xffffffffu, z, o);
      // lanes [0, H2): al, [H2, 2 H2): be, [2 H2, 3 H2): ga (of this slot's lane group)
      al = (double)__shfl_sync(0xffffffffu, z, base);
      be = (double)__shfl_sync(0xffffffffu, z, base + H2);
      ga = (double)__shfl_sync(0xffffffffu, z, base + 2 * H2);
    }
    double c = 1.0, sn = 0.0;
    {
      const double ab = al * be, g2 = ga * ga;
      if (act && g2 > skip2 * ab && fabs(ga) > 1e-300) {
        cmax2 = fmax(cmax2, g2 * fast_rcp(ab));
        if (nrot && hl == 0) atomicAdd(nrot, 1);
        const float zf = (float)((be - al) * fast_rcp(2.0 * ga));
        const float tf = copysignf(1.0f, zf) / (fabsf(zf) + sqrtf(fmaf(zf, zf, 1.0f)));
        const double t = (double)tf;
        c = fast_rsqrt(fma(t, t, 1.0));
        sn = c * t;
      }
    }
    if (act) {
      // rotate (p, q) = (co, ce) and trade places: the "q" result stays in the registers, the "p" result goes to ce
      double ve[RPL];
#pragma unroll
      for (int r = 0; r < RPL; ++r) ve[r] = V[ce * VPP + hl + LPS * r];
#pragma unroll
      for (int r = 0; r < RPL; ++r) {
        const double np = c * xo[r] - sn * xe[r], nq = sn * xo[r] + c * xe[r];
        const double wp = c * vo[r] - sn * ve[r], wq = sn * vo[r] + c * ve[r];
        xo[r] = nq; vo[r] = wq;
        X[ce * VPP + hl + LPS * r] = np;
        V[ce * VPP + hl + LPS * r] = wp;
      }
    }
    __syncthreads();
  }
  if (on) {
    const int co = 2 * slot + 1;
#pragma unroll
    for (int r = 0; r < RPL; ++r) { X[co * VPP + hl + LPS * r] = xo[r]; V[co * VPP + hl + LPS * r] = vo[r]; }
  }
  __syncthreads();
  // the sweep reversed the order of the columns: put them back (column q <-> column pe - 1 - q, X and V alike)
  for (int e = threadIdx.x; e < (pe >> 1) * VP; e += blockDim.x) {
    const int q = e >> 6, r = e & 63, q2 = pe - 1 - q;
    const double a = X[q * VPP + r], b = X[q2 * VPP + r];
    X[q * VPP + r] = b; X[q2 * VPP + r] = a;
    const double va = V[q * VPP + r], vb = V[q2 * VPP + r];
    V[q * VPP + r] = vb; V[q2 * VPP + r] = va;
  }
  if (hl == 0) s_max[(32 / LPS) * warp + lane / LPS] = cmax2;
  __syncthreads();
  return 1;
}

// Two-stage deterministic reduction of the per-CTA partials: every CTA owns a slice of the
// VSLOT entries; 16 lanes share one entry (strided over the CTAs) and combine by shuffles.
__device__ __forceinline__ void reduce_partials(const double* partial, double* reduced, int nslots) {
  const int per_cta = (VSLOT + (int)gridDim.x - 1) / (int)gridDim.x;
  const int sub = threadIdx.x & 15, loc = threadIdx.x >> 4;          // VTHREADS / 16 = 32 entries per pass
  for (int base = 0; base < per_cta; base += VTHREADS / 16) {
    const int e = blockIdx.x * per_cta + base + loc;
    const bool ok = (base + loc) < per_cta && e < VSLOT;
    double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
    if (ok) {
      int k = sub;
      for (; k + 48 < nslots; k += 64) {
        s0 += partial[(int64_t)k * VSLOT + e];
        s1 += partial[(int64_t)(k + 16) * VSLOT + e];
        s2 += partial[(int64_t)(k + 32) * VSLOT + e];
        s3 += partial[(int64_t)(k + 48) * VSLOT + e];
      }
      for (; k < nslots; k += 16) s0 += partial[(int64_t)k * VSLOT + e];
    }
    double s = (s0 + s1) + (s2 + s3);
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (ok && sub == 0) reduced[e] = s;
  }
}

// Grid-wide barrier of the co-resident CTAs (cooperative launch): one monotonically increasing arrival counter,
// red.release / ld.acquire by thread 0 between two CTA barriers (the same barrier as in tridiag.cu; about half the
// latency of cooperative_groups' grid.sync(), which the iteration pays twice).
__device__ __forceinline__ void vm_grid_barrier(unsigned int* counter, unsigned int& epoch) {
  __syncthreads();
  if (threadIdx.x == 0) {
    epoch += gridDim.x;
    asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(counter) : "memory");
    unsigned int seen;
    do {
      asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(counter) : "memory");
    } while (seen < epoch);
  }
  __syncthreads();
}

template <typename TS>
__global__ void __launch_bounds__(VTHREADS, 1) varimax_kernel(VarimaxParams P) {
  cg::grid_group grid = cg::this_grid();
  extern __shared__ double sm[];
  double* Rs = sm;                    // rotation (row-major; all p x p matrices: 64 x VPP)
  double* Gs = Rs + VP * VPP;         // A^T A
  double* Ws = Gs + VP * VPP;         // scratch (G R)
  double* Vs = Ws + VP * VPP;         // right singular vectors, warm start   (column-major, stride VPP)
  double* Xs = Vs + VP * VPP;         // SVD work matrix, then U              (column-major, stride VPP)
  double* As = Xs + VP * VPP;         // tile [32][64]
  double* Bs = As + VT * VP;          // tile [32][64]
  double* Ts = As;                    // T^T (stride VPP), aliases the tiles between the streaming passes
  double* cs = As + VP * VPP;         // [64] column sums / sigma
  unsigned char* rr = reinterpret_cast<unsigned char*>(cs + VP);   // [(pe-1)*pe]
  __shared__ double s_max[VTHREADS / 32];
  __shared__ double s_max2[4 * (VTHREADS / 32)];   // (per slot of the part-warp sweeps)
  __shared__ int s_cnt[4];            // diagnostics (out[10..12]): rotations applied by the in-loop sweeps, pairs whose
                                      // cosine exceeded 1e-4 / 1e-3 when a Gram matrix was inspected

  const int tid = threadIdx.x, p = P.p, pe = (p + 1) & ~1;
  if (tid < 4) s_cnt[tid] = 0;
  const int64_t n = P.n;
  const TS* L = reinterpret_cast<const TS*>(P.L);
  TS* An = reinterpret_cast<TS*>(P.An);
  const int64_t ntiles = (n + VT - 1) / VT;

  // round-robin table for pe columns
  if (tid == 0) {
    unsigned char idx[VP];
    for (int i = 0; i < pe; ++i) idx[i] = (unsigned char)i;
    for (int r = 0; r < pe - 1; ++r) {
      for (int i = 0; i < pe / 2; ++i) { rr[r * pe + 2 * i] = idx[i]; rr[r * pe + 2 * i + 1] = idx[pe - 1 - i]; }
      unsigned char last = idx[pe - 1];
      for (int i = pe - 1; i > 1; --i) idx[i] = idx[i - 1];
      idx[1] = last;
    }
  }
  for (int e = tid; e < VP * VP; e += VTHREADS) {
    int i = e >> 6, j = e & 63;
    double id = (i == j && i < p) ? 1.0 : 0.0;
    Rs[i * RS + j] = id; Vs[i * VPP + j] = (i == j) ? 1.0 : 0.0;
  }

  // T1-phase thread mapping: 4x4 register block of the 64x64 accumulator, two row-halves
  const int half = tid >> 8, bi = (tid & 255) >> 4, bj = tid & 15;
  // B-phase mapping: column j, 4 rows
  const int jcol = tid & 63, rg = tid >> 6;

  double acc[4][4];

  // ---------------- phase 0: h, An = L / h, G = An^T An ----------------
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.0;
  for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int64_t r0 = tile * VT;
    __syncthreads();
    for (int e = tid; e < VT * VP; e += VTHREADS) {
      int r = e >> 6, c = e & 63;
      int64_t row = r0 + r;
      As[e] = (row < n && c < p) ? (double)L[row * P.ldl + c] : 0.0;
    }
    __syncthreads();
    {   // row norms in the storage precision of the reference (rotation.py:46-48)
      const int warp = tid >> 5, lane = tid & 31;
      for (int r = warp; r < VT; r += VTHREADS / 32) {
        TS a0 = (TS)As[r * VP + lane], a1 = (TS)As[r * VP + lane + 32];
        double ss = warp_sum((double)(a0 * a0) + (double)(a1 * a1));
        TS hh = (TS)sqrt((TS)ss);
        TS inv = (TS)1 / hh;
        int64_t row = r0 + r;
        if (row < n) {
          if (lane == 0) P.h[row] = (double)hh;
          TS n0 = inv * a0, n1 = inv * a1;
          if (lane < p) An[row * p + lane] = n0;
          if (lane + 32 < p) An[row * p + lane + 32] = n1;
          Bs[r * VP + lane] = (double)n0; Bs[r * VP + lane + 32] = (double)n1;
        } else {
          Bs[r * VP + lane] = 0.0; Bs[r * VP + lane + 32] = 0.0;
        }
      }
    }
    __syncthreads();
#pragma unroll 4
    for (int r = half * 16; r < half * 16 + 16; ++r) {
      double a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) { a[i] = Bs[r * VP + bi + 16 * i]; b[i] = Bs[r * VP + bj + 16 * i]; }
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fma(a[i], b[j], acc[i][j]);
    }
  }
  {
    double* slot = P.partial + ((int64_t)blockIdx.x * 2 + half) * VSLOT;
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) slot[(bi + 16 * i) * VP + bj + 16 * j] = acc[i][j];
    if ((tid & 255) < VP) slot[VP * VP + (tid & 255)] = 0.0;
  }
  __threadfence();
  grid.sync();
  reduce_partials(P.partial, P.reduced, 2 * (int)gridDim.x);
  __threadfence();
  grid.sync();
  for (int e = tid; e < VP * VP; e += VTHREADS) Gs[(e >> 6) * VPP + (e & 63)] = P.reduced[e];
  __syncthreads();

  // ---------------- fixed-point iteration ----------------
  double d = 0.0;
  int it = 0, converged = 0, svd_sweeps = 0;
  unsigned int bar_epoch = 0;
  long long tk[6] = {0, 0, 0, 0, 0, 0};   // per-phase clock64 totals (block 0), returned in out[4..9]
  for (it = 1; it <= P.max_iter; ++it) {
    const double d_old = d;
    long long c0 = clock64();
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[i][j] = 0.0;
    // Streaming pass on the fp64 tensor-core path (mma.sync.m8n8k4): per 32-row tile  b = a R  (warp: 8 rows x 16
    // columns), then T1 += a^T (b^3) (warp: a 16 x 16 block of T1, accumulators kept over all tiles of the CTA) and
    // the column sums of b^2.  Tiles live in shared memory with row stride RS (conflict-free fragment loads); the
    // next tile's elements are fetched into registers while the current one is processed.
    {
      double* At = As;                 // [VT][RS] tile of An            (aliases Ts: free during the pass)
      double* Bt = Ws;                 // [VT][RS] tile of (a R)^3       (Ws: scratch of the polar phase)
      const int lane = tid & 31, w = tid >> 5, gid = lane >> 2, tig = lane & 3;
      const int rb = w & 3, cg = w >> 2;             // b = a R: rows 8 rb .., columns 16 cg ..
      const int ksteps = (p + 3) >> 2;
      double t1[2][2][2], cq[2][2];
#pragma unroll
      for (int a = 0; a < 2; ++a)
#pragma unroll
        for (int b = 0; b < 2; ++b) { t1[a][b][0] = 0.0; t1[a][b][1] = 0.0; cq[a][b] = 0.0; }
      const int lr = tid >> 6, lc = tid & 63;        // loader: rows lr + 8 q, column lc
      double pre[4];
      auto fetch = [&](int64_t tile) {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int64_t row = tile * VT + lr + 8 * q;
          pre[q] = (row < n && lc < p) ? (double)An[row * p + lc] : 0.0;
        }
      };
      if ((int64_t)blockIdx.x < ntiles) fetch(blockIdx.x);
      for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        __syncthreads();
#pragma unroll
        for (int q = 0; q < 4; ++q) At[(lr + 8 * q) * RS + lc] = pre[q];
        __syncthreads();
        if (tile + gridDim.x < ntiles) fetch(tile + gridDim.x);
        {
          double c[2][2] = {{0.0, 0.0}, {0.0, 0.0}};
          const double* ap = At + (8 * rb + gid) * RS + tig;
          const double* rp = Rs + tig * RS + 16 * cg + gid;
#pragma unroll 4
          for (int kk = 0; kk < ksteps; ++kk) {
            const double a = ap[4 * kk], b0 = rp[4 * kk * RS], b1 = rp[4 * kk * RS + 8];
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                         : "+d"(c[0][0]), "+d"(c[0][1]) : "d"(a), "d"(b0));
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                         : "+d"(c[1][0]), "+d"(c[1][1]) : "d"(a), "d"(b1));
          }
#pragma unroll
          for (int nb = 0; nb < 2; ++nb) {
            double2 v;
            cq[nb][0] = fma(c[nb][0], c[nb][0], cq[nb][0]);
            cq[nb][1] = fma(c[nb][1], c[nb][1], cq[nb][1]);
            v.x = c[nb][0] * c[nb][0] * c[nb][0];
            v.y = c[nb][1] * c[nb][1] * c[nb][1];
            *reinterpret_cast<double2*>(Bt + (8 * rb + gid) * RS + 16 * cg + 8 * nb + 2 * tig) = v;
          }
        }
        __syncthreads();
        {
          const double* ap = At + tig * RS + 16 * rb + gid;      // A fragment: a[row 4 ks + tig][column i0 + gid]
          const double* bp = Bt + tig * RS + 16 * cg + gid;      // B fragment: b3[row 4 ks + tig][column j0 + gid]
#pragma unroll
          for (int ks = 0; ks < VT / 4; ++ks) {
            const double a0 = ap[4 * ks * RS], a1 = ap[4 * ks * RS + 8];
            const double b0 = bp[4 * ks * RS], b1 = bp[4 * ks * RS + 8];
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                         : "+d"(t1[0][0][0]), "+d"(t1[0][0][1]) : "d"(a0), "d"(b0));
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                         : "+d"(t1[0][1][0]), "+d"(t1[0][1][1]) : "d"(a0), "d"(b1));
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                         : "+d"(t1[1][0][0]), "+d"(t1[1][0][1]) : "d"(a1), "d"(b0));
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                         : "+d"(t1[1][1][0]), "+d"(t1[1][1][1]) : "d"(a1), "d"(b1));
          }
        }
      }
      // per-CTA partial (ONE slot per CTA in the iteration): T1 block of this warp, column sums of b^2
      __syncthreads();
#pragma unroll
      for (int nb = 0; nb < 2; ++nb)
#pragma unroll
        for (int u = 0; u < 2; ++u) {
          double v = cq[nb][u];
          v += __shfl_xor_sync(0xffffffffu, v, 4);
          v += __shfl_xor_sync(0xffffffffu, v, 8);
          v += __shfl_xor_sync(0xffffffffu, v, 16);
          if (gid == 0) At[w * 16 + 8 * nb + 2 * tig + u] = v;
        }
      __syncthreads();
      double* slot = P.partial + (int64_t)blockIdx.x * VSLOT;
#pragma unroll
      for (int ib = 0; ib < 2; ++ib)
#pragma unroll
        for (int jb = 0; jb < 2; ++jb) {
          double2 v;
          v.x = t1[ib][jb][0];
          v.y = t1[ib][jb][1];
          *reinterpret_cast<double2*>(slot + (16 * rb + 8 * ib + gid) * VP + 16 * cg + 8 * jb + 2 * tig) = v;
        }
      if (tid < VP) {
        const int g4 = (tid >> 4) * 4, wi = tid & 15;
        slot[VP * VP + tid] = (At[g4 * 16 + wi] + At[(g4 + 1) * 16 + wi]) + (At[(g4 + 2) * 16 + wi] + At[(g4 + 3) * 16 + wi]);
      }
    }
    { long long c1 = clock64(); tk[0] += c1 - c0; c0 = c1; }
    vm_grid_barrier(P.bar, bar_epoch);
    { long long c1 = clock64(); tk[1] += c1 - c0; c0 = c1; }
    reduce_partials(P.partial, P.reduced, (int)gridDim.x);
    vm_grid_barrier(P.bar, bar_epoch);
    { long long c1 = clock64(); tk[2] += c1 - c0; c0 = c1; }

    // ---- phase 2 (redundant on every CTA): T, polar factor, convergence ----
    if (tid < VP) cs[tid] = P.reduced[VP * VP + tid];
    small_matmul(Gs, VPP, 1, Rs, RS, 1, Ws, VPP, 1, p);    // Ws = G R
    __syncthreads();
    const double gn = P.gamma / (double)n;
    for (int e = tid; e < VP * VP; e += VTHREADS) {
      int i = e >> 6, k = e & 63;                        // T^T[k][i] = T[i][k]
      Ts[k * VPP + i] = (i < p && k < p) ? P.reduced[e] - gn * Ws[i * VPP + k] * cs[k] : 0.0;
    }
    __syncthreads();
    // X = T V (warm start): X(i,j) = sum_k T(i,k) V(k,j); T(i,k) = Ts[k*VPP+i], V(k,j) = Vs[j*VPP+k]; lanes <-> i
    small_matmul(Vs, VPP, 1, Ts, VPP, 1, Xs, VPP, 1, p);   // computed as X^T = V^T T^T: Z(j,i) = sum_k V^T(j,k) T^T(k,i)
    __syncthreads();
    { long long c1 = clock64(); tk[3] += c1 - c0; c0 = c1; }
    // Polar factor inside the iteration: X = T V (V warm-started) already has nearly orthogonal columns.  The
    // Gram matrix G = X^T X tells exactly how nearly: while a cosine exceeds 1e-3 one Jacobi sweep is applied
    // (X and V rotated); then polar(X) = X G^{-1/2} comes from the second-order expansion above (error ~1e-8)
    // and R = polar(X) V^T.  The converged rotation is polished to full accuracy by Jacobi sweeps below.
    double d_new = 0.0;
    for (int guard = 0; guard < 40; ++guard) {
      small_matmul(Xs, VPP, 1, Xs, 1, VPP, Ws, VPP, 1, p);           // G(i,j) = x_i . x_j
      __syncthreads();
      if (tid < VP) cs[tid] = (tid < p) ? sqrt(Ws[tid * VPP + tid]) : 0.0;
      __syncthreads();
      double m2 = 0.0;
      for (int e = tid; e < VP * VP; e += VTHREADS) {
        const int i = e >> 6, j = e & 63;
        const double den = cs[i] * cs[j];
        if (i != j && i < p && j < p && den > 0.0) {
          const double c = Ws[i * VPP + j] / den;
          m2 = fmax(m2, c * c);
          if (i < j && c * c > 1e-8) atomicAdd(&s_cnt[1], 1);
          if (i < j && c * c > 1e-6) atomicAdd(&s_cnt[2], 1);
        }
      }
      m2 = block_max(m2, s_max);
      if (m2 <= 1e-6) break;
      svd_sweeps += P.jacobi_oe == 2 ? polar_jacobi_oe_part<16>(Xs, Vs, pe, s_max2, 1e-8, &s_cnt[0]) :
                    P.jacobi_oe == 1 ? polar_jacobi_oe(Xs, Vs, pe, s_max, 1e-8, &s_cnt[0]) : polar_jacobi(Xs, Vs, pe, rr, s_max, 1e-6, 1, 1e-8);
      __syncthreads();
    }
    for (int e = tid; e < VP * VP; e += VTHREADS) {                  // G -> Et in place
      const int i = e >> 6, j = e & 63;
      const double ssum = cs[i] + cs[j];
      Ws[i * VPP + j] = (i != j && i < p && j < p && ssum > 0.0) ? Ws[i * VPP + j] / ssum : 0.0;
    }
    __syncthreads();
    d_new = gram_inv_sqrt2(Ws, cs, Rs, p, s_max);                    // Rs = Z
    __syncthreads();
    small_matmul(Rs, RS, 1, Vs, VPP, 1, Ws, VPP, 1, p);              // Y = Z V^T:  Y(i,l) = sum_j Z(i,j) V(l,j)
    __syncthreads();
    small_matmul(Xs, 1, VPP, Ws, VPP, 1, Rs, RS, 1, p);              // R = X Y
    __syncthreads();
    { long long c1 = clock64(); tk[4] += c1 - c0; c0 = c1; }
    auto finish_polar = [&]() {
      // sigma_j = ||x_j||, U = X / sigma (in place), R = U V^T, d = sum sigma
      const int warp = tid >> 5, lane = tid & 31;
      for (int j = warp; j < VP; j += VTHREADS / 32) {
        const double x0 = Xs[j * VPP + lane], x1 = Xs[j * VPP + lane + 32];
        const double nn = sqrt(warp_sum(x0 * x0 + x1 * x1));
        const bool live = (j < p) && nn > 0.0;
        if (lane == 0) cs[j] = (j < p) ? nn : 0.0;
        Xs[j * VPP + lane] = live ? x0 / nn : 0.0;
        Xs[j * VPP + lane + 32] = live ? x1 / nn : 0.0;
      }
      __syncthreads();
      // R(i,l) = sum_j U(i,j) V(l,j); U(i,j) = Xs[j*VPP+i], V(l,j) = Vs[j*VPP+l]
      small_matmul(Xs, 1, VPP, Vs, VPP, 1, Rs, RS, 1, p);
      double dd = 0.0;
      for (int j = 0; j < p; ++j) dd += cs[j];
      __syncthreads();
      return dd;
    };
    d = d_new;
    if (fabs(d - d_old) / d < P.tol) {
      // converged: redo the last polar factor to full accuracy (cosines <= 1e-11 before the final sweep)
      small_matmul(Vs, VPP, 1, Ts, VPP, 1, Xs, VPP, 1, p);          // X = T V with the accumulated V
      __syncthreads();
      svd_sweeps += polar_jacobi(Xs, Vs, pe, rr, s_max, 1e-22);
      __syncthreads();
      d = finish_polar();
      converged = 1;
      break;
    }
    { long long c1 = clock64(); tk[5] += c1 - c0; c0 = c1; }
  }
  if (it > P.max_iter) it = P.max_iter;

  // ---------------- final: B = (h * An) R  (rotation.py:74-77) ----------------
  for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int64_t r0 = tile * VT;
    __syncthreads();
    for (int e = tid; e < VT * VP; e += VTHREADS) {
      int r = e >> 6, c = e & 63;
      int64_t row = r0 + r;
      // de-normalise in storage precision like the reference (h * A), then promote
      As[e] = (row < n && c < p) ? (double)((TS)P.h[row] * An[row * p + c]) : 0.0;
    }
    __syncthreads();
    if (jcol < p) {
      double b4[4] = {0.0, 0.0, 0.0, 0.0};
      for (int k = 0; k < p; ++k) {
        const double rk = Rs[k * RS + jcol];
#pragma unroll
        for (int q = 0; q < 4; ++q) b4[q] = fma(As[(rg * 4 + q) * VP + k], rk, b4[q]);
      }
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        int64_t row = r0 + rg * 4 + q;
        if (row < n) P.B[row * P.ldb + jcol] = b4[q];
      }
    }
  }
  if (blockIdx.x == 0) {
    for (int e = tid; e < p * p; e += VTHREADS) P.R[e] = Rs[(e / p) * RS + (e % p)];
    if (tid == 0) { P.out[0] = (double)it; P.out[1] = (double)converged; P.out[2] = d; P.out[3] = (double)svd_sweeps;
      for (int q = 0; q < 6; ++q) P.out[4 + q] = (double)tk[q];
      for (int q = 0; q < 3; ++q) P.out[10 + q] = (double)s_cnt[q]; }
  }
}

static size_t varimax_smem_bytes() {
  return (size_t)(6 * VP * VPP + VP) * sizeof(double) + (size_t)VP * VP;
}

template <typename TS>
static int varimax_grid(int64_t n, int* grid_out) {
  int occ = 0;
  XMCA_CUDA(cudaFuncSetAttribute(varimax_kernel<TS>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (int)varimax_smem_bytes()));
  XMCA_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, varimax_kernel<TS>, VTHREADS,
                                                          varimax_smem_bytes()));
  if (occ < 1) return fail(XMCA_CUDA_ERROR, "varimax kernel does not fit on an SM", __FILE__, __LINE__);
  int64_t g = (int64_t)occ * sm_count();
  int64_t tiles = (n + VT - 1) / VT;
  if (g > tiles) g = tiles;
  *grid_out = (int)g;
  return XMCA_OK;
}

}  // namespace xmca

using namespace xmca;

extern "C" size_t xmca_varimax_workspace_bytes(int64_t n, int p) {
  size_t max_grid = (size_t)4 * 148 + 64;
  size_t b = 0;
  b += ((size_t)n * p * 8 + 255) / 256 * 256;     // An (storage dtype, <= 8 bytes)
  b += ((size_t)n * 8 + 255) / 256 * 256;         // h
  b += max_grid * 2 * VSLOT * 8;                  // partials
  b += (size_t)VSLOT * 8 + 256;                   // reduced
  return b;
}

extern "C" int xmca_varimax(const void* d_L, int l_dtype, int64_t n, int p, int64_t ldl,
                            double gamma, int max_iter, double tol,
                            double* d_B, int64_t ldb, double* d_R, int* iterations_out, double* d_out,
                            void* d_workspace, size_t workspace_bytes, void* stream) {
  XMCA_REQUIRE(d_L && d_B && d_R && d_out && d_workspace, "xmca_varimax: null argument");
  XMCA_REQUIRE(dtype_ok(l_dtype), "xmca_varimax: bad dtype");
  XMCA_REQUIRE(n > 0 && p >= 2 && p <= VP, "xmca_varimax: need 2 <= p <= 64");
  XMCA_REQUIRE(ldl >= p && ldb >= p, "xmca_varimax: leading dimension too small");
  XMCA_REQUIRE(workspace_bytes >= xmca_varimax_workspace_bytes(n, p), "xmca_varimax: workspace too small");
  cudaStream_t st = (cudaStream_t)stream;
  int grid = 0, rc;
  rc = (l_dtype == XMCA_F64) ? varimax_grid<double>(n, &grid) : varimax_grid<float>(n, &grid);
  if (rc != XMCA_OK) return rc;
  XMCA_REQUIRE((size_t)grid <= (size_t)4 * 148 + 64, "xmca_varimax: grid larger than workspace plan");

  char* ws = reinterpret_cast<char*>(d_workspace);
  VarimaxParams P;
  P.L = d_L; P.ldt = l_dtype; P.n = n; P.p = p; P.ldl = ldl;
  P.gamma = gamma; P.max_iter = max_iter; P.tol = tol;
  size_t o = 0;
  P.An = ws + o; o += ((size_t)n * p * 8 + 255) / 256 * 256;
  P.h = reinterpret_cast<double*>(ws + o); o += ((size_t)n * 8 + 255) / 256 * 256;
  P.partial = reinterpret_cast<double*>(ws + o); o += ((size_t)4 * 148 + 64) * 2 * VSLOT * 8;
  P.reduced = reinterpret_cast<double*>(ws + o);
  P.bar = reinterpret_cast<unsigned int*>(P.reduced + VSLOT);        // (inside the 256-byte pad of the plan)
  XMCA_CUDA(cudaMemsetAsync(P.bar, 0, sizeof(unsigned int), st));
  P.B = d_B; P.ldb = ldb; P.R = d_R; P.out = d_out;
  {
    const char* e = getenv("XMCA_VARIMAX_JACOBI");       // A/B runs: "rr" round-robin, "oe" odd-even with a warp per slot
    P.jacobi_oe = (e && e[0] == 'r') ? 0 : (e && e[0] == 'o') ? 1 : 2;
  }

  void* args[] = {&P};
  if (l_dtype == XMCA_F64)
    XMCA_CUDA(cudaLaunchCooperativeKernel((void*)varimax_kernel<double>, dim3(grid), dim3(VTHREADS),
                                          args, varimax_smem_bytes(), st));
  else
    XMCA_CUDA(cudaLaunchCooperativeKernel((void*)varimax_kernel<float>, dim3(grid), dim3(VTHREADS),
                                          args, varimax_smem_bytes(), st));
  XMCA_LAUNCHED();
  double h_out[4];
  XMCA_CUDA(cudaMemcpyAsync(h_out, d_out, sizeof h_out, cudaMemcpyDeviceToHost, st));
  XMCA_CUDA(cudaStreamSynchronize(st));
  if (iterations_out) *iterations_out = (int)h_out[0];
  if (h_out[1] == 0.0)
    return fail(XMCA_NOT_CONVERGED, "Rotation process did not converge.", __FILE__, __LINE__);
  return XMCA_OK;
}
