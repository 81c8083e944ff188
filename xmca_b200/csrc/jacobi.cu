// Blocked one-sided Jacobi SVD (Hestenes) in fp64 -- the engine's replacement
// for np.linalg.svd at array.py:479 / :570 and (for small n) rotation.py:59.
//
// The m x n matrix is held COLUMN-major in HBM.  Columns are grouped in blocks
// of W = 32; a round-robin tournament pairs the blocks so that every round
// works on n_blocks/2 disjoint 64-column panels in parallel:
//
//   pair_gram   G_p = X_p^T X_p           (64 x 64, split over row chunks)
//   pair_eig    G_p = R_p L R_p^T         (two-sided Jacobi in shared memory,
//                                          one CTA per panel, 1 barrier / step)
//   pair_apply  X_p <- X_p R_p, J_p <- J_p R_p   (streaming, in place)
//
// Per sweep the HBM traffic is 3 n^2 m e / W (+ the same for J); the fp64 flop
// count 8 m n^2 (+4 m n^2 for J) does not depend on W.  Convergence measure:
// the largest cosine |g_ij| / sqrt(g_ii g_jj) seen over all panels of a sweep
// (scaled criterion: clustered small singular values are resolved as well as
// the leading ones); columns whose norm is below 1e-11 of the largest count as
// zero.  Panels whose largest cosine is already below tol / 10 are skipped
// (no eigen-solve, no update), which makes the final, confirming sweep cheap.
#include "common.cuh"
#include <vector>
#include <cstring>
#include <math.h>
#include <stdlib.h>

namespace xmca {

constexpr int W = 32;          // block width
constexpr int P = 2 * W;       // panel width (64)
constexpr int GS = P * P;      // doubles per 64 x 64 matrix

__constant__ unsigned char c_rr[(P - 1) * W * 2];   // inner round-robin table for 64 indices

// ------------------------------------------------------------------ pair_gram
// grid (npairs, nsplit); block 256.  partial[(p*nsplit+z)][64][64]
__global__ void __launch_bounds__(256)
pair_gram_kernel(const double* __restrict__ Kc, int64_t ldk, int64_t m, const int* __restrict__ pairs,
                 int nsplit, double* __restrict__ partial) {
  __shared__ double Xs[P][W + 1];           // [col][row]
  const int p = blockIdx.x, z = blockIdx.y, tid = threadIdx.x;
  const int64_t cI = (int64_t)pairs[2 * p] * W, cJ = (int64_t)pairs[2 * p + 1] * W;
  const int64_t chunk = ((m + nsplit - 1) / nsplit + W - 1) / W * W;
  const int64_t r_begin = (int64_t)z * chunk, r_end = min(m, r_begin + chunk);
  const int tx = tid & 15, ty = tid >> 4;
  double acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.0;

  const int lr = tid & 31, lc = tid >> 5;   // loader: 32 rows x 8 columns per pass
  for (int64_t r0 = r_begin; r0 < r_end; r0 += W) {
#pragma unroll
    for (int pass = 0; pass < 8; ++pass) {
      int c = lc + 8 * pass;
      int64_t col = c < W ? cI + c : cJ + (c - W);
      int64_t r = r0 + lr;
      Xs[c][lr] = r < r_end ? Kc[col * ldk + r] : 0.0;
    }
    __syncthreads();
#pragma unroll 8
    for (int r = 0; r < W; ++r) {
      double a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        a[i] = Xs[ty + 16 * i][r];
        b[i] = Xs[tx + 16 * i][r];
      }
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fma(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
  double* out = partial + ((int64_t)p * nsplit + z) * GS;
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) out[(ty + 16 * i) * P + tx + 16 * j] = acc[i][j];
}

// ------------------------------------------------------------------- pair_eig
__device__ __forceinline__ void sym_rotation(double app, double aqq, double apq, double& c, double& s) {
  // Golub & Van Loan 8.4: J = [[c, s], [-s, c]],  J^T A J diagonal
  if (apq == 0.0 || fabs(apq) <= 1e-300) { c = 1.0; s = 0.0; return; }
  if (fabs(apq) <= 1.1e-16 * 0.5 * sqrt(fabs(app * aqq)) && app != aqq) { c = 1.0; s = 0.0; return; }
  double zeta = (aqq - app) / (2.0 * apq);
  double t = (zeta >= 0.0 ? 1.0 : -1.0) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
  c = rsqrt(1.0 + t * t);
  s = c * t;
}

// grid npairs; block 1024; dynamic smem 3 * 64*64 doubles (G ping-pong + R)
// Two-sided cyclic Jacobi on the 64 x 64 pivot Gram.  Per step the 32 disjoint
// rotations are computed ONCE (32 threads, fp64 div/sqrt are slow) and
// broadcast through shared memory; all 1024 threads then apply them to their
// 2 x 2 block of G (rows and columns) and to two rows of R.  At most
// `max_inner` sweeps: the pivot block only has to be diagonalised as far as the
// outer iteration can use (the off-block coupling is of the same size), and
// close to convergence one sweep is enough (quadratic).
__global__ void __launch_bounds__(1024)
pair_eig_kernel(const double* __restrict__ partial, int nsplit, double* __restrict__ Rout,
                unsigned long long* __restrict__ offmax_bits, int max_inner,
                const unsigned long long* __restrict__ gmax_bits, double skip_tol, int* __restrict__ skip) {
  extern __shared__ double sm[];
  double* Ga = sm;
  double* Gb = sm + GS;
  double* R = sm + 2 * GS;
  __shared__ double red[32];
  __shared__ double2 s_cs[W];
  __shared__ int s_perm[P];
  __shared__ int s_rot[2];
  const int tid = threadIdx.x, p = blockIdx.x;
  const int a = tid >> 5, b = tid & 31;

  // load + reduce partial Grams, init R = I
  for (int e = tid; e < GS; e += 1024) {
    double g = 0.0;
    const double* src = partial + (int64_t)p * nsplit * GS + e;
    for (int z = 0; z < nsplit; ++z) g += src[(int64_t)z * GS];
    int i = e >> 6, j = e & 63;
    Ga[e] = g;
    R[e] = (i == j) ? 1.0 : 0.0;
  }
  if (tid < 2) s_rot[tid] = 0;
  __syncthreads();
  // pre-rotation measure: largest cosine between two non-negligible columns of the panel
  const double zero2 = 1e-22 * __longlong_as_double((long long)*gmax_bits);
  double local_off = 0.0;
  for (int e = tid; e < GS; e += 1024) {
    int i = e >> 6, j = e & 63;
    if (i != j) {
      const double gi = Ga[i * P + i], gj = Ga[j * P + j];
      if (gi > zero2 && gj > zero2) local_off = fmax(local_off, fabs(Ga[e]) * rsqrt(gi * gj));
    }
  }
  local_off = warp_max(local_off);
  if ((tid & 31) == 0) red[tid >> 5] = local_off;
  __syncthreads();
  double pair_cos = 0.0;
  for (int i = 0; i < 32; ++i) pair_cos = fmax(pair_cos, red[i]);
  if (tid == 0) {
    atomicMax(offmax_bits, (unsigned long long)__double_as_longlong(pair_cos));
    skip[p] = pair_cos <= skip_tol;
  }
  if (pair_cos <= skip_tol) return;          // uniform: panel already orthogonal to working accuracy
  // symmetrise (partials are symmetric up to rounding order; enforce exactly)
  for (int e = tid; e < GS; e += 1024) {
    int i = e >> 6, j = e & 63;
    if (i < j) { double v = 0.5 * (Ga[i * P + j] + Ga[j * P + i]); Gb[i * P + j] = v; Gb[j * P + i] = v; }
    else if (i == j) Gb[e] = Ga[e];
  }
  __syncthreads();
  double* cur = Gb;
  double* nxt = Ga;

  for (int sweep = 0; sweep < max_inner; ++sweep) {
    if (tid == 0) s_rot[(sweep + 1) & 1] = 0;          // flag of the NEXT sweep (nobody reads it now)
    for (int step = 0; step < P - 1; ++step) {
      const unsigned char* tb = c_rr + step * P;
      if (tid < W) {
        const int pp = tb[2 * tid], qq = tb[2 * tid + 1];
        double c, s;
        sym_rotation(cur[pp * P + pp], cur[qq * P + qq], cur[pp * P + qq], c, s);
        s_cs[tid] = make_double2(c, s);
        if (s != 0.0) s_rot[sweep & 1] = 1;
      }
      __syncthreads();
      const int pa = tb[2 * a], qa = tb[2 * a + 1], pb = tb[2 * b], qb = tb[2 * b + 1];
      const double2 ra = s_cs[a], rb = s_cs[b];
      const double ca = ra.x, sa = ra.y, cb = rb.x, sb = rb.y;
      const double x00 = cur[pa * P + pb], x01 = cur[pa * P + qb];
      const double x10 = cur[qa * P + pb], x11 = cur[qa * P + qb];
      const double y00 = cb * x00 - sb * x01, y01 = sb * x00 + cb * x01;
      const double y10 = cb * x10 - sb * x11, y11 = sb * x10 + cb * x11;
      double z00 = ca * y00 - sa * y10, z10 = sa * y00 + ca * y10;
      double z01 = ca * y01 - sa * y11, z11 = sa * y01 + ca * y11;
      if (a == b) { z01 = 0.0; z10 = 0.0; }
      nxt[pa * P + pb] = z00; nxt[pa * P + qb] = z01;
      nxt[qa * P + pb] = z10; nxt[qa * P + qb] = z11;
      // eigenvector accumulation: rows 2a, 2a+1 of column pair b (in place)
#pragma unroll
      for (int rr = 0; rr < 2; ++rr) {
        const int row = 2 * a + rr;
        const double rp = R[row * P + pb], rq = R[row * P + qb];
        R[row * P + pb] = cb * rp - sb * rq;
        R[row * P + qb] = sb * rp + cb * rq;
      }
      __syncthreads();
      double* t = cur; cur = nxt; nxt = t;
    }
    if (s_rot[sweep & 1] == 0) break;                    // no rotation in this sweep: diagonal
  }
  __syncthreads();
  // sort eigenvalues descending (rank by counting), write permuted eigenvectors
  if (tid < P) {
    const double li = cur[tid * P + tid];
    int rank = 0;
    for (int j = 0; j < P; ++j) {
      const double lj = cur[j * P + j];
      rank += (lj > li) || (lj == li && j < tid);
    }
    s_perm[rank] = tid;
  }
  __syncthreads();
  double* out = Rout + (int64_t)p * GS;
  for (int e = tid; e < GS; e += 1024) {
    int k = e >> 6, c = e & 63;
    out[e] = R[k * P + s_perm[c]];
  }
}

// ----------------------------------------------------------------- pair_apply
// X[:, panel] <- X[:, panel] * R_p on a 64-row tile.  grid (npairs, row tiles); block 256.
__global__ void __launch_bounds__(256)
pair_apply_kernel(double* __restrict__ Xc, int64_t ld, int64_t rows, const int* __restrict__ pairs,
                  const double* __restrict__ Rall, const int* __restrict__ skip) {
  if (skip[blockIdx.x]) return;
  extern __shared__ double sm[];
  double* Xs = sm;                 // [col k][row]  stride 65
  double* Rs = sm + P * (P + 1);   // [k][c]        stride 64
  const int p = blockIdx.x, tid = threadIdx.x;
  const int64_t r0 = (int64_t)blockIdx.y * P;
  const int64_t cI = (int64_t)pairs[2 * p] * W, cJ = (int64_t)pairs[2 * p + 1] * W;
  const double* Rp = Rall + (int64_t)p * GS;
  for (int e = tid; e < GS; e += 256) Rs[e] = Rp[e];
  {
    const int lr = tid & 63, lc = tid >> 6;          // 64 rows x 4 columns per pass
#pragma unroll
    for (int pass = 0; pass < 16; ++pass) {
      int c = lc + 4 * pass;
      int64_t col = c < W ? cI + c : cJ + (c - W);
      int64_t r = r0 + lr;
      Xs[c * (P + 1) + lr] = r < rows ? Xc[col * ld + r] : 0.0;
    }
  }
  __syncthreads();
  const int tx = tid & 15, ty = tid >> 4;            // tx -> rows, ty -> cols
  double acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.0;
#pragma unroll 8
  for (int k = 0; k < P; ++k) {
    double x[4], r[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      x[i] = Xs[k * (P + 1) + tx + 16 * i];
      r[i] = Rs[k * P + ty + 16 * i];
    }
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[i][j] = fma(x[i], r[j], acc[i][j]);
  }
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    int c = ty + 16 * j;
    int64_t col = c < W ? cI + c : cJ + (c - W);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      int64_t r = r0 + tx + 16 * i;
      if (r < rows) Xc[col * ld + r] = acc[i][j];
    }
  }
}

// ------------------------------------------------------------------- helpers
__global__ void set_identity_kernel(double* __restrict__ Jc, int64_t ld, int64_t n) {
  int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n * n) return;
  int64_t c = idx / n, r = idx % n;
  Jc[c * ld + r] = (r == c) ? 1.0 : 0.0;
}

// one block per column: sumsq -> out[c]; also atomicMax of sumsq into gmax_bits
__global__ void col_norm_kernel(const double* __restrict__ Kc, int64_t ldk, int64_t m,
                                double* __restrict__ out, int take_sqrt,
                                unsigned long long* __restrict__ gmax_bits) {
  __shared__ double red[8];
  const double* col = Kc + (int64_t)blockIdx.x * ldk;
  double s = 0.0;
  for (int64_t r = threadIdx.x; r < m; r += blockDim.x) s = fma(col[r], col[r], s);
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int i = 0; i < (int)(blockDim.x >> 5); ++i) t += red[i];
    if (out) out[blockIdx.x] = take_sqrt ? sqrt(t) : t;
    if (gmax_bits) atomicMax(gmax_bits, (unsigned long long)__double_as_longlong(t));
  }
}

static void build_round_robin(int nb, std::vector<int>& table) {
  // nb even; rounds nb-1; each round nb/2 pairs (lo, hi)
  std::vector<int> idx(nb);
  for (int i = 0; i < nb; ++i) idx[i] = i;
  table.resize((size_t)(nb - 1) * nb);
  for (int r = 0; r < nb - 1; ++r) {
    for (int i = 0; i < nb / 2; ++i) {
      int x = idx[i], y = idx[nb - 1 - i];
      table[(size_t)r * nb + 2 * i] = x < y ? x : y;
      table[(size_t)r * nb + 2 * i + 1] = x < y ? y : x;
    }
    // rotate all but the first
    int last = idx[nb - 1];
    for (int i = nb - 1; i > 1; --i) idx[i] = idx[i - 1];
    idx[1] = last;
  }
}

static int g_rr_uploaded_device = -1;

static int upload_inner_table() {
  int dev = 0;
  XMCA_CUDA(cudaGetDevice(&dev));
  if (g_rr_uploaded_device == dev) return XMCA_OK;
  std::vector<int> t;
  build_round_robin(P, t);
  std::vector<unsigned char> bytes(t.size());
  for (size_t i = 0; i < t.size(); ++i) bytes[i] = (unsigned char)t[i];
  XMCA_CUDA(cudaMemcpyToSymbol(c_rr, bytes.data(), bytes.size()));
  XMCA_CUDA(cudaFuncSetAttribute(pair_eig_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 3 * GS * (int)sizeof(double)));
  XMCA_CUDA(cudaFuncSetAttribute(pair_apply_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (P * (P + 1) + GS) * (int)sizeof(double)));
  g_rr_uploaded_device = dev;
  return XMCA_OK;
}

struct JacobiPlan {
  int64_t n_pad; int nb, npairs, nsplit;
  size_t off_pairs, off_partial, off_R, off_scalars, off_skip, total;
};

static JacobiPlan make_plan(int64_t m, int64_t n) {
  JacobiPlan pl;
  int64_t nb = (n + W - 1) / W;
  if (nb < 2) nb = 2;
  if (nb & 1) ++nb;
  pl.nb = (int)nb; pl.n_pad = nb * W; pl.npairs = (int)(nb / 2);
  int64_t want = (2LL * sm_count() + pl.npairs - 1) / pl.npairs;     // >= 2 CTAs per SM
  int64_t maxsplit = (m + 4 * W - 1) / (4 * W);                       // >= 128 rows per split
  int64_t ns = want < maxsplit ? want : maxsplit;
  if (ns < 1) ns = 1;
  if (ns > 64) ns = 64;
  pl.nsplit = (int)ns;
  size_t o = 0;
  pl.off_pairs = o;   o += ((size_t)(nb - 1) * nb * sizeof(int) + 255) / 256 * 256;
  pl.off_partial = o; o += (size_t)pl.npairs * pl.nsplit * GS * sizeof(double);
  pl.off_R = o;       o += (size_t)pl.npairs * GS * sizeof(double);
  pl.off_scalars = o; o += 256;
  pl.off_skip = o;    o += ((size_t)pl.npairs * sizeof(int) + 255) / 256 * 256;
  pl.total = o;
  return pl;
}

}  // namespace xmca

using namespace xmca;

extern "C" int64_t xmca_jacobi_padded_cols(int64_t n) { return make_plan(1, n).n_pad; }

extern "C" size_t xmca_jacobi_workspace_bytes(int64_t m, int64_t n) { return make_plan(m, n).total; }

extern "C" int xmca_jacobi_svd(int64_t m, int64_t n, double* d_Kc, int64_t ldk,
                               double* d_Jc, int64_t ldj, double* d_sigma,
                               int max_sweeps, double tol, int* sweeps_out, double* offnorm_out,
                               void* d_workspace, size_t workspace_bytes, void* stream) {
  XMCA_REQUIRE(m > 0 && n > 0 && d_Kc && d_sigma && d_workspace, "xmca_jacobi_svd: bad argument");
  XMCA_REQUIRE(ldk >= m, "xmca_jacobi_svd: ldk < m");
  JacobiPlan pl = make_plan(m, n);
  XMCA_REQUIRE(workspace_bytes >= pl.total, "xmca_jacobi_svd: workspace too small");
  XMCA_REQUIRE(!d_Jc || ldj >= pl.n_pad, "xmca_jacobi_svd: ldj < padded n");
  XMCA_REQUIRE((m + P - 1) / P <= 65535 && (pl.n_pad + P - 1) / P <= 65535, "xmca_jacobi_svd: too large");
  if (max_sweeps <= 0) max_sweeps = 40;
  if (tol <= 0.0) tol = 1e-11;
  cudaStream_t st = (cudaStream_t)stream;
  int rc = upload_inner_table();
  if (rc != XMCA_OK) return rc;

  char* ws = reinterpret_cast<char*>(d_workspace);
  int* d_pairs = reinterpret_cast<int*>(ws + pl.off_pairs);
  double* d_partial = reinterpret_cast<double*>(ws + pl.off_partial);
  double* d_R = reinterpret_cast<double*>(ws + pl.off_R);
  unsigned long long* d_scal = reinterpret_cast<unsigned long long*>(ws + pl.off_scalars);
  int* d_skip = reinterpret_cast<int*>(ws + pl.off_skip);

  std::vector<int> table;
  build_round_robin(pl.nb, table);
  XMCA_CUDA(cudaMemcpyAsync(d_pairs, table.data(), table.size() * sizeof(int),
                            cudaMemcpyHostToDevice, st));
  // the host vector must outlive the async copy
  XMCA_CUDA(cudaStreamSynchronize(st));

  if (d_Jc) {
    int64_t tot = pl.n_pad * pl.n_pad;
    set_identity_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, st>>>(d_Jc, ldj, pl.n_pad);
    XMCA_LAUNCHED();
  }
  const size_t eig_smem = 3 * GS * sizeof(double);
  const size_t app_smem = (P * (P + 1) + GS) * sizeof(double);
  const double skip_tol = 0.1 * tol;
  static const int inner_sweeps = getenv("XMCA_JACOBI_INNER") ? atoi(getenv("XMCA_JACOBI_INNER")) : 1;
  static const bool trace = getenv("XMCA_JACOBI_TRACE") != nullptr;
  int sweeps = 0;
  double measure = INFINITY;
  bool converged = false;
  for (int sweep = 0; sweep < max_sweeps; ++sweep) {
    XMCA_CUDA(cudaMemsetAsync(d_scal, 0, 16, st));
    col_norm_kernel<<<(unsigned)pl.n_pad, 256, 0, st>>>(d_Kc, ldk, m, nullptr, 0, d_scal + 1);
    XMCA_LAUNCHED();
    for (int r = 0; r < pl.nb - 1; ++r) {
      const int* pr = d_pairs + (size_t)r * pl.nb;
      pair_gram_kernel<<<dim3(pl.npairs, pl.nsplit), 256, 0, st>>>(d_Kc, ldk, m, pr, pl.nsplit, d_partial);
      XMCA_LAUNCHED();
      pair_eig_kernel<<<pl.npairs, 1024, eig_smem, st>>>(d_partial, pl.nsplit, d_R, d_scal, inner_sweeps, d_scal + 1,
                                                         skip_tol, d_skip);
      XMCA_LAUNCHED();
      pair_apply_kernel<<<dim3(pl.npairs, (unsigned)((m + P - 1) / P)), 256, app_smem, st>>>(
          d_Kc, ldk, m, pr, d_R, d_skip);
      XMCA_LAUNCHED();
      if (d_Jc) {
        pair_apply_kernel<<<dim3(pl.npairs, (unsigned)((pl.n_pad + P - 1) / P)), 256, app_smem, st>>>(
            d_Jc, ldj, pl.n_pad, pr, d_R, d_skip);
        XMCA_LAUNCHED();
      }
    }
    unsigned long long h[2];
    XMCA_CUDA(cudaMemcpyAsync(h, d_scal, 16, cudaMemcpyDeviceToHost, st));
    XMCA_CUDA(cudaStreamSynchronize(st));
    double offmax, gmax;
    memcpy(&offmax, &h[0], 8);
    memcpy(&gmax, &h[1], 8);
    sweeps = sweep + 1;
    if (!(gmax > 0.0)) { measure = 0.0; converged = true; break; }   // zero matrix
    if (!isfinite(offmax) || !isfinite(gmax)) {
      if (sweeps_out) *sweeps_out = sweeps;
      return fail(XMCA_NUMERIC, "xmca_jacobi_svd: non-finite entries (SVD failed. NaN entries may be the problem.)",
                  __FILE__, __LINE__);
    }
    measure = offmax;                  // largest cosine BEFORE this sweep's rotations
    if (trace) fprintf(stderr, "[xmca jacobi] m=%lld n=%lld sweep %d: max cosine = %.3e\n", (long long)m, (long long)n, sweeps, measure);
    if (measure <= tol) { converged = true; break; }
  }
  col_norm_kernel<<<(unsigned)pl.n_pad, 256, 0, st>>>(d_Kc, ldk, m, d_sigma, 1, nullptr);
  XMCA_LAUNCHED();
  if (sweeps_out) *sweeps_out = sweeps;
  if (offnorm_out) *offnorm_out = measure;
  if (!converged) return fail(XMCA_NOT_CONVERGED, "xmca_jacobi_svd: sweep limit reached", __FILE__, __LINE__);
  return XMCA_OK;
}
