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

namespace cg = cooperative_groups;

namespace xmca {

constexpr int VP = 64;          // padded number of rotated modes (p <= 64)
constexpr int VT = 32;          // rows per tile
constexpr int VTHREADS = 512;
constexpr int VSLOT = VP * VP + VP;   // doubles per partial: T1 (64x64) + c (64)

struct VarimaxParams {
  const void* L; int ldt; int64_t n; int p; int64_t ldl;
  double gamma; int max_iter; double tol;
  void* An;            // n x p normalised loadings, same dtype as L, ld = p
  double* h;           // n row norms
  double* partial;     // [2*grid][VSLOT]
  double* reduced;     // [VSLOT]
  double* B; int64_t ldb; double* R; double* out;   // out: [0]=iterations [1]=converged [2]=d [3]=svd sweeps total
};

__device__ __forceinline__ void block_matmul_64(const double* X, const double* Y, double* Z, int p,
                                                bool y_transposed) {
  // Z[i][j] = sum_k X[i][k] * (y_transposed ? Y[j][k] : Y[k][j]),  i,j,k < p; padded entries -> 0
  for (int e = threadIdx.x; e < VP * VP; e += blockDim.x) {
    int i = e >> 6, j = e & 63;
    double s = 0.0;
    if (i < p && j < p) {
      if (y_transposed) for (int k = 0; k < p; ++k) s = fma(X[i * VP + k], Y[j * VP + k], s);
      else              for (int k = 0; k < p; ++k) s = fma(X[i * VP + k], Y[k * VP + j], s);
    }
    Z[e] = s;
  }
}

// One-sided Jacobi on the columns of X (p x p, stride VP), accumulating V.
// pe = p rounded up to even (column p is a zero column when p is odd).
// rr: round-robin table [(pe-1)][pe].  Returns number of sweeps used.
__device__ int polar_jacobi(double* X, double* V, int pe, const unsigned char* rr, int* s_flag) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  int sweeps = 0;
  for (; sweeps < 40; ++sweeps) {
    if (threadIdx.x == 0) *s_flag = 0;
    __syncthreads();
    for (int step = 0; step < pe - 1; ++step) {
      for (int pr = warp; pr < pe / 2; pr += nwarps) {
        const int cp = rr[step * pe + 2 * pr], cq = rr[step * pe + 2 * pr + 1];
        double xp0 = X[lane * VP + cp], xp1 = X[(lane + 32) * VP + cp];
        double xq0 = X[lane * VP + cq], xq1 = X[(lane + 32) * VP + cq];
        double al = warp_sum(xp0 * xp0 + xp1 * xp1);
        double be = warp_sum(xq0 * xq0 + xq1 * xq1);
        double ga = warp_sum(xp0 * xq0 + xp1 * xq1);
        if (fabs(ga) > 1e-14 * sqrt(al * be) && fabs(ga) > 1e-300) {
          double zeta = (be - al) / (2.0 * ga);
          double t = (zeta >= 0.0 ? 1.0 : -1.0) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
          double c = rsqrt(1.0 + t * t), s = c * t;
          X[lane * VP + cp] = c * xp0 - s * xq0;        X[lane * VP + cq] = s * xp0 + c * xq0;
          X[(lane + 32) * VP + cp] = c * xp1 - s * xq1; X[(lane + 32) * VP + cq] = s * xp1 + c * xq1;
          double vp0 = V[lane * VP + cp], vp1 = V[(lane + 32) * VP + cp];
          double vq0 = V[lane * VP + cq], vq1 = V[(lane + 32) * VP + cq];
          V[lane * VP + cp] = c * vp0 - s * vq0;        V[lane * VP + cq] = s * vp0 + c * vq0;
          V[(lane + 32) * VP + cp] = c * vp1 - s * vq1; V[(lane + 32) * VP + cq] = s * vp1 + c * vq1;
          if (lane == 0) *s_flag = 1;
        }
      }
      __syncthreads();
    }
    if (*s_flag == 0) { ++sweeps; break; }
    __syncthreads();
  }
  return sweeps;
}

template <typename TS>
__global__ void __launch_bounds__(VTHREADS, 1) varimax_kernel(VarimaxParams P) {
  cg::grid_group grid = cg::this_grid();
  extern __shared__ double sm[];
  double* Rs = sm;                    // rotation (64x64)
  double* Vs = Rs + VP * VP;          // right singular vectors, warm start
  double* Gs = Vs + VP * VP;          // A^T A
  double* Xs = Gs + VP * VP;          // SVD work matrix
  double* Ws = Xs + VP * VP;          // scratch (G R, U)
  double* As = Ws + VP * VP;          // tile [32][64]
  double* Bs = As + VT * VP;          // tile [32][64]
  double* cs = Bs + VT * VP;          // [64] column sums / sigma
  unsigned char* rr = reinterpret_cast<unsigned char*>(cs + VP);   // [(pe-1)*pe]
  __shared__ int s_flag;

  const int tid = threadIdx.x, p = P.p, pe = (p + 1) & ~1;
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
    Rs[e] = id; Vs[e] = (i == j) ? 1.0 : 0.0;
  }

  // T1-phase thread mapping: 4x4 register block of the 64x64 accumulator, two row-halves
  const int half = tid >> 8, bi = (tid & 255) >> 4, bj = tid & 15;
  // B-phase mapping: column j, 4 rows
  const int jcol = tid & 63, rg = tid >> 6;

  double acc[4][4];
  double csq;

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
  for (int64_t e = (int64_t)blockIdx.x * VTHREADS + tid; e < VSLOT; e += (int64_t)gridDim.x * VTHREADS) {
    double s = 0.0;
    for (int k = 0; k < 2 * (int)gridDim.x; ++k) s += P.partial[(int64_t)k * VSLOT + e];
    P.reduced[e] = s;
  }
  __threadfence();
  grid.sync();
  for (int e = tid; e < VP * VP; e += VTHREADS) Gs[e] = P.reduced[e];
  __syncthreads();

  // ---------------- fixed-point iteration ----------------
  double d = 0.0;
  int it = 0, converged = 0, svd_sweeps = 0;
  for (it = 1; it <= P.max_iter; ++it) {
    const double d_old = d;
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[i][j] = 0.0;
    csq = 0.0;
    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
      const int64_t r0 = tile * VT;
      __syncthreads();
      for (int e = tid; e < VT * VP; e += VTHREADS) {
        int r = e >> 6, c = e & 63;
        int64_t row = r0 + r;
        As[e] = (row < n && c < p) ? (double)An[row * p + c] : 0.0;
      }
      __syncthreads();
      {   // b = a R for 4 rows x 1 column
        double b4[4] = {0.0, 0.0, 0.0, 0.0};
        if (jcol < p) {
          for (int k = 0; k < p; ++k) {
            const double rk = Rs[k * VP + jcol];
#pragma unroll
            for (int q = 0; q < 4; ++q) b4[q] = fma(As[(rg * 4 + q) * VP + k], rk, b4[q]);
          }
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          csq = fma(b4[q], b4[q], csq);
          Bs[(rg * 4 + q) * VP + jcol] = b4[q] * b4[q] * b4[q];
        }
      }
      __syncthreads();
#pragma unroll 4
      for (int r = half * 16; r < half * 16 + 16; ++r) {
        double a[4], b[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) { a[i] = As[r * VP + bi + 16 * i]; b[i] = Bs[r * VP + bj + 16 * i]; }
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[i][j] = fma(a[i], b[j], acc[i][j]);
      }
    }
    // per-CTA partials: T1 blocks (two halves) and column sums of b^2
    __syncthreads();
    As[rg * VP + jcol] = csq;           // 8 row-groups x 64 columns
    __syncthreads();
    {
      double* slot = P.partial + ((int64_t)blockIdx.x * 2 + half) * VSLOT;
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) slot[(bi + 16 * i) * VP + bj + 16 * j] = acc[i][j];
      if ((tid & 255) < VP) {
        double s = 0.0;
        if (half == 0) for (int g = 0; g < VTHREADS / 64; ++g) s += As[g * VP + (tid & 255)];
        slot[VP * VP + (tid & 255)] = s;
      }
    }
    __threadfence();
    grid.sync();
    for (int64_t e = (int64_t)blockIdx.x * VTHREADS + tid; e < VSLOT; e += (int64_t)gridDim.x * VTHREADS) {
      double s = 0.0;
      for (int k = 0; k < 2 * (int)gridDim.x; ++k) s += P.partial[(int64_t)k * VSLOT + e];
      P.reduced[e] = s;
    }
    __threadfence();
    grid.sync();

    // ---- phase 2 (redundant on every CTA): T, polar factor, convergence ----
    if (tid < VP) cs[tid] = P.reduced[VP * VP + tid];
    block_matmul_64(Gs, Rs, Ws, p, false);               // Ws = G R
    __syncthreads();
    const double gn = P.gamma / (double)n;
    for (int e = tid; e < VP * VP; e += VTHREADS) {
      int i = e >> 6, j = e & 63;
      Rs[e] = (i < p && j < p) ? P.reduced[e] - gn * Ws[e] * cs[j] : 0.0;    // Rs now holds T
    }
    __syncthreads();
    block_matmul_64(Rs, Vs, Xs, p, false);               // X = T V   (warm start)
    __syncthreads();
    svd_sweeps += polar_jacobi(Xs, Vs, pe, rr, &s_flag);
    __syncthreads();
    // sigma_j = ||x_j||, d = sum sigma, U = X / sigma
    {
      const int warp = tid >> 5, lane = tid & 31;
      for (int j = warp; j < VP; j += VTHREADS / 32) {
        double x0 = Xs[lane * VP + j], x1 = Xs[(lane + 32) * VP + j];
        double nn = sqrt(warp_sum(x0 * x0 + x1 * x1));
        if (lane == 0) cs[j] = (j < p) ? nn : 0.0;
      }
    }
    __syncthreads();
    for (int e = tid; e < VP * VP; e += VTHREADS) {
      int j = e & 63;
      Ws[e] = (j < p && cs[j] > 0.0) ? Xs[e] / cs[j] : 0.0;                // U
    }
    __syncthreads();
    block_matmul_64(Ws, Vs, Rs, p, true);                 // R = U V^T
    d = 0.0;
    for (int j = 0; j < p; ++j) d += cs[j];
    __syncthreads();
    if (fabs(d - d_old) / d < P.tol) { converged = 1; break; }
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
        const double rk = Rs[k * VP + jcol];
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
    for (int e = tid; e < p * p; e += VTHREADS) P.R[e] = Rs[(e / p) * VP + (e % p)];
    if (tid == 0) { P.out[0] = (double)it; P.out[1] = (double)converged; P.out[2] = d; P.out[3] = (double)svd_sweeps; }
  }
}

static size_t varimax_smem_bytes() {
  return (size_t)(5 * VP * VP + 2 * VT * VP + VP) * sizeof(double) + (size_t)VP * VP;
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
  P.B = d_B; P.ldb = ldb; P.R = d_R; P.out = d_out;

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
