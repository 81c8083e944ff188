// Blocked Cholesky factorisation and triangular solve in fp64 -- the
// orthogonalisation step of the engine's "Cholesky-QR" route for T < S:
//
//   G_A = A A^T = L_A L_A^T,  G_B = B B^T = L_B L_B^T        (T x T)
//   C = A^T B = Q_A (L_A^T L_B) Q_B^T,  Q_X = X^T L_X^{-T} orthonormal,
//
// so the singular values of the S1 x S2 cross-covariance are those of the
// T x T matrix M = L_A^T L_B and ONE Jacobi SVD replaces the three the
// reference takes (array.py:479 twice, :570).  V = X^T (L_X^{-T} P) needs the
// triangular solve below.
//
// Right-looking, block size 64:  the 64 x 64 diagonal block is factorised and
// inverted by one CTA in shared memory; the panel solve and the trailing
// update are products on the general fp64 kernel (xmca_gemm), which is where
// all the n^3/3 work is.
#include "common.cuh"
#include "small64.cuh"
#include <math.h>
#include <stdlib.h>

namespace xmca {

constexpr int CB = 64;

// One CTA of 64 threads (two warps, synchronised by a named barrier).  A: n x n row-major (lda), block starting
// at (k0, k0) of size nb (<= 64).  Writes L_kk in place (strict upper part of the block zeroed) and inv(L_kk)
// (lower triangular, padded to 64 x 64 with an identity tail) to inv.  flag[0] is set to k0 + 1 + (failing
// column) when a pivot is not above min_pivot / finite.
// The block is on the critical path of the whole factorisation (n / 64 strictly sequential launches), so both
// parts are written for latency: thread i keeps ROW i of the block in registers and the factorisation is
// right-looking (per column: pivot, scale, then 63 - j independent FMAs per thread); the inverse is a
// column-oriented forward substitution with thread c holding column c of the right-hand side in registers.
__device__ __forceinline__ void bar64() { asm volatile("bar.sync 1, 64;" ::: "memory"); }

__global__ void __launch_bounds__(CB)
chol_diag_kernel(double* __restrict__ A, int64_t lda, int64_t k0, int nb, double* __restrict__ inv,
                 int* __restrict__ flag, double min_pivot) {
  __shared__ double Ls[CB][CB + 1];
  __shared__ double colbuf[CB], rdiag[CB];
  __shared__ double s_rinv;
  __shared__ int s_bad;
  const int tid = threadIdx.x;
  if (tid == 0) s_bad = 0;
  for (int e = tid; e < CB * CB; e += CB) {
    const int i = e >> 6, j = e & 63;
    Ls[i][j] = (i < nb && j < nb && j <= i) ? A[(k0 + i) * lda + k0 + j] : ((i == j) ? 1.0 : 0.0);
  }
  bar64();
  double row[CB];
#pragma unroll
  for (int k = 0; k < CB; ++k) row[k] = Ls[tid][k];
  bool bad = false;
#pragma unroll
  for (int j = 0; j < CB; ++j) {
    if (tid == j) {
      const double d = row[j];
      if (!(d > min_pivot) || !isfinite(d)) s_bad = j + 1;
      const double piv = sqrt(d);
      row[j] = piv;
      const double r = 1.0 / piv;
      s_rinv = r;
      rdiag[j] = r;
    }
    bar64();
    if (s_bad) { bad = true; break; }              // uniform
    if (tid > j) { row[j] *= s_rinv; colbuf[tid] = row[j]; }
    bar64();
    if (tid > j) {
      const double lij = row[j];
#pragma unroll
      for (int k = j + 1; k < CB; ++k)
        if (k <= tid) row[k] = fma(-lij, colbuf[k], row[k]);
    }
  }
  if (bad) {
    if (tid == 0) atomicCAS(flag, 0, (int)(k0 + s_bad));
    return;
  }
#pragma unroll
  for (int k = 0; k < CB; ++k) Ls[tid][k] = (k <= tid) ? row[k] : 0.0;
  bar64();
  // inverse: column c = tid of inv solves L x = e_c; x_k = s_k / L_kk, then s_i -= L_ik x_k for all i > k
  double sv[CB];
#pragma unroll
  for (int i = 0; i < CB; ++i) sv[i] = (i == tid) ? 1.0 : 0.0;
#pragma unroll
  for (int k = 0; k < CB; ++k) {
    const double xk = sv[k] * rdiag[k];
    sv[k] = xk;
#pragma unroll
    for (int i = k + 1; i < CB; ++i) sv[i] = fma(-Ls[i][k], xk, sv[i]);
  }
#pragma unroll
  for (int i = 0; i < CB; ++i) inv[i * CB + tid] = (i >= tid) ? sv[i] : 0.0;
  for (int e = tid; e < CB * CB; e += CB) {
    const int i = e >> 6, j = e & 63;
    if (i < nb && j < nb) A[(k0 + i) * lda + k0 + j] = Ls[i][j];
  }
}

// The same diagonal-block step with 256 threads and the blocked-by-8 factorisation of small64.cuh (the 64-thread
// kernel above keeps a row per thread in registers and runs its 64-step recurrences fully unrolled: ~40 us warm, bound
// by instruction fetch).  inv(L_kk) follows from the 8 x 8 diagonal-block inverses by block forward substitution,
// one warp per block column J:  Inv_JJ = Dinv_J,  Inv_IJ = -Dinv_I sum_{K = J}^{I - 1} L_IK Inv_KJ.
constexpr int CD_T = 256;

__global__ void __launch_bounds__(CD_T)
chol_diag_blocked_kernel(double* __restrict__ A, int64_t lda, int64_t k0, int nb, double* __restrict__ inv,
                         int* __restrict__ flag, double min_pivot, double* __restrict__ first_out, int first_rows) {
  pdl_enter();
  extern __shared__ double cd_smem[];
  double (*Ls)[PLD] = reinterpret_cast<double (*)[PLD]>(cd_smem);
  double (*Is)[PLD] = reinterpret_cast<double (*)[PLD]>(cd_smem + PB * PLD);
  __shared__ double s_dinv[512];
  __shared__ int s_fail, s_bad;
  const int tid = threadIdx.x;
  if (tid == 0) { s_fail = 0; s_bad = 0; }
  for (int e = tid; e < PB * PB; e += CD_T) {
    const int i = e >> 6, j = e & 63;
    Ls[i][j] = (i < nb && j < nb) ? (j <= i ? A[(k0 + i) * lda + k0 + j] : 0.0) : ((i == j) ? 1.0 : 0.0);
    Is[i][j] = 0.0;
    // copy of the first tile of the panel below (block (b + 1, b), still unsolved): every CTA of chol_panel_step_kernel
    // needs it while CTA 0 of that kernel overwrites the original in place
    if (first_out) first_out[e] = i < first_rows ? A[(k0 + PB + i) * lda + k0 + j] : 0.0;
  }
  __syncthreads();
  chol64_blocked(Ls, s_dinv, tid, &s_fail, min_pivot, &s_bad);
  __syncthreads();
  if (s_bad) {
    if (tid == 0) atomicCAS(flag, 0, (int)(k0 + s_bad));
    return;
  }
  {
    const int J = tid >> 5, lane = tid & 31;
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      const int e = lane + 32 * q, r = e >> 3, c = e & 7;
      Is[8 * J + r][8 * J + c] = s_dinv[J * 64 + r * 8 + c];
    }
    __syncwarp();
    for (int I = J + 1; I < 8; ++I) {
      double sv[2];
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        const int e = lane + 32 * q, r = e >> 3, c = e & 7;
        double acc = 0.0;
        for (int k = 8 * J; k < 8 * I; ++k) acc = fma(Ls[8 * I + r][k], Is[k][8 * J + c], acc);
        sv[q] = acc;
      }
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        const int e = lane + 32 * q;
        Is[8 * I + (e >> 3)][8 * J + (e & 7)] = sv[q];
      }
      __syncwarp();
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        const int e = lane + 32 * q, r = e >> 3, c = e & 7;
        double acc = 0.0;
#pragma unroll
        for (int t = 0; t < 8; ++t) acc = fma(s_dinv[I * 64 + r * 8 + t], Is[8 * I + t][8 * J + c], acc);
        sv[q] = -acc;
      }
      __syncwarp();
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        const int e = lane + 32 * q;
        Is[8 * I + (e >> 3)][8 * J + (e & 7)] = sv[q];
      }
      __syncwarp();
    }
  }
  __syncthreads();
  for (int e = tid; e < PB * PB; e += CD_T) {
    const int i = e >> 6, j = e & 63;
    inv[e] = Is[i][j];
    if (i < nb && j < nb) A[(k0 + i) * lda + k0 + j] = Ls[i][j];
  }
}

// One block step below the diagonal block in ONE launch (chain of the factorisation: diagonal block -> this kernel):
//   L_i = A_i inv(L_kk)^T            for the 64-row tile i of the panel (written in place and to the panel buffer),
//   A[i, next block column] -= L_i L_f^T,   L_f = the FIRST tile of the panel (rows k0 + 64 ..), which every CTA recomputes
// instead of waiting for the CTA that owns it (from a copy of the unsolved tile that the diagonal-block kernel made: the
// original is overwritten in place by CTA 0).  Three 64 x 64 x 64 products per CTA on the fp64 DMMA path (8 warps, warp
// tile 16 x 32), operands in shared memory at pitch 68.  Replaces three launches (panel product, copy back, update of the
// next block column) of the generic kernels.
constexpr int CF_T = 256, CF_LD = 68;

// acc[a][b][e]: C[i0 + 8 a + gid][j0 + 8 b + 2 tig + e], i0 = 16 (warp & 3), j0 = 32 (warp >> 2);
// C = sum_k X(i, k) Y(k, j), X(i, k) = X[i * xs_i + k * xs_k], Y(k, j) = Y[k * ys_k + j * ys_j], k < 64
__device__ __forceinline__ void cf_mm64(const double* X, int xs_i, int xs_k, const double* Y, int ys_k, int ys_j,
                                        double (&acc)[2][4][2]) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, gid = lane >> 2, tig = lane & 3;
  const int i0 = 16 * (w & 3), j0 = 32 * (w >> 2);
#pragma unroll
  for (int a = 0; a < 2; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b) { acc[a][b][0] = 0.0; acc[a][b][1] = 0.0; }
  const double* xp = X + (i0 + gid) * xs_i + tig * xs_k;
  const double* yp = Y + tig * ys_k + (j0 + gid) * ys_j;
#pragma unroll 4
  for (int kk = 0; kk < 16; ++kk) {
    double af[2], bf[4];
#pragma unroll
    for (int a = 0; a < 2; ++a) af[a] = xp[4 * kk * xs_k + 8 * a * xs_i];
#pragma unroll
    for (int b = 0; b < 4; ++b) bf[b] = yp[4 * kk * ys_k + 8 * b * ys_j];
#pragma unroll
    for (int a = 0; a < 2; ++a)
#pragma unroll
      for (int b = 0; b < 4; ++b)
        asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                     : "+d"(acc[a][b][0]), "+d"(acc[a][b][1]) : "d"(af[a]), "d"(bf[b]));
  }
}

__global__ void __launch_bounds__(CF_T)
chol_panel_step_kernel(double* __restrict__ A, int64_t lda, int64_t k0, int64_t rem, const double* __restrict__ inv,
                       const double* __restrict__ first, double* __restrict__ panel) {
  pdl_enter();
  extern __shared__ double cf_smem[];
  double* Xs = cf_smem;                       // tile i of the panel, then L_i      [64][68]
  double* Fs = Xs + 64 * CF_LD;               // first tile of the panel, then L_f  [64][68]
  double* Is = Fs + 64 * CF_LD;               // inv(L_kk), lower                   [64][68]
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5, gid = lane >> 2, tig = lane & 3;
  const int i0w = 16 * (w & 3), j0w = 32 * (w >> 2);
  const int64_t r0 = (int64_t)blockIdx.x * 64;                       // first panel row of this tile
  const int rows = (int)min((int64_t)64, rem - r0);
  const int rows_f = (int)min((int64_t)64, rem);
  const double* Ap = A + (k0 + 64) * lda + k0;                       // the panel: rem x 64, pitch lda
  for (int e = tid; e < 64 * 64; e += CF_T) {
    const int r = e >> 6, c = e & 63;
    Xs[r * CF_LD + c] = r < rows ? Ap[(r0 + r) * lda + c] : 0.0;
    Fs[r * CF_LD + c] = first[e];                 // (copy made by the diagonal-block kernel; rows >= rows_f are zero)
    Is[r * CF_LD + c] = inv[e];
  }
  __syncthreads();
  double li[2][4][2], lf[2][4][2];
  cf_mm64(Xs, CF_LD, 1, Is, 1, CF_LD, li);                           // L_i[r][c] = sum_k X[r][k] inv[c][k]
  cf_mm64(Fs, CF_LD, 1, Is, 1, CF_LD, lf);
  __syncthreads();
#pragma unroll
  for (int a = 0; a < 2; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b) {
      const int r = i0w + 8 * a + gid, c = j0w + 8 * b + 2 * tig;
      Xs[r * CF_LD + c] = li[a][b][0]; Xs[r * CF_LD + c + 1] = li[a][b][1];
      Fs[r * CF_LD + c] = lf[a][b][0]; Fs[r * CF_LD + c + 1] = lf[a][b][1];
      if (r < rows) {
        double* ai = A + (k0 + 64 + r0 + r) * lda + k0 + c;
        ai[0] = li[a][b][0]; ai[1] = li[a][b][1];
        double* pp = panel + (r0 + r) * 64 + c;
        pp[0] = li[a][b][0]; pp[1] = li[a][b][1];
      }
    }
  __syncthreads();
  // update of the next block column: A[k0 + 64 + r0 + r][k0 + 64 + c] -= sum_k L_i[r][k] L_f[c][k],  c < min(64, rem)
  double u[2][4][2];
  cf_mm64(Xs, CF_LD, 1, Fs, 1, CF_LD, u);
#pragma unroll
  for (int a = 0; a < 2; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b) {
      const int r = i0w + 8 * a + gid, c = j0w + 8 * b + 2 * tig;
      if (r < rows) {
        double* an = A + (k0 + 64 + r0 + r) * lda + k0 + 64 + c;
        if (c < rows_f) an[0] -= u[a][b][0];
        if (c + 1 < rows_f) an[1] -= u[a][b][1];
      }
    }
}

// zero the strict upper triangle of an n x n row-major matrix
__global__ void zero_upper_kernel(double* __restrict__ A, int64_t lda, int64_t n) {
  const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= n) return;
  for (int64_t r = blockIdx.y; r < n; r += gridDim.y)
    if (c > r) A[r * lda + c] = 0.0;
}

// Y (rows x cols, ldy) = X (ldx)
__global__ void copy_block_kernel(const double* __restrict__ X, int64_t ldx, double* __restrict__ Y,
                                  int64_t ldy, int64_t rows, int64_t cols) {
  const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= cols) return;
  for (int64_t r = blockIdx.y; r < rows; r += gridDim.y) Y[r * ldy + c] = X[r * ldx + c];
}

static inline unsigned yblocks(int64_t rows) {
  int64_t want = 8LL * sm_count();
  int64_t g = rows < want ? rows : want;
  return (unsigned)(g < 1 ? 1 : (g > 65535 ? 65535 : g));
}

static int copy_block(const double* X, int64_t ldx, double* Y, int64_t ldy, int64_t rows, int64_t cols,
                      cudaStream_t st) {
  if (rows <= 0 || cols <= 0) return XMCA_OK;
  dim3 grid((unsigned)((cols + 127) / 128), yblocks(rows));
  copy_block_kernel<<<grid, 128, 0, st>>>(X, ldx, Y, ldy, rows, cols);
  XMCA_LAUNCHED();
  return XMCA_OK;
}

// internal high-priority stream + events of the look-ahead, one set per (device, caller's stream), created on first use
struct CholStreams { cudaStream_t owner = nullptr; bool used = false; cudaStream_t chain = nullptr; cudaEvent_t ev_panel = nullptr, ev_bulk = nullptr; };
static CholStreams* chol_streams(cudaStream_t caller) {
  constexpr int SLOTS = 4;
  static CholStreams pool[64][SLOTS];
  static int next_slot[64];
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return nullptr;
  CholStreams* s = nullptr;
  for (int i = 0; i < SLOTS; ++i)
    if (pool[dev][i].used && pool[dev][i].owner == caller) s = &pool[dev][i];
  if (!s) {
    s = &pool[dev][next_slot[dev]];
    next_slot[dev] = (next_slot[dev] + 1) % SLOTS;
    s->owner = caller;
    s->used = true;
  }
  if (!s->chain) {
    int lo = 0, hi = 0;
    cudaDeviceGetStreamPriorityRange(&lo, &hi);
    if (cudaStreamCreateWithPriority(&s->chain, cudaStreamNonBlocking, hi) != cudaSuccess) { s->chain = nullptr; return nullptr; }
    cudaEventCreateWithFlags(&s->ev_panel, cudaEventDisableTiming);
    cudaEventCreateWithFlags(&s->ev_bulk, cudaEventDisableTiming);
  }
  return s;
}

struct CholPlan { int64_t nblk; size_t off_inv, off_panel, off_flag, total; };

static CholPlan chol_plan(int64_t n, int64_t nrhs) {
  CholPlan p;
  p.nblk = (n + CB - 1) / CB;
  size_t o = 0;
  p.off_inv = o;   o += (size_t)p.nblk * CB * CB * sizeof(double);
  int64_t w = nrhs > CB ? nrhs : CB;
  p.off_panel = o; o += (2 * (size_t)n * CB + (size_t)CB * w) * sizeof(double);   // two panel buffers (look-ahead)
  p.off_flag = o;  o += 256 + (size_t)CB * CB * sizeof(double);          // flag + copy of the first panel tile (fused step)
  p.total = o;
  return p;
}

}  // namespace xmca

using namespace xmca;

extern "C" size_t xmca_cholesky_workspace_bytes(int64_t n) { return chol_plan(n, CB).total; }
extern "C" size_t xmca_cholesky_invdiag_bytes(int64_t n) {
  return (size_t)((n + CB - 1) / CB) * CB * CB * sizeof(double);
}

extern "C" int xmca_cholesky(int64_t n, double* d_A, int64_t lda, double* d_invdiag, double min_pivot,
                             int* info_out, void* d_workspace, size_t workspace_bytes, void* stream) {
  XMCA_REQUIRE(n > 0 && d_A && d_invdiag && d_workspace, "xmca_cholesky: bad argument");
  XMCA_REQUIRE(lda >= n, "xmca_cholesky: lda < n");
  CholPlan pl = chol_plan(n, CB);
  XMCA_REQUIRE(workspace_bytes >= pl.total, "xmca_cholesky: workspace too small");
  cudaStream_t st = (cudaStream_t)stream;
  char* ws = reinterpret_cast<char*>(d_workspace);
  double* panels[2] = {reinterpret_cast<double*>(ws + pl.off_panel), reinterpret_cast<double*>(ws + pl.off_panel) + (size_t)n * CB};
  int* flag = reinterpret_cast<int*>(ws + pl.off_flag);
  XMCA_CUDA(cudaMemsetAsync(flag, 0, sizeof(int), st));
  // Look-ahead: the critical chain (diagonal block -> panel -> update of the NEXT block column) runs on an
  // internal HIGH-priority stream; the rest of each trailing update (block columns >= b + 2) stays on the caller's
  // stream, one step behind, under the next step's chain.  The stream / event objects are kept per (device, caller's
  // stream): creating and destroying them per call costs more than a block step and synchronises the device.
  CholStreams* cs = chol_streams(st);
  XMCA_REQUIRE(cs != nullptr, "xmca_cholesky: cannot create the internal stream");
  cudaStream_t chain = cs->chain;
  cudaEvent_t ev_panel = cs->ev_panel, ev_bulk = cs->ev_bulk;
  auto cleanup = [&]() { cudaStreamSynchronize(chain); };     // (the stream / events are kept for the next call)
  int rc = XMCA_OK;
  const size_t cd_smem_bytes = 2 * PB * PLD * sizeof(double);
  const char* cd_mode = getenv("XMCA_CHOL_DIAG");                 // "rows": the 64-thread row-per-thread kernel (A/B runs)
  const bool blocked_diag = !(cd_mode && cd_mode[0] == 'r');
  const char* cs_mode = getenv("XMCA_CHOL_STEP");                 // "generic": panel product / copy / update as three launches
  const bool fused_step = blocked_diag && !(cs_mode && cs_mode[0] == 'g');
  double* first_tile = reinterpret_cast<double*>(ws + pl.off_flag) + 16;      // 64 x 64 behind the flag
  const size_t cf_smem_bytes = 3 * 64 * CF_LD * sizeof(double);
  if (fused_step && cudaFuncSetAttribute(chol_panel_step_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)cf_smem_bytes) != cudaSuccess) rc = XMCA_CUDA_ERROR;
  if (blocked_diag && cudaFuncSetAttribute(chol_diag_blocked_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           (int)cd_smem_bytes) != cudaSuccess) rc = XMCA_CUDA_ERROR;
  // the chain starts after everything already queued on the caller's stream (the matrix, the flag reset)
  if (cudaEventRecord(ev_bulk, st) != cudaSuccess || cudaStreamWaitEvent(chain, ev_bulk, 0) != cudaSuccess) rc = XMCA_CUDA_ERROR;
  bool bulk_pending = false;
  for (int64_t b = 0; b < pl.nblk && rc == XMCA_OK; ++b) {
    const int64_t k0 = b * CB;
    const int nb = (int)((n - k0) < CB ? (n - k0) : CB);
    double* inv = d_invdiag + (size_t)b * CB * CB;
    double* panel = panels[b & 1];
    const int64_t rem_b = n - k0 - nb;
    const bool fuse = fused_step && nb == CB && rem_b > 0;
    if (blocked_diag)
      launch_pdl(true, chol_diag_blocked_kernel, dim3(1), dim3(CD_T), cd_smem_bytes, chain, d_A, lda, k0, nb, inv, flag,
                 min_pivot > 0.0 ? min_pivot : 0.0, fuse ? first_tile : nullptr, (int)(rem_b < CB ? rem_b : CB));
    else
      chol_diag_kernel<<<1, CB, 0, chain>>>(d_A, lda, k0, nb, inv, flag, min_pivot > 0.0 ? min_pivot : 0.0);
    if (cudaGetLastError() != cudaSuccess) { rc = XMCA_CUDA_ERROR; break; }
    g_launches.fetch_add(1, std::memory_order_relaxed);
    const int64_t rem = n - k0 - nb;
    if (rem <= 0) break;
    double* Aik = d_A + (k0 + nb) * lda + k0;
    double* Att = d_A + (k0 + nb) * lda + (k0 + nb);
    const int64_t c1 = rem < CB ? rem : CB;
    if (fuse) {
      // the previous step's bulk update also touched block column b + 1: it has to land first
      if (bulk_pending && cudaStreamWaitEvent(chain, ev_bulk, 0) != cudaSuccess) { rc = XMCA_CUDA_ERROR; break; }
      launch_pdl(true, chol_panel_step_kernel, dim3((unsigned)((rem + 63) / 64)), dim3(CF_T), cf_smem_bytes, chain, d_A, lda, k0, rem,
                 inv, first_tile, panel);
      if (cudaGetLastError() != cudaSuccess) { rc = XMCA_CUDA_ERROR; break; }
      g_launches.fetch_add(1, std::memory_order_relaxed);
      if (cudaEventRecord(ev_panel, chain) != cudaSuccess) { rc = XMCA_CUDA_ERROR; break; }
    } else {
      // panel: L_ik = A_ik inv(L_kk)^T   (rem x nb) -> workspace, then back in place
      rc = xmca_gemm(1, 1, rem, nb, nb, 1.0, Aik, XMCA_F64, lda, inv, XMCA_F64, CB, panel, XMCA_F64, CB, 0,
                     XMCA_F64, 1, nullptr, 0, (void*)chain);
      if (rc != XMCA_OK) break;
      if ((rc = copy_block(panel, CB, Aik, lda, rem, nb, chain)) != XMCA_OK) break;
      if (cudaEventRecord(ev_panel, chain) != cudaSuccess) { rc = XMCA_CUDA_ERROR; break; }
      // the previous step's bulk update also touched block column b + 1: it has to land first
      if (bulk_pending && cudaStreamWaitEvent(chain, ev_bulk, 0) != cudaSuccess) { rc = XMCA_CUDA_ERROR; break; }
      // trailing update  A_ij -= L_ik L_jk^T  (tiles on or below the diagonal): the next block column on the chain ...
      rc = xmca_gemm_ex(1, 1, rem, c1, nb, -1.0, panel, XMCA_F64, CB, panel, XMCA_F64, CB, Att, XMCA_F64, lda, 1,
                        XMCA_F64, 1, nullptr, 0, 0, (void*)chain);       // (the strict upper part is zeroed at the end)
      if (rc != XMCA_OK) break;
    }
    // ... and the remaining block columns on the caller's stream
    const int64_t rem2 = rem - c1;
    if (rem2 > 0) {
      if (cudaStreamWaitEvent(st, ev_panel, 0) != cudaSuccess) { rc = XMCA_CUDA_ERROR; break; }
      rc = xmca_gemm_ex(1, 1, rem2, rem2, nb, -1.0, panel + c1 * CB, XMCA_F64, CB, panel + c1 * CB, XMCA_F64, CB,
                        Att + c1 * lda + c1, XMCA_F64, lda, 1, XMCA_F64, 1, nullptr, 0, XMCA_GEMM_LOWER_ONLY, stream);
      if (rc != XMCA_OK) break;
      if (cudaEventRecord(ev_bulk, st) != cudaSuccess) { rc = XMCA_CUDA_ERROR; break; }
      bulk_pending = true;
    }
  }
  // join: the caller's stream continues after the chain
  if (rc == XMCA_OK && (cudaEventRecord(ev_panel, chain) != cudaSuccess || cudaStreamWaitEvent(st, ev_panel, 0) != cudaSuccess))
    rc = XMCA_CUDA_ERROR;
  if (rc != XMCA_OK) {
    cleanup();
    return rc == XMCA_CUDA_ERROR ? fail(XMCA_CUDA_ERROR, "xmca_cholesky: CUDA error in the block loop", __FILE__, __LINE__) : rc;
  }
  {
    dim3 grid((unsigned)((n + 127) / 128), yblocks(n));
    zero_upper_kernel<<<grid, 128, 0, st>>>(d_A, lda, n);
    XMCA_LAUNCHED();
  }
  int h_flag = 0;
  cudaError_t ce = cudaMemcpyAsync(&h_flag, flag, sizeof(int), cudaMemcpyDeviceToHost, st);
  if (ce == cudaSuccess) ce = cudaStreamSynchronize(st);
  cleanup();
  XMCA_CUDA(ce);
  if (info_out) *info_out = h_flag;
  if (h_flag != 0)
    return fail(XMCA_NUMERIC, "xmca_cholesky: matrix is not (numerically) positive definite", __FILE__, __LINE__);
  return XMCA_OK;
}

extern "C" size_t xmca_trsm_workspace_bytes(int64_t n, int64_t nrhs) { return chol_plan(n, nrhs).total; }

// Solve L^T W = R in place (R: n x nrhs row-major, ldr), L lower triangular from
// xmca_cholesky with its inverted diagonal blocks.
extern "C" int xmca_trsm_lt(int64_t n, int64_t nrhs, const double* d_L, int64_t ldl,
                            const double* d_invdiag, double* d_R, int64_t ldr,
                            void* d_workspace, size_t workspace_bytes, void* stream) {
  XMCA_REQUIRE(n > 0 && nrhs > 0 && d_L && d_invdiag && d_R && d_workspace, "xmca_trsm_lt: bad argument");
  XMCA_REQUIRE(ldl >= n && ldr >= nrhs, "xmca_trsm_lt: leading dimension too small");
  CholPlan pl = chol_plan(n, nrhs);
  XMCA_REQUIRE(workspace_bytes >= pl.total, "xmca_trsm_lt: workspace too small");
  cudaStream_t st = (cudaStream_t)stream;
  double* tmp = reinterpret_cast<double*>(reinterpret_cast<char*>(d_workspace) + pl.off_panel);  // 64 x nrhs
  for (int64_t b = pl.nblk - 1; b >= 0; --b) {
    const int64_t k0 = b * CB;
    const int nb = (int)((n - k0) < CB ? (n - k0) : CB);
    const double* inv = d_invdiag + (size_t)b * CB * CB;
    double* Rk = d_R + k0 * ldr;
    // W_k = inv(L_kk)^T R_k : A[m][kk] = inv[kk][m] (stored K x M)
    int rc = xmca_gemm(0, 0, nb, nrhs, nb, 1.0, inv, XMCA_F64, CB, Rk, XMCA_F64, ldr, tmp, XMCA_F64, nrhs, 0,
                       XMCA_F64, 1, nullptr, 0, stream);
    if (rc != XMCA_OK) return rc;
    if ((rc = copy_block(tmp, nrhs, Rk, ldr, nb, nrhs, st)) != XMCA_OK) return rc;
    if (k0 == 0) break;
    // R_i -= L_ki^T W_k for all i < k :  A[m][kk] = L[k0 + kk][m]  (stored K x M, lda = ldl)
    rc = xmca_gemm(0, 0, k0, nrhs, nb, -1.0, d_L + k0 * ldl, XMCA_F64, ldl, tmp, XMCA_F64, nrhs, d_R, XMCA_F64,
                   ldr, 1, XMCA_F64, 1, nullptr, 0, stream);
    if (rc != XMCA_OK) return rc;
  }
  return XMCA_OK;
}
