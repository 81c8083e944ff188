// General dense product on the FP64 / FP32 CUDA cores.
//
// Used wherever the reference calls `@` on operands that must keep fp64
// accuracy (the T x T "kernel" of array.py:556-566, the back-projection of
// array.py:584, the Promax fit of rotation.py:128-147, fp64 fields of config 5)
// or whose shape is too skinny for the tensor-core tile (projections onto a few
// modes, array.py:640/:667).  tcgen05 has no fp64 kind, so this is the fp64
// path of the engine; the fp32 cross-covariance goes through gemm_tc.cu.
//
// Tile 128 x 128 x 16, 256 threads, 8 x 8 register micro-tile, register
// prefetch of the next k-slab, optional split-K with a deterministic second
// pass (no atomics -> bit-reproducible).
#include "common.cuh"

namespace xmca {

constexpr int BM = 128, BN = 128, BK = 16, PAD = 1, NT = 256;

template <typename T>
__device__ __forceinline__ T ld_elem(const void* p, int dt, int64_t i) {
  return dt == XMCA_F64 ? (T) reinterpret_cast<const double*>(p)[i]
                        : (T) reinterpret_cast<const float*>(p)[i];
}

// Loads one BM x BK (or BN x BK) operand slab into registers.
// KMAJOR: operand stored [mn][k] (k contiguous); else stored [k][mn].
template <typename T, bool KMAJOR>
__device__ __forceinline__ void load_slab(T (&r)[8], const void* p, int dt, int64_t ld,
                                          int64_t mn0, int64_t mn_end, int64_t k0, int64_t k_end,
                                          int tid) {
  if (KMAJOR) {
    const int kk = tid & 15, m = tid >> 4;       // 16 consecutive k per row: one 128B line (fp64)
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      int64_t mm = mn0 + m + 16 * i, k = k0 + kk;
      r[i] = (mm < mn_end && k < k_end) ? ld_elem<T>(p, dt, mm * ld + k) : T(0);
    }
  } else {
    const int m = tid & 127, kk = tid >> 7;      // 128 consecutive mn per k row
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      int64_t mm = mn0 + m, k = k0 + kk + 2 * i;
      r[i] = (mm < mn_end && k < k_end) ? ld_elem<T>(p, dt, k * ld + mm) : T(0);
    }
  }
}

template <typename T, bool KMAJOR>
__device__ __forceinline__ void store_slab(const T (&r)[8], T (*s)[BM + PAD], int tid) {
  if (KMAJOR) {
    const int kk = tid & 15, m = tid >> 4;
#pragma unroll
    for (int i = 0; i < 8; ++i) s[kk][m + 16 * i] = r[i];
  } else {
    const int m = tid & 127, kk = tid >> 7;
#pragma unroll
    for (int i = 0; i < 8; ++i) s[kk + 2 * i][m] = r[i];
  }
}

template <typename T, bool AK, bool BKM>
__global__ void __launch_bounds__(NT)
gemm_simt_kernel(int64_t M, int64_t N, int64_t K, double alpha,
                 const void* __restrict__ A, int adt, int64_t lda,
                 const void* __restrict__ B, int bdt, int64_t ldb,
                 void* __restrict__ D, int ddt, int64_t ldd, int accumulate,
                 T* __restrict__ partial, int64_t k_chunk, int flags) {
  pdl_enter();
  __shared__ T As[BK][BM + PAD];
  __shared__ T Bs[BK][BN + PAD];
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int64_t m0 = (int64_t)blockIdx.y * BM, n0 = (int64_t)blockIdx.x * BN;
  // structure flags (M == N): symmetric result -> only tiles on or below the diagonal are
  // computed and mirrored; lower-triangular operands -> the k range that is all zeros is skipped
  if ((flags & (XMCA_GEMM_SYMMETRIC | XMCA_GEMM_LOWER_ONLY)) && n0 > m0) return;
  int64_t kb = (int64_t)blockIdx.z * k_chunk;
  const int64_t ke = min(K, kb + k_chunk);
  if (flags & XMCA_GEMM_A_LOWER_T) kb = max(kb, m0 / BK * BK);      // opA = L^T: L[k][m] = 0 for k < m
  if (flags & XMCA_GEMM_B_LOWER) kb = max(kb, n0 / BK * BK);        // opB = L  : L[k][n] = 0 for k < n

  T acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = T(0);

  T ra[8], rb[8];
  if (kb < ke) {
    load_slab<T, AK>(ra, A, adt, lda, m0, M, kb, ke, tid);
    load_slab<T, BKM>(rb, B, bdt, ldb, n0, N, kb, ke, tid);
  }
  for (int64_t k0 = kb; k0 < ke; k0 += BK) {
    store_slab<T, AK>(ra, As, tid);
    store_slab<T, BKM>(rb, Bs, tid);
    __syncthreads();
    if (k0 + BK < ke) {
      load_slab<T, AK>(ra, A, adt, lda, m0, M, k0 + BK, ke, tid);
      load_slab<T, BKM>(rb, B, bdt, ldb, n0, N, k0 + BK, ke, tid);
    }
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      T a[8], b[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {   // interleaved micro-tile: conflict-free LDS, coalesced stores
        a[i] = As[kk][ty + 16 * i];
        b[i] = Bs[kk][tx + 16 * i];
      }
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fma(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }

#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int64_t m = m0 + ty + 16 * i;
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int64_t n = n0 + tx + 16 * j;
      if (n >= N) continue;
      if (partial) {
        partial[((int64_t)blockIdx.z * M + m) * N + n] = acc[i][j];
      } else {
        double v = alpha * (double)acc[i][j];
        if (accumulate) v += load_as_double(D, ddt, m * ldd + n);
        store_from_double(D, ddt, m * ldd + n, v);
        if ((flags & XMCA_GEMM_SYMMETRIC) && n0 < m0) store_from_double(D, ddt, n * ldd + m, v);
      }
    }
  }
}

// ------------------------------------------------------------------ fp64 on the DMMA pipe
// Same tiling contract (128 x 128 x 16 block tile, register prefetch of the next k-slab, split-K
// partials, structure flags) with the inner product on mma.sync.m8n8k4.f64: 16 warps, each a
// 32 x 32 warp tile = 4 x 4 MMA tiles (32 accumulator doubles per thread); per k-step of 4 a thread
// loads 4 + 4 fragment doubles for 16 MMAs, a quarter of the shared-memory traffic per flop of the
// 8 x 8 scalar micro-tile above, which is what kept that kernel at ~45 % of the DFMA peak.
// Operand slabs are stored [k][mn] with a row pitch of 132 doubles: the fragment loads (k = lane % 4,
// mn = lane / 4) then hit 16 distinct 8-byte banks per half-warp.
constexpr int DT = 512, DLD = BM + 4;

template <bool KMAJOR>
__device__ __forceinline__ void dload_slab(double (&r)[4], const void* p, int dt, int64_t ld,
                                           int64_t mn0, int64_t mn_end, int64_t k0, int64_t k_end, int tid) {
  if (KMAJOR) {
    const int kk = tid & 15, m = tid >> 4;        // 16 consecutive k per row
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int64_t mm = mn0 + m + 32 * i, k = k0 + kk;
      r[i] = (mm < mn_end && k < k_end) ? ld_elem<double>(p, dt, mm * ld + k) : 0.0;
    }
  } else {
    const int m = tid & 127, kk = tid >> 7;       // 128 consecutive mn per k row
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int64_t mm = mn0 + m, k = k0 + kk + 4 * i;
      r[i] = (mm < mn_end && k < k_end) ? ld_elem<double>(p, dt, k * ld + mm) : 0.0;
    }
  }
}

// Column index XOR-swizzled by (k / 4) & 3: the K-major store (16 lanes = 16 consecutive k of one row)
// then hits 16 distinct 8-byte banks instead of 4, and the fragment loads see a permutation that is
// uniform per k-step, so they stay conflict free.
template <bool KMAJOR>
__device__ __forceinline__ void dstore_slab(const double (&r)[4], double (*s)[DLD], int tid) {
  if (KMAJOR) {
    const int kk = tid & 15, m = tid >> 4;
#pragma unroll
    for (int i = 0; i < 4; ++i) s[kk][(m + 32 * i) ^ ((kk >> 2) & 3)] = r[i];
  } else {
    const int m = tid & 127, kk = tid >> 7;
#pragma unroll
    for (int i = 0; i < 4; ++i) s[kk + 4 * i][m ^ (((kk + 4 * i) >> 2) & 3)] = r[i];
  }
}

template <bool AK, bool BKM>
__global__ void __launch_bounds__(DT)
gemm_dmma_kernel(int64_t M, int64_t N, int64_t K, double alpha,
                 const void* __restrict__ A, int adt, int64_t lda,
                 const void* __restrict__ B, int bdt, int64_t ldb,
                 void* __restrict__ D, int ddt, int64_t ldd, int accumulate,
                 double* __restrict__ partial, int64_t k_chunk, int flags) {
  pdl_enter();
  __shared__ double As[BK][DLD];
  __shared__ double Bs[BK][DLD];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int gid = lane >> 2, tig = lane & 3;
  const int wm = (warp & 3) * 32, wn = (warp >> 2) * 32;
  const int64_t m0 = (int64_t)blockIdx.y * BM, n0 = (int64_t)blockIdx.x * BN;
  if ((flags & (XMCA_GEMM_SYMMETRIC | XMCA_GEMM_LOWER_ONLY)) && n0 > m0) return;
  int64_t kb = (int64_t)blockIdx.z * k_chunk;
  const int64_t ke = min(K, kb + k_chunk);
  if (flags & XMCA_GEMM_A_LOWER_T) kb = max(kb, m0 / BK * BK);
  if (flags & XMCA_GEMM_B_LOWER) kb = max(kb, n0 / BK * BK);

  double c[4][4][2];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) { c[i][j][0] = 0.0; c[i][j][1] = 0.0; }

  double ra[4], rb[4];
  if (kb < ke) {
    dload_slab<AK>(ra, A, adt, lda, m0, M, kb, ke, tid);
    dload_slab<BKM>(rb, B, bdt, ldb, n0, N, kb, ke, tid);
  }
  for (int64_t k0 = kb; k0 < ke; k0 += BK) {
    dstore_slab<AK>(ra, As, tid);
    dstore_slab<BKM>(rb, Bs, tid);
    __syncthreads();
    if (k0 + BK < ke) {
      dload_slab<AK>(ra, A, adt, lda, m0, M, k0 + BK, ke, tid);
      dload_slab<BKM>(rb, B, bdt, ldb, n0, N, k0 + BK, ke, tid);
    }
#pragma unroll
    for (int ks = 0; ks < BK; ks += 4) {
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

  // accumulate: every old value is loaded before the first store (a load after a store to the same array cannot be
  // hoisted by the compiler, and 32 dependent DRAM round trips per thread used to cost more than the tile's MMAs)
  if (!partial && accumulate) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int64_t m = m0 + wm + 8 * i + gid;
#pragma unroll
      for (int j = 0; j < 4; ++j)
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const int64_t n = n0 + wn + 8 * j + 2 * tig + e;
          const double old = (m < M && n < N) ? load_as_double(D, ddt, m * ldd + n) : 0.0;
          c[i][j][e] = fma(alpha, c[i][j][e], old);
        }
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int64_t m = m0 + wm + 8 * i + gid;
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int64_t n = n0 + wn + 8 * j + 2 * tig + e;
        if (n >= N) continue;
        if (partial) {
          partial[((int64_t)blockIdx.z * M + m) * N + n] = c[i][j][e];
        } else {
          const double v = accumulate ? c[i][j][e] : alpha * c[i][j][e];
          store_from_double(D, ddt, m * ldd + n, v);
          if ((flags & XMCA_GEMM_SYMMETRIC) && n0 < m0) store_from_double(D, ddt, n * ldd + m, v);
        }
      }
    }
  }
}

// ------------------------------------------------------------------ fp64 DMMA, skinny N (projections onto a few modes)
// N <= 64 (array.py:640 / :667: fields times 20-50 vectors): on the 128 x 128 tile half of the warps compute columns that do
// not exist.  Same inner product with a 256 x 64 block tile: 16 warps as 8 x 2, warp tile 32 x 32, operand slabs [k][mn]
// with pitches 260 / 68 (the same XOR swizzle).  No structure flags.
constexpr int SK_BM = 256, SK_BN = 64, SK_LDA = SK_BM + 4, SK_LDB = SK_BN + 4;

template <bool KMAJOR>
__device__ __forceinline__ void sk_load_a(double (&r)[8], const void* p, int dt, int64_t ld,
                                          int64_t m0, int64_t M, int64_t k0, int64_t k_end, int tid) {
  if (KMAJOR) {
    const int kk = tid & 15, m = tid >> 4;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int64_t mm = m0 + m + 32 * i, k = k0 + kk;
      r[i] = (mm < M && k < k_end) ? ld_elem<double>(p, dt, mm * ld + k) : 0.0;
    }
  } else {
    const int m = tid & 255, kk = tid >> 8;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int64_t mm = m0 + m, k = k0 + kk + 2 * i;
      r[i] = (mm < M && k < k_end) ? ld_elem<double>(p, dt, k * ld + mm) : 0.0;
    }
  }
}
template <bool KMAJOR>
__device__ __forceinline__ void sk_store_a(const double (&r)[8], double (*s)[SK_LDA], int tid) {
  if (KMAJOR) {
    const int kk = tid & 15, m = tid >> 4;
#pragma unroll
    for (int i = 0; i < 8; ++i) s[kk][(m + 32 * i) ^ ((kk >> 2) & 3)] = r[i];
  } else {
    const int m = tid & 255, kk = tid >> 8;
#pragma unroll
    for (int i = 0; i < 8; ++i) s[kk + 2 * i][m ^ (((kk + 2 * i) >> 2) & 3)] = r[i];
  }
}
template <bool KMAJOR>
__device__ __forceinline__ void sk_load_b(double (&r)[2], const void* p, int dt, int64_t ld,
                                          int64_t n0, int64_t N, int64_t k0, int64_t k_end, int tid) {
  if (KMAJOR) {
    const int kk = tid & 15, n = tid >> 4;
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int64_t nn = n0 + n + 32 * i, k = k0 + kk;
      r[i] = (nn < N && k < k_end) ? ld_elem<double>(p, dt, nn * ld + k) : 0.0;
    }
  } else {
    const int n = tid & 63, kk = tid >> 6;
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int64_t nn = n0 + n, k = k0 + kk + 8 * i;
      r[i] = (nn < N && k < k_end) ? ld_elem<double>(p, dt, k * ld + nn) : 0.0;
    }
  }
}
template <bool KMAJOR>
__device__ __forceinline__ void sk_store_b(const double (&r)[2], double (*s)[SK_LDB], int tid) {
  if (KMAJOR) {
    const int kk = tid & 15, n = tid >> 4;
#pragma unroll
    for (int i = 0; i < 2; ++i) s[kk][(n + 32 * i) ^ ((kk >> 2) & 3)] = r[i];
  } else {
    const int n = tid & 63, kk = tid >> 6;
#pragma unroll
    for (int i = 0; i < 2; ++i) s[kk + 8 * i][n ^ (((kk + 8 * i) >> 2) & 3)] = r[i];
  }
}

template <bool AK, bool BKM>
__global__ void __launch_bounds__(DT)
gemm_dmma_skinny_kernel(int64_t M, int64_t N, int64_t K, double alpha,
                        const void* __restrict__ A, int adt, int64_t lda,
                        const void* __restrict__ B, int bdt, int64_t ldb,
                        void* __restrict__ D, int ddt, int64_t ldd, int accumulate,
                        double* __restrict__ partial, int64_t k_chunk) {
  pdl_enter();
  __shared__ double As[BK][SK_LDA];
  __shared__ double Bs[BK][SK_LDB];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int gid = lane >> 2, tig = lane & 3;
  const int wm = (warp & 7) * 32, wn = (warp >> 3) * 32;
  const int64_t m0 = (int64_t)blockIdx.y * SK_BM, n0 = (int64_t)blockIdx.x * SK_BN;
  const int64_t kb = (int64_t)blockIdx.z * k_chunk;
  const int64_t ke = min(K, kb + k_chunk);

  double c[4][4][2];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) { c[i][j][0] = 0.0; c[i][j][1] = 0.0; }
  double ra[8], rb[2];
  if (kb < ke) {
    sk_load_a<AK>(ra, A, adt, lda, m0, M, kb, ke, tid);
    sk_load_b<BKM>(rb, B, bdt, ldb, n0, N, kb, ke, tid);
  }
  for (int64_t k0 = kb; k0 < ke; k0 += BK) {
    sk_store_a<AK>(ra, As, tid);
    sk_store_b<BKM>(rb, Bs, tid);
    __syncthreads();
    if (k0 + BK < ke) {
      sk_load_a<AK>(ra, A, adt, lda, m0, M, k0 + BK, ke, tid);
      sk_load_b<BKM>(rb, B, bdt, ldb, n0, N, k0 + BK, ke, tid);
    }
#pragma unroll
    for (int ks = 0; ks < BK; ks += 4) {
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
  if (!partial && accumulate) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int64_t m = m0 + wm + 8 * i + gid;
#pragma unroll
      for (int j = 0; j < 4; ++j)
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const int64_t n = n0 + wn + 8 * j + 2 * tig + e;
          const double old = (m < M && n < N) ? load_as_double(D, ddt, m * ldd + n) : 0.0;
          c[i][j][e] = fma(alpha, c[i][j][e], old);
        }
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int64_t m = m0 + wm + 8 * i + gid;
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int64_t n = n0 + wn + 8 * j + 2 * tig + e;
        if (n >= N) continue;
        if (partial) partial[((int64_t)blockIdx.z * M + m) * N + n] = c[i][j][e];
        else store_from_double(D, ddt, m * ldd + n, accumulate ? c[i][j][e] : alpha * c[i][j][e]);
      }
  }
}

template <typename T>
__global__ void splitk_reduce_kernel(int64_t M, int64_t N, int split, double alpha,
                                     const T* __restrict__ partial,
                                     void* __restrict__ D, int ddt, int64_t ldd, int accumulate) {
  pdl_enter();
  int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= M * N) return;
  double s = 0.0;
  for (int z = 0; z < split; ++z) s += (double)partial[(int64_t)z * M * N + idx];
  int64_t m = idx / N, n = idx % N;
  double v = alpha * s;
  if (accumulate) v += load_as_double(D, ddt, m * ldd + n);
  store_from_double(D, ddt, m * ldd + n, v);
}

template <typename T>
static int launch_gemm(int ak, int bk, int64_t M, int64_t N, int64_t K, double alpha,
                       const void* A, int adt, int64_t lda, const void* B, int bdt, int64_t ldb,
                       void* D, int ddt, int64_t ldd, int accumulate, int split,
                       void* ws, int flags, cudaStream_t st) {
  int64_t k_chunk = ((K + split - 1) / split + BK - 1) / BK * BK;
  if (k_chunk < BK) k_chunk = BK;
  dim3 grid((unsigned)((N + BN - 1) / BN), (unsigned)((M + BM - 1) / BM), (unsigned)split);
  T* partial = split > 1 ? reinterpret_cast<T*>(ws) : nullptr;
  if (sizeof(T) == 8 && N <= SK_BN && M >= 2 * SK_BM && K >= 1024 && flags == 0) {     // skinny N, long K: 256 x 64 tiles
    double* dpart = reinterpret_cast<double*>(partial);
    dim3 sgrid(1, (unsigned)((M + SK_BM - 1) / SK_BM), (unsigned)split);
#define GOS(AKF, BKF)                                                                            \
  launch_pdl(true, gemm_dmma_skinny_kernel<AKF, BKF>, sgrid, dim3(DT), 0, st, M, N, K, alpha, A, adt, lda, B, bdt, ldb, \
             D, ddt, ldd, accumulate, dpart, k_chunk)
    if (ak && bk) GOS(true, true);
    else if (ak && !bk) GOS(true, false);
    else if (!ak && bk) GOS(false, true);
    else GOS(false, false);
#undef GOS
    XMCA_LAUNCHED();
    if (split > 1) {
      int64_t tot = M * N;
      launch_pdl(true, splitk_reduce_kernel<T>, dim3((unsigned)((tot + 255) / 256)), dim3(256), 0, st,
                 M, N, split, alpha, (const T*)partial, D, ddt, ldd, accumulate);
      XMCA_LAUNCHED();
    }
    return XMCA_OK;
  }
  if (sizeof(T) == 8) {                 // fp64 accumulation: DMMA kernel
    double* dpart = reinterpret_cast<double*>(partial);
#define GOD(AKF, BKF)                                                                    \
  launch_pdl(true, gemm_dmma_kernel<AKF, BKF>, grid, dim3(DT), 0, st, M, N, K, alpha, A, adt, lda, B, bdt, ldb, \
             D, ddt, ldd, accumulate, dpart, k_chunk, flags)
    if (ak && bk) GOD(true, true);
    else if (ak && !bk) GOD(true, false);
    else if (!ak && bk) GOD(false, true);
    else GOD(false, false);
#undef GOD
    XMCA_LAUNCHED();
    if (split > 1) {
      int64_t tot = M * N;
      launch_pdl(true, splitk_reduce_kernel<T>, dim3((unsigned)((tot + 255) / 256)), dim3(256), 0, st,
                 M, N, split, alpha, (const T*)partial, D, ddt, ldd, accumulate);
      XMCA_LAUNCHED();
    }
    return XMCA_OK;
  }
#define GO(AKF, BKF)                                                                     \
  launch_pdl(true, gemm_simt_kernel<T, AKF, BKF>, grid, dim3(NT), 0, st, M, N, K, alpha, A, adt, lda, B, bdt, \
             ldb, D, ddt, ldd, accumulate, partial, k_chunk, flags)
  if (ak && bk) GO(true, true);
  else if (ak && !bk) GO(true, false);
  else if (!ak && bk) GO(false, true);
  else GO(false, false);
#undef GO
  XMCA_LAUNCHED();
  if (split > 1) {
    int64_t tot = M * N;
    splitk_reduce_kernel<T><<<(unsigned)((tot + 255) / 256), 256, 0, st>>>(
        M, N, split, alpha, partial, D, ddt, ldd, accumulate);
    XMCA_LAUNCHED();
  }
  return XMCA_OK;
}

}  // namespace xmca

using namespace xmca;

extern "C" size_t xmca_gemm_workspace_bytes(int64_t M, int64_t N, int split_k, int acc_dtype) {
  if (split_k <= 1) return 0;
  return (size_t)M * (size_t)N * (size_t)split_k * (size_t)dtype_size(acc_dtype);
}

extern "C" int xmca_gemm(int a_kmajor, int b_kmajor, int64_t M, int64_t N, int64_t K, double alpha,
                         const void* d_A, int a_dtype, int64_t lda,
                         const void* d_B, int b_dtype, int64_t ldb,
                         void* d_D, int d_dtype, int64_t ldd, int accumulate,
                         int acc_dtype, int split_k, void* d_workspace, size_t workspace_bytes,
                         void* stream) {
  return xmca_gemm_ex(a_kmajor, b_kmajor, M, N, K, alpha, d_A, a_dtype, lda, d_B, b_dtype, ldb, d_D, d_dtype, ldd,
                      accumulate, acc_dtype, split_k, d_workspace, workspace_bytes, 0, stream);
}

extern "C" int xmca_gemm_ex(int a_kmajor, int b_kmajor, int64_t M, int64_t N, int64_t K, double alpha,
                            const void* d_A, int a_dtype, int64_t lda,
                            const void* d_B, int b_dtype, int64_t ldb,
                            void* d_D, int d_dtype, int64_t ldd, int accumulate,
                            int acc_dtype, int split_k, void* d_workspace, size_t workspace_bytes,
                            int flags, void* stream) {
  if (flags) {
    XMCA_REQUIRE(M == N, "xmca_gemm_ex: structure flags need a square result");
    XMCA_REQUIRE(split_k <= 1, "xmca_gemm_ex: structure flags cannot be combined with split_k");
    XMCA_REQUIRE(!(flags & XMCA_GEMM_A_LOWER_T) || (!a_kmajor && K == M), "xmca_gemm_ex: A_LOWER_T needs opA = L^T (K x M storage)");
    XMCA_REQUIRE(!(flags & XMCA_GEMM_B_LOWER) || (!b_kmajor && K == N), "xmca_gemm_ex: B_LOWER needs opB = L (K x N storage)");
  }
  XMCA_REQUIRE(M > 0 && N > 0 && K > 0, "xmca_gemm: empty problem");
  XMCA_REQUIRE(d_A && d_B && d_D, "xmca_gemm: null operand");
  XMCA_REQUIRE(dtype_ok(a_dtype) && dtype_ok(b_dtype) && dtype_ok(d_dtype) && dtype_ok(acc_dtype),
               "xmca_gemm: bad dtype");
  XMCA_REQUIRE(lda >= (a_kmajor ? K : M) && ldb >= (b_kmajor ? K : N) && ldd >= N,
               "xmca_gemm: leading dimension too small");
  if (split_k < 1) split_k = 1;
  XMCA_REQUIRE(split_k <= 65535, "xmca_gemm: split_k too large");
  XMCA_REQUIRE((M + BM - 1) / BM <= 65535, "xmca_gemm: M too large for grid.y");
  if (split_k > 1)
    XMCA_REQUIRE(d_workspace && workspace_bytes >= xmca_gemm_workspace_bytes(M, N, split_k, acc_dtype),
                 "xmca_gemm: workspace too small for split_k");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (acc_dtype == XMCA_F64)
    return launch_gemm<double>(a_kmajor, b_kmajor, M, N, K, alpha, d_A, a_dtype, lda, d_B, b_dtype,
                               ldb, d_D, d_dtype, ldd, accumulate, split_k, d_workspace, flags, st);
  return launch_gemm<float>(a_kmajor, b_kmajor, M, N, K, alpha, d_A, a_dtype, lda, d_B, b_dtype, ldb,
                            d_D, d_dtype, ldd, accumulate, split_k, d_workspace, flags, st);
}
