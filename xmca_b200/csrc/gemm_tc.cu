// Cross-covariance GEMM on the 5th-generation tensor cores:
//     D[M,N] = alpha * sum_k A[m,k] * B[n,k]        (both operands K-major)
// with fp32 operands fed as THREE TF32 products (hi*hi + hi*lo + lo*hi) so the
// result carries fp32-level accuracy (error ~2^-22 per product) while running
// on tcgen05.mma.kind::tf32 with fp32 accumulators in tensor memory.
//
// This is the engine's replacement for the reference's route to the
// covariance "kernel" (array.py:552-566): instead of SVD-ing each field and
// multiplying the reduced factors, C = A^T B / (T-1) is formed directly.
//
// Structure (one 128 x BN output tile per CTA, 320 threads):
//   warp 0      TMA producer : cp.async.bulk.tensor.2d of the 4 operand slabs
//                              (A_hi, A_lo, B_hi, B_lo; 32 fp32 = 128 B rows,
//                              SWIZZLE_128B) into a ring of shared-memory stages
//   warp 1      MMA issuer   : one elected thread issues 12 tcgen05.mma per
//                              stage (4 k-steps of 8 x 3 products), commits the
//                              stage back to the producer with tcgen05.commit
//   warps 2..9  accumulate + : the tensor core TRUNCATES the fp32 accumulator
//               epilogue       on every MMA (measured: error grows linearly
//                              with the number of chained MMAs, ~2^-24 |acc|
//                              each), so the K loop is cut into chunks of
//                              TC_CHUNK_KB stages.  Each chunk accumulates into
//                              one of two TMEM buffers; these warps drain the
//                              finished buffer with tcgen05.ld and add it into
//                              fp32 REGISTER sums (round-to-nearest) while the
//                              tensor core fills the other buffer.  At the end:
//                              scale, optional sum(D^2), store to HBM.
// Tiles are rasterised in groups of 16 tile-rows so that the ~148 concurrently
// resident CTAs share operand slabs through the 126 MB L2.
#include "common.cuh"
#include <cuda.h>
#include <mutex>

namespace xmca {

constexpr int TC_BM = 128;
constexpr int TC_BK = 32;                 // fp32 elements per stage row = 128 bytes (one swizzle atom)
constexpr int TC_UMMA_K = 8;              // tf32: 32 bytes per MMA k-step
constexpr int TC_THREADS = 320;           // TMA warp + MMA warp + 8 accumulate/epilogue warps
constexpr int TC_EPI_WARPS = 8;
constexpr int TC_CHUNK_KB = 4;            // stages (of 32 k) chained in TMEM before a register flush: 48 MMAs
constexpr int TC_GROUP_M = 16;

// ------------------------------------------------------------------ PTX glue
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred P1;\n"
      "WAIT_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
      "@P1 bra DONE;\n"
      "bra WAIT_LOOP;\n"
      "DONE:\n"
      "}\n" ::"r"(smem_u32(bar)), "r"(parity)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* map, int c0, int c1,
                                            uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
      ::"r"(smem_u32(smem_dst)), "l"(map), "r"(c0), "r"(c1), "r"(smem_u32(bar))
      : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
        "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]),
        "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]),
        "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]),
        "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
        "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// K-major, SWIZZLE_128B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor):
//   [0,14) start address >> 4 | [16,30) LBO >> 4 (unused for swizzled K-major, 1)
//   [32,46) SBO >> 4 = 1024 B (8 rows x 128 B) | [46,48) version = 1 | [61,64) layout = 2 (SW128)
__device__ __forceinline__ uint64_t make_kmajor_sw128_desc(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}

template <int BN>
struct TcCfg {
  static constexpr int kStages = (BN == 256) ? 2 : 3;
  static constexpr int kABytes = TC_BM * TC_BK * 4;          // 16 KB per plane
  static constexpr int kBBytes = BN * TC_BK * 4;
  static constexpr int kStageBytes = 2 * kABytes + 2 * kBBytes;
  static constexpr int kSmemBytes = kStages * kStageBytes + 1024 /*align*/ + 256 /*barriers*/;
  // instruction descriptor (cute::UMMA::InstrDescriptor): c=F32, a=b=TF32, K-major both, N, M
  static constexpr uint32_t kIdesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(BN >> 3) << 17) |
                                     ((uint32_t)(TC_BM >> 4) << 24);
};

template <int BN>
__global__ void __launch_bounds__(TC_THREADS, 1)
tc_gemm_nt_kernel(const __grid_constant__ CUtensorMap mapAhi, const __grid_constant__ CUtensorMap mapAlo,
                  const __grid_constant__ CUtensorMap mapBhi, const __grid_constant__ CUtensorMap mapBlo,
                  int M, int N, int K, float alpha, float* __restrict__ D, int64_t ldd,
                  double* __restrict__ frob2, int tiles_m, int tiles_n) {
  using Cfg = TcCfg<BN>;
  constexpr int kCols = BN / 2;             // accumulator columns per epilogue thread (two column halves)
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + Cfg::kStages * Cfg::kStageBytes);
  uint64_t* empty_bar = full_bar + Cfg::kStages;
  uint64_t* tmem_full_bar = empty_bar + Cfg::kStages;      // [2]
  uint64_t* tmem_empty_bar = tmem_full_bar + 2;            // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty_bar + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  // grouped rasterisation: TC_GROUP_M tile-rows per group, column-major inside a group
  int tile_m, tile_n;
  {
    const int id = blockIdx.x;
    const int per_group = TC_GROUP_M * tiles_n;
    const int g = id / per_group;
    const int first_m = g * TC_GROUP_M;
    const int gm = min(TC_GROUP_M, tiles_m - first_m);
    const int in = id - g * per_group;
    tile_m = first_m + in % gm;
    tile_n = in / gm;
  }
  const int m0 = tile_m * TC_BM, n0 = tile_n * BN;
  const int num_kb = (K + TC_BK - 1) / TC_BK;
  const int num_chunks = (num_kb + TC_CHUNK_KB - 1) / TC_CHUNK_KB;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&mapAhi) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&mapAlo) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&mapBhi) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&mapBlo) : "memory");
    for (int s = 0; s < Cfg::kStages; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    for (int b = 0; b < 2; ++b) { mbar_init(&tmem_full_bar[b], 1); mbar_init(&tmem_empty_bar[b], TC_EPI_WARPS); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"((uint32_t)(2 * BN))
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      for (int kb = 0; kb < num_kb; ++kb) {
        mbar_wait(&empty_bar[stage], phase ^ 1);
        uint8_t* st = smem + stage * Cfg::kStageBytes;
        mbar_expect_tx(&full_bar[stage], Cfg::kStageBytes);
        const int k0 = kb * TC_BK;
        tma_load_2d(st, &mapAhi, k0, m0, &full_bar[stage]);
        tma_load_2d(st + Cfg::kABytes, &mapAlo, k0, m0, &full_bar[stage]);
        tma_load_2d(st + 2 * Cfg::kABytes, &mapBhi, k0, n0, &full_bar[stage]);
        tma_load_2d(st + 2 * Cfg::kABytes + Cfg::kBBytes, &mapBlo, k0, n0, &full_bar[stage]);
        if (++stage == Cfg::kStages) { stage = 0; phase ^= 1; }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      int kb = 0;
      for (int c = 0; c < num_chunks; ++c) {
        const int buf = c & 1;
        mbar_wait(&tmem_empty_bar[buf], ((uint32_t)(c >> 1) & 1u) ^ 1u);   // drained by the epilogue warps
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + (uint32_t)(buf * BN);
        const int kb_end = min(num_kb, kb + TC_CHUNK_KB);
        bool first = true;
        for (; kb < kb_end; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + stage * Cfg::kStageBytes);
          const uint32_t a_hi = sa, a_lo = sa + Cfg::kABytes;
          const uint32_t b_hi = sa + 2 * Cfg::kABytes, b_lo = b_hi + Cfg::kBBytes;
#pragma unroll
          for (int k = 0; k < TC_BK / TC_UMMA_K; ++k) {
            const uint32_t off = k * TC_UMMA_K * 4;       // 32 bytes along K inside the swizzle atom
            const uint64_t dah = make_kmajor_sw128_desc(a_hi + off), dal = make_kmajor_sw128_desc(a_lo + off);
            const uint64_t dbh = make_kmajor_sw128_desc(b_hi + off), dbl = make_kmajor_sw128_desc(b_lo + off);
            umma_tf32(tmem_d, dal, dbh, Cfg::kIdesc, first ? 0u : 1u);   // small terms first
            umma_tf32(tmem_d, dah, dbl, Cfg::kIdesc, 1u);
            umma_tf32(tmem_d, dah, dbh, Cfg::kIdesc, 1u);
            first = false;
          }
          umma_commit(&empty_bar[stage]);                  // frees the smem stage when the MMAs retire
          if (++stage == Cfg::kStages) { stage = 0; phase ^= 1; }
        }
        umma_commit(&tmem_full_bar[buf]);                  // this chunk's partial sum is complete
      }
    }
    __syncwarp();
  } else {
    // accumulate/epilogue warps 2..9 -> TMEM lane quadrant (warp % 4), column half (warp - 2) / 4
    const int q = warp & 3, half = (warp - 2) >> 2;
    float acc[kCols];
#pragma unroll
    for (int j = 0; j < kCols; ++j) acc[j] = 0.f;
    for (int c = 0; c < num_chunks; ++c) {
      const int buf = c & 1;
      mbar_wait(&tmem_full_bar[buf], (uint32_t)(c >> 1) & 1u);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(buf * BN + half * kCols);
#pragma unroll
      for (int g = 0; g < kCols / 32; ++g) {
        uint32_t v[32];
        tmem_ld32(taddr + (uint32_t)(g * 32), v);
#pragma unroll
        for (int j = 0; j < 32; ++j) acc[g * 32 + j] += __uint_as_float(v[j]);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty_bar[buf]);
    }
    const int row = m0 + q * 32 + lane;
    double ss = 0.0;
    const bool vec_ok = ((ldd & 3) == 0) && ((reinterpret_cast<uintptr_t>(D) & 15) == 0);
    const int col0 = n0 + half * kCols;
    if (row < M && col0 < N) {
      float* dst = D + (int64_t)row * ldd + col0;
      if (vec_ok && col0 + kCols <= N) {
#pragma unroll
        for (int j = 0; j < kCols; j += 4) {
          float4 o;
          o.x = acc[j] * alpha; o.y = acc[j + 1] * alpha; o.z = acc[j + 2] * alpha; o.w = acc[j + 3] * alpha;
          ss += (double)o.x * o.x + (double)o.y * o.y + (double)o.z * o.z + (double)o.w * o.w;
          *reinterpret_cast<float4*>(dst + j) = o;
        }
      } else {
#pragma unroll
        for (int j = 0; j < kCols; ++j) {
          if (col0 + j < N) {
            float o = acc[j] * alpha;
            ss += (double)o * o;
            dst[j] = o;
          }
        }
      }
    }
    if (frob2) {
      ss = warp_sum(ss);
      if (lane == 0 && ss != 0.0) atomicAdd(frob2, ss);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)(2 * BN))
                 : "memory");
  }
}

// ------------------------------------------------------------------ fp64-output variant (Gram matrices)
// D[M,N] (fp64) = alpha * sum_k A[m,k] B[n,k], same 3xTF32 operand split, built for the long
// contractions of the Gram matrices G = X X^T (K = S up to 65536, all diagonal terms positive).
// The tensor core truncates the fp32 accumulator after every MMA, a bias that grows with the
// length of the chain and with |acc|, so here
//   * the large products (hi*hi) and the small ones (hi*lo, lo*hi) go to SEPARATE TMEM
//     accumulators: the main chain is 8 MMAs per chunk of 2 stages (64 k), the small chain
//     carries values 2^-11 smaller whose truncation is negligible,
//   * every chunk is drained to registers, main + small added in fp32 (one rounding) and
//     accumulated in FP64 registers (64 per epilogue thread): the sum over K/64 chunks adds
//     no further error.  Measured against the fp64 product: ~1e-7 relative on the diagonal.
// symmetric != 0: tiles above the diagonal are skipped and mirrored from below.
constexpr int TG_BN = 128;
constexpr int TG_CHUNK_KB = 2;

__global__ void __launch_bounds__(TC_THREADS, 1)
tc_gemm_nt_f64_kernel(const __grid_constant__ CUtensorMap mapAhi, const __grid_constant__ CUtensorMap mapAlo,
                      const __grid_constant__ CUtensorMap mapBhi, const __grid_constant__ CUtensorMap mapBlo,
                      int M, int N, int K, double alpha, double* __restrict__ D, int64_t ldd,
                      int symmetric, int tiles_m, int tiles_n) {
  using Cfg = TcCfg<TG_BN>;
  constexpr int kCols = TG_BN / 2;          // 64 accumulator columns per epilogue thread
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + Cfg::kStages * Cfg::kStageBytes);
  uint64_t* empty_bar = full_bar + Cfg::kStages;
  uint64_t* tmem_full_bar = empty_bar + Cfg::kStages;      // [2]
  uint64_t* tmem_empty_bar = tmem_full_bar + 2;            // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty_bar + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  int tile_m, tile_n;
  {
    const int id = blockIdx.x;
    const int per_group = TC_GROUP_M * tiles_n;
    const int g = id / per_group;
    const int first_m = g * TC_GROUP_M;
    const int gm = min(TC_GROUP_M, tiles_m - first_m);
    const int in = id - g * per_group;
    tile_m = first_m + in % gm;
    tile_n = in / gm;
  }
  const int m0 = tile_m * TC_BM, n0 = tile_n * TG_BN;
  if (symmetric && n0 > m0) return;                         // uniform: mirrored from the tile below the diagonal
  const int num_kb = (K + TC_BK - 1) / TC_BK;
  const int num_chunks = (num_kb + TG_CHUNK_KB - 1) / TG_CHUNK_KB;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&mapAhi) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&mapAlo) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&mapBhi) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&mapBlo) : "memory");
    for (int s = 0; s < Cfg::kStages; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    for (int b = 0; b < 2; ++b) { mbar_init(&tmem_full_bar[b], 1); mbar_init(&tmem_empty_bar[b], TC_EPI_WARPS); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    // 2 buffers x (main 128 + small 128) columns = the whole tensor memory of the SM
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"(512u)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      for (int kb = 0; kb < num_kb; ++kb) {
        mbar_wait(&empty_bar[stage], phase ^ 1);
        uint8_t* st = smem + stage * Cfg::kStageBytes;
        mbar_expect_tx(&full_bar[stage], Cfg::kStageBytes);
        const int k0 = kb * TC_BK;
        tma_load_2d(st, &mapAhi, k0, m0, &full_bar[stage]);
        tma_load_2d(st + Cfg::kABytes, &mapAlo, k0, m0, &full_bar[stage]);
        tma_load_2d(st + 2 * Cfg::kABytes, &mapBhi, k0, n0, &full_bar[stage]);
        tma_load_2d(st + 2 * Cfg::kABytes + Cfg::kBBytes, &mapBlo, k0, n0, &full_bar[stage]);
        if (++stage == Cfg::kStages) { stage = 0; phase ^= 1; }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      int kb = 0;
      for (int c = 0; c < num_chunks; ++c) {
        const int buf = c & 1;
        mbar_wait(&tmem_empty_bar[buf], ((uint32_t)(c >> 1) & 1u) ^ 1u);
        tc_fence_after();
        const uint32_t tmem_main = tmem_base + (uint32_t)(buf * 2 * TG_BN);
        const uint32_t tmem_small = tmem_main + (uint32_t)TG_BN;
        const int kb_end = min(num_kb, kb + TG_CHUNK_KB);
        bool first = true;
        for (; kb < kb_end; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + stage * Cfg::kStageBytes);
          const uint32_t a_hi = sa, a_lo = sa + Cfg::kABytes;
          const uint32_t b_hi = sa + 2 * Cfg::kABytes, b_lo = b_hi + Cfg::kBBytes;
#pragma unroll
          for (int k = 0; k < TC_BK / TC_UMMA_K; ++k) {
            const uint32_t off = k * TC_UMMA_K * 4;
            const uint64_t dah = make_kmajor_sw128_desc(a_hi + off), dal = make_kmajor_sw128_desc(a_lo + off);
            const uint64_t dbh = make_kmajor_sw128_desc(b_hi + off), dbl = make_kmajor_sw128_desc(b_lo + off);
            umma_tf32(tmem_small, dal, dbh, Cfg::kIdesc, first ? 0u : 1u);
            umma_tf32(tmem_small, dah, dbl, Cfg::kIdesc, 1u);
            umma_tf32(tmem_main, dah, dbh, Cfg::kIdesc, first ? 0u : 1u);
            first = false;
          }
          umma_commit(&empty_bar[stage]);
          if (++stage == Cfg::kStages) { stage = 0; phase ^= 1; }
        }
        umma_commit(&tmem_full_bar[buf]);
      }
    }
    __syncwarp();
  } else {
    const int q = warp & 3, half = (warp - 2) >> 2;
    double acc[kCols];
#pragma unroll
    for (int j = 0; j < kCols; ++j) acc[j] = 0.0;
    for (int c = 0; c < num_chunks; ++c) {
      const int buf = c & 1;
      mbar_wait(&tmem_full_bar[buf], (uint32_t)(c >> 1) & 1u);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(buf * 2 * TG_BN + half * kCols);
#pragma unroll
      for (int g = 0; g < kCols / 16; ++g) {
        uint32_t v[16], w[16];
        tmem_ld16(taddr + (uint32_t)(g * 16), v);
        tmem_ld16(taddr + (uint32_t)(TG_BN + g * 16), w);
#pragma unroll
        for (int j = 0; j < 16; ++j) acc[g * 16 + j] += (double)(__uint_as_float(v[j]) + __uint_as_float(w[j]));
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty_bar[buf]);
    }
    const int row = m0 + q * 32 + lane;
    const int col0 = n0 + half * kCols;
    if (row < M) {
      double* dst = D + (int64_t)row * ldd + col0;
      const bool mirror = symmetric && n0 < m0;
#pragma unroll
      for (int j = 0; j < kCols; ++j) {
        if (col0 + j < N) {
          const double o = acc[j] * alpha;
          dst[j] = o;
          if (mirror) D[(int64_t)(col0 + j) * ldd + row] = o;     // lanes <-> consecutive rows: coalesced
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

// ------------------------------------------------------------ host: tensor maps
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                  CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                  CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  });
  return fn;
}

static int make_map(CUtensorMap* map, const float* base, int64_t rows, int64_t K, int64_t ld, int box_rows) {
  EncodeTiledFn enc = get_encode_fn();
  if (!enc) return fail(XMCA_CUDA_ERROR, "cuTensorMapEncodeTiled entry point not found", __FILE__, __LINE__);
  cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * 4};
  cuuint32_t box[2] = {(cuuint32_t)TC_BK, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)base, dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    char msg[128];
    snprintf(msg, sizeof msg, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
    return fail(XMCA_CUDA_ERROR, msg, __FILE__, __LINE__);
  }
  return XMCA_OK;
}

template <int BN>
static int launch_tc(int64_t M, int64_t N, int64_t K, float alpha, const float* Ahi, const float* Alo,
                     int64_t lda, const float* Bhi, const float* Blo, int64_t ldb, float* D, int64_t ldd,
                     double* frob2, cudaStream_t st) {
  using Cfg = TcCfg<BN>;
  CUtensorMap mAh, mAl, mBh, mBl;
  int rc;
  if ((rc = make_map(&mAh, Ahi, M, K, lda, TC_BM)) != XMCA_OK) return rc;
  if ((rc = make_map(&mAl, Alo, M, K, lda, TC_BM)) != XMCA_OK) return rc;
  if ((rc = make_map(&mBh, Bhi, N, K, ldb, BN)) != XMCA_OK) return rc;
  if ((rc = make_map(&mBl, Blo, N, K, ldb, BN)) != XMCA_OK) return rc;
  // (function attributes are per DEVICE: set on every call -- a process-wide one-time flag breaks a second GPU)
  XMCA_CUDA(cudaFuncSetAttribute(tc_gemm_nt_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 Cfg::kSmemBytes));
  const int tiles_m = (int)((M + TC_BM - 1) / TC_BM), tiles_n = (int)((N + BN - 1) / BN);
  tc_gemm_nt_kernel<BN><<<tiles_m * tiles_n, TC_THREADS, Cfg::kSmemBytes, st>>>(
      mAh, mAl, mBh, mBl, (int)M, (int)N, (int)K, alpha, D, ldd, frob2, tiles_m, tiles_n);
  XMCA_LAUNCHED();
  return XMCA_OK;
}

}  // namespace xmca

using namespace xmca;

extern "C" int xmca_tc_gemm_nt(int64_t M, int64_t N, int64_t K, float alpha,
                               const float* d_Ahi, const float* d_Alo, int64_t lda,
                               const float* d_Bhi, const float* d_Blo, int64_t ldb,
                               float* d_D, int64_t ldd, double* d_frob2, void* stream) {
  XMCA_REQUIRE(M > 0 && N > 0 && K > 0, "xmca_tc_gemm_nt: empty problem");
  XMCA_REQUIRE(d_Ahi && d_Alo && d_Bhi && d_Blo && d_D, "xmca_tc_gemm_nt: null operand");
  XMCA_REQUIRE(lda >= K && ldb >= K && ldd >= N, "xmca_tc_gemm_nt: leading dimension too small");
  XMCA_REQUIRE((lda & 3) == 0 && (ldb & 3) == 0, "xmca_tc_gemm_nt: operand pitch must be a multiple of 4 floats (TMA)");
  XMCA_REQUIRE(((uintptr_t)d_Ahi & 15) == 0 && ((uintptr_t)d_Alo & 15) == 0 && ((uintptr_t)d_Bhi & 15) == 0 &&
                   ((uintptr_t)d_Blo & 15) == 0,
               "xmca_tc_gemm_nt: operand planes must be 16-byte aligned (TMA)");
  XMCA_REQUIRE(M < (1LL << 31) && N < (1LL << 31) && K < (1LL << 31), "xmca_tc_gemm_nt: dimension too large");
  cudaStream_t st = (cudaStream_t)stream;
  if (N > 128)
    return launch_tc<256>(M, N, K, alpha, d_Ahi, d_Alo, lda, d_Bhi, d_Blo, ldb, d_D, ldd, d_frob2, st);
  return launch_tc<128>(M, N, K, alpha, d_Ahi, d_Alo, lda, d_Bhi, d_Blo, ldb, d_D, ldd, d_frob2, st);
}

extern "C" int xmca_tc_gemm_nt_f64(int64_t M, int64_t N, int64_t K, double alpha,
                                   const float* d_Ahi, const float* d_Alo, int64_t lda,
                                   const float* d_Bhi, const float* d_Blo, int64_t ldb,
                                   double* d_D, int64_t ldd, int symmetric, void* stream) {
  XMCA_REQUIRE(M > 0 && N > 0 && K > 0, "xmca_tc_gemm_nt_f64: empty problem");
  XMCA_REQUIRE(d_Ahi && d_Alo && d_Bhi && d_Blo && d_D, "xmca_tc_gemm_nt_f64: null operand");
  XMCA_REQUIRE(lda >= K && ldb >= K && ldd >= N, "xmca_tc_gemm_nt_f64: leading dimension too small");
  XMCA_REQUIRE((lda & 3) == 0 && (ldb & 3) == 0, "xmca_tc_gemm_nt_f64: operand pitch must be a multiple of 4 floats (TMA)");
  XMCA_REQUIRE(((uintptr_t)d_Ahi & 15) == 0 && ((uintptr_t)d_Alo & 15) == 0 && ((uintptr_t)d_Bhi & 15) == 0 &&
                   ((uintptr_t)d_Blo & 15) == 0,
               "xmca_tc_gemm_nt_f64: operand planes must be 16-byte aligned (TMA)");
  XMCA_REQUIRE(M < (1LL << 31) && N < (1LL << 31) && K < (1LL << 31), "xmca_tc_gemm_nt_f64: dimension too large");
  XMCA_REQUIRE(!symmetric || M == N, "xmca_tc_gemm_nt_f64: symmetric needs M == N");
  using Cfg = TcCfg<TG_BN>;
  cudaStream_t st = (cudaStream_t)stream;
  CUtensorMap mAh, mAl, mBh, mBl;
  int rc;
  if ((rc = make_map(&mAh, d_Ahi, M, K, lda, TC_BM)) != XMCA_OK) return rc;
  if ((rc = make_map(&mAl, d_Alo, M, K, lda, TC_BM)) != XMCA_OK) return rc;
  if ((rc = make_map(&mBh, d_Bhi, N, K, ldb, TG_BN)) != XMCA_OK) return rc;
  if ((rc = make_map(&mBl, d_Blo, N, K, ldb, TG_BN)) != XMCA_OK) return rc;
  XMCA_CUDA(cudaFuncSetAttribute(tc_gemm_nt_f64_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes));
  const int tiles_m = (int)((M + TC_BM - 1) / TC_BM), tiles_n = (int)((N + TG_BN - 1) / TG_BN);
  tc_gemm_nt_f64_kernel<<<tiles_m * tiles_n, TC_THREADS, Cfg::kSmemBytes, st>>>(
      mAh, mAl, mBh, mBl, (int)M, (int)N, (int)K, alpha, d_D, ldd, symmetric, tiles_m, tiles_n);
  XMCA_LAUNCHED();
  return XMCA_OK;
}
