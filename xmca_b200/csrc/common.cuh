// Shared helpers for libxmca_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string>
#include <atomic>
#include "../../include/xmca_b200.h"

namespace xmca {

extern thread_local std::string g_last_error;
extern std::atomic<long long> g_launches;

inline int fail(int code, const char* what, const char* file, int line) {
  char buf[512];
  snprintf(buf, sizeof buf, "%s (%s:%d)", what, file, line);
  g_last_error = buf;
  return code;
}

#define XMCA_REQUIRE(cond, msg)                                                  \
  do {                                                                           \
    if (!(cond)) return ::xmca::fail(XMCA_BAD_ARG, msg, __FILE__, __LINE__);     \
  } while (0)

#define XMCA_CUDA(expr)                                                          \
  do {                                                                           \
    cudaError_t _e = (expr);                                                     \
    if (_e != cudaSuccess)                                                       \
      return ::xmca::fail(XMCA_CUDA_ERROR, cudaGetErrorString(_e), __FILE__, __LINE__); \
  } while (0)

// count + check a kernel launch
#define XMCA_LAUNCHED()                                                          \
  do {                                                                           \
    ::xmca::g_launches.fetch_add(1, std::memory_order_relaxed);                  \
    XMCA_CUDA(cudaGetLastError());                                               \
  } while (0)

inline int dtype_size(int dt) { return dt == XMCA_F64 ? 8 : 4; }
inline bool dtype_ok(int dt) { return dt == XMCA_F32 || dt == XMCA_F64; }

// runtime-typed scalar access (dtype is warp-uniform, so the branch is free)
__device__ __forceinline__ double load_as_double(const void* p, int dt, int64_t i) {
  return dt == XMCA_F64 ? reinterpret_cast<const double*>(p)[i]
                        : (double)reinterpret_cast<const float*>(p)[i];
}
__device__ __forceinline__ void store_from_double(void* p, int dt, int64_t i, double v) {
  if (dt == XMCA_F64) reinterpret_cast<double*>(p)[i] = v;
  else reinterpret_cast<float*>(p)[i] = (float)v;
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// fp64 reciprocal / reciprocal square root from the hardware approximations (~20 bits) and two
// Newton steps: relative error ~1e-15, a fraction of the latency of the IEEE division / sqrt.
__device__ __forceinline__ double fast_rcp(double x) {
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
  r = fma(r, fma(-x, r, 1.0), r);
  r = fma(r, fma(-x, r, 1.0), r);
  return r;
}
__device__ __forceinline__ double fast_rsqrt(double x) {
  double y;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
  const double hx = 0.5 * x;
  y = y * fma(-hx * y, y, 1.5);
  y = y * fma(-hx * y, y, 1.5);
  y = y * fma(-hx * y, y, 1.5);
  return y;
}

inline int sm_count() {                 // of the CURRENT device (cached per device id)
  static int cache[64] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  if (cache[dev] == 0) {
    int n = 0;
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    cache[dev] = n > 0 ? n : 148;
  }
  return cache[dev];
}

// Programmatic dependent launch for chains of small dependent kernels (13 launches per panel of the two-stage reduction,
// the block steps of the Cholesky factorisation, the per-panel products of the back-transformation): a kernel launched
// with the attribute may be set up before its predecessor in the stream has finished and blocks in `griddepcontrol.wait`
// until that grid has completed and flushed -- only the launch latency between the two disappears (measured: stage 1 of
// xmca_sytrd2 71.3 -> 68.4 ms at n = 8192).  Triggering the dependents early (`griddepcontrol.launch_dependents` at
// kernel entry) was measured too and is slower (74.3 ms): the waiting CTAs take SM slots from the last wave of the big
// products.  Every kernel launched this way starts with pdl_enter() (a no-op under an ordinary launch).
__device__ __forceinline__ void pdl_enter() {
  asm volatile("griddepcontrol.wait;" ::: "memory");
}
template <typename... KArgs, typename... Args>
static cudaError_t launch_pdl(bool pdl, void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st,
                              Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}

}  // namespace xmca
