// Latency micro-benchmark (B200): dependent chains of the operations the latency-bound kernels (bulge chasing, 64 x 64
// factorisations, Varimax polar factor) are made of.  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o latency latency.cu
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdint.h>

#define N 512
__device__ __forceinline__ double fast_rcp(double x) { double r; asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x)); r = fma(r, fma(-x, r, 1.0), r); r = fma(r, fma(-x, r, 1.0), r); return r; }
__device__ __forceinline__ double fast_rsqrt(double x) { double y; asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x)); const double hx = 0.5 * x; y = y * fma(-hx * y, y, 1.5); y = y * fma(-hx * y, y, 1.5); y = y * fma(-hx * y, y, 1.5); return y; }

__global__ void lat_kernel(double* out, long long* clk, const int* chase, int chase_n, double* gbuf, int* flag) {
  __shared__ double sm[1024];
  __shared__ int smi[1024];
  const int tid = threadIdx.x;
  for (int i = tid; i < 1024; i += blockDim.x) { sm[i] = 1.0 + i * 1e-9; smi[i] = (i * 17 + 1) & 1023; }
  __syncthreads();
  double x = 1.0 + tid * 1e-12, y = 1.000001;
  long long t0, t1;
  int slot = 0;
#define BEGIN() __syncthreads(); t0 = clock64();
#define END() t1 = clock64(); if (tid == 0) clk[slot] = t1 - t0; slot++;
  BEGIN(); for (int i = 0; i < N; ++i) x = fma(x, y, 1e-9); END();                       // 0 DFMA chain
  BEGIN(); for (int i = 0; i < N; ++i) x = x + y; END();                                 // 1 DADD chain
  BEGIN(); for (int i = 0; i < N; ++i) x += __shfl_xor_sync(0xffffffffu, x, 1 + (i & 15)); END();   // 2 shfl64 + dadd chain
  { int p = tid & 1023; BEGIN(); for (int i = 0; i < N; ++i) p = smi[p]; END(); x += p; }  // 3 LDS chain
  BEGIN(); for (int i = 0; i < N; ++i) { x = fma(x, sm[(tid + i) & 1023], 1e-9); } END();   // 4 LDS.64 + DFMA (independent loads)
  BEGIN(); for (int i = 0; i < N; ++i) __syncthreads(); END();                           // 5 __syncthreads (all threads)
  if (tid < 64) { t0 = clock64(); for (int i = 0; i < N; ++i) asm volatile("bar.sync 1, 64;" ::: "memory"); t1 = clock64(); if (tid == 0) clk[slot] = t1 - t0; } slot++;   // 6 bar.sync 64
  BEGIN(); for (int i = 0; i < N; ++i) x = fast_rsqrt(x + 2.0); END();                   // 7 fast_rsqrt chain
  BEGIN(); for (int i = 0; i < N; ++i) x = fast_rcp(x + 2.0); END();                     // 8 fast_rcp chain
  BEGIN(); for (int i = 0; i < N; ++i) x = sqrt(x + 2.0); END();                         // 9 sqrt chain
  BEGIN(); for (int i = 0; i < N; ++i) x = 1.0 / (x + 2.0); END();                       // 10 div chain
  { int p = tid % chase_n; BEGIN(); if (tid < 32) for (int i = 0; i < N; ++i) p = __ldcg(chase + p); END(); x += p; }     // 11 ld.cg chain (L2)
  { BEGIN(); if (tid == 0) for (int i = 0; i < N; ++i) { int v; asm volatile("ld.relaxed.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(flag) : "memory"); x += v; } END(); }  // 12 ld.relaxed.gpu chain
  { BEGIN(); if (tid == 0) for (int i = 0; i < 64; ++i) { asm volatile("st.release.gpu.global.s32 [%0], %1;" :: "l"(flag), "r"(i) : "memory"); } END(); }  // 13 st.release x64 (nothing outstanding)
  { BEGIN(); for (int i = 0; i < 64; ++i) { for (int q = 0; q < 12; ++q) gbuf[(size_t)(q * 512 + tid) * 1 + (size_t)i * 8192] = x; __syncthreads(); if (tid == 0) asm volatile("st.release.gpu.global.s32 [%0], %1;" :: "l"(flag), "r"(i) : "memory"); } END(); }  // 14 49 KB stores + sync + release, x64
  { BEGIN(); for (int i = 0; i < 64; ++i) { double a = 0; for (int q = 0; q < 12; ++q) a += __ldcg(gbuf + (size_t)(q * 512 + tid) + (size_t)i * 8192); x += a; __syncthreads(); } END(); }   // 15 49 KB ld.cg + sync, x64
  { BEGIN(); if (tid == 0) for (int i = 0; i < 64; ++i) __threadfence(); END(); }        // 16 threadfence x64
  { BEGIN(); if (tid < 32) for (int i = 0; i < N; ++i) { double c0 = x, c1 = x * 2, c2 = x * 3, c3 = x * 4;
      for (int o = 16; o > 0; o >>= 1) { c0 += __shfl_xor_sync(0xffffffffu, c0, o); c1 += __shfl_xor_sync(0xffffffffu, c1, o); c2 += __shfl_xor_sync(0xffffffffu, c2, o); c3 += __shfl_xor_sync(0xffffffffu, c3, o); }
      x = c0 + c1 + c2 + c3; } END(); }                                                  // 17 four interleaved warp sums (1 warp)
  { BEGIN(); for (int i = 0; i < N; ++i) { double c0 = x, c1 = x * 2, c2 = x * 3, c3 = x * 4;
      for (int o = 16; o > 0; o >>= 1) { c0 += __shfl_xor_sync(0xffffffffu, c0, o); c1 += __shfl_xor_sync(0xffffffffu, c1, o); c2 += __shfl_xor_sync(0xffffffffu, c2, o); c3 += __shfl_xor_sync(0xffffffffu, c3, o); }
      x = c0 + c1 + c2 + c3; } END(); }                                                  // 18 the same on all 16 warps
  out[tid] = x;
}

int main() {
  const int chase_n = 1 << 21;   // 8 MB of ints: L2 resident, beyond L1
  int* h = (int*)malloc(sizeof(int) * chase_n);
  for (int i = 0; i < chase_n; ++i) h[i] = (int)(((long long)i * 1048583 + 12345) % chase_n);
  int *d_chase, *d_flag; double *d_out, *d_g; long long* d_clk;
  cudaMalloc(&d_chase, sizeof(int) * chase_n); cudaMemcpy(d_chase, h, sizeof(int) * chase_n, cudaMemcpyHostToDevice);
  cudaMalloc(&d_flag, 256); cudaMemset(d_flag, 0, 256);
  cudaMalloc(&d_out, 8 * 1024); cudaMalloc(&d_g, 8ull * 8192 * 80); cudaMalloc(&d_clk, 8 * 64); cudaMemset(d_clk, 0, 8 * 64);
  for (int rep = 0; rep < 2; ++rep) lat_kernel<<<1, 512>>>(d_out, d_clk, d_chase, chase_n, d_g, d_flag);
  cudaError_t e = cudaDeviceSynchronize();
  long long clk[64]; cudaMemcpy(clk, d_clk, sizeof(clk), cudaMemcpyDeviceToHost);
  const char* names[] = {"DFMA chain", "DADD chain", "shfl64+dadd chain", "LDS pointer chain", "LDS.64+DFMA (indep loads)", "__syncthreads (512 thr)",
                         "bar.sync 1,64", "fast_rsqrt(+add) chain", "fast_rcp(+add) chain", "sqrt(+add) chain", "div(+add) chain", "ld.cg pointer chain (L2, 1 warp)",
                         "ld.relaxed.gpu same address", "st.release.gpu (idle)", "49KB stores+sync+release", "49KB ld.cg+sync", "__threadfence (idle)",
                         "4 interleaved warp sums (1 warp)", "4 interleaved warp sums (16 warps)"};
  const int per[] = {N, N, N, N, N, N, N, N, N, N, N, N, N, 64, 64, 64, 64, N, N};
  printf("%s\n", cudaGetErrorString(e));
  for (int i = 0; i < 19; ++i) printf("%-36s %8.1f clocks per op\n", names[i], (double)clk[i] / per[i]);
  return 0;
}
