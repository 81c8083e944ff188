// Micro-benchmark: fp64 throughput of DFMA (CUDA cores) vs mma.sync m8n8k4 f64 (DMMA) on sm_100a.
// Decides which instruction the blocked-Jacobi panel kernels should be built on.
#include <cstdio>
#include <cuda_runtime.h>
__global__ void dfma_kernel(double* out, int iters, double x) {
  double a[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) a[i] = threadIdx.x * 1e-9 + i;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) a[i] = fma(a[i], x, 1e-9);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += a[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void dmma_kernel(double* out, int iters, double x) {
  double c[8][2];
#pragma unroll
  for (int i = 0; i < 8; ++i) { c[i][0] = threadIdx.x * 1e-9; c[i][1] = i; }
  double a = x, b = 1.0 - x;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i)
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                   : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += c[i][0] + c[i][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
int main() {
  double* out; cudaMalloc(&out, 148 * 8 * 256 * sizeof(double));
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int iters = 20000, blocks = 148 * 4, threads = 256;
  for (int rep = 0; rep < 2; ++rep) {
    cudaEventRecord(e0); dfma_kernel<<<blocks, threads>>>(out, iters, 0.999); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double fl = 2.0 * 16 * iters * (double)blocks * threads;
    printf("DFMA : %.3f ms  %.2f TFLOP/s\n", ms, fl / ms / 1e9);
    cudaEventRecord(e0); dmma_kernel<<<blocks, threads>>>(out, iters, 0.999); cudaEventRecord(e1); cudaEventSynchronize(e1);
    cudaEventElapsedTime(&ms, e0, e1);
    fl = 2.0 * 8 * 8 * 4 * 8 * iters * (double)blocks * (threads / 32);
    printf("DMMA : %.3f ms  %.2f TFLOP/s\n", ms, fl / ms / 1e9);
  }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
