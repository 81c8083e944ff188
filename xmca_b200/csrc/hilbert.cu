// Analytic signal (Hilbert transform along time) without an FFT library.
//
// The reference calls scipy.signal.hilbert(field, axis=0) (xmca/array.py:464): FFT over
// time, negative frequencies zeroed, positive ones doubled (DC and, for even T, Nyquist
// kept once), inverse FFT.  Two linear-operator forms of the same map are used here, both
// applied as GEMMs over the T x S field (tensor cores for fp32 fields, fp64 cores for fp64):
//
//   time domain      z = x + i H x,  H[t][t'] = h[(t - t') mod T],
//                    h[k] = (2/T) sum_{0 < f < T/2} sin(2 pi f k / T)          (circulant)
//                    -> xmca_hilbert_matrix; used for the PCs and the `_fields` mirror
//   frequency domain Z^ = diag(w) F x / sqrt(T), f = 1 .. floor(T/2)  (DC vanishes: centred)
//                    z = E Z^ with orthonormal columns E[t][f] = exp(2 pi i f t / T)/sqrt(T),
//                    hence C = Z_A^H Z_B = Z^_A^H Z^_B: the solve runs on the T/2 x S
//                    coefficient matrices (half the rows, and full row rank -- the analytic
//                    field itself has rank <= T/2)
//                    -> xmca_dft_matrix (stacked [Re; Im] rows), xmca_embed_complex
#include "common.cuh"
#include <math.h>

namespace xmca {

// h[k], k = 0..T-1, one thread per k (exact integer argument reduction, fp64 sinpi)
__global__ void hilbert_taps_kernel(int64_t T, double* __restrict__ h) {
  const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= T) return;
  const int64_t fmax = (T - 1) / 2;              // strictly below Nyquist
  double s = 0.0, comp = 0.0;
  for (int64_t f = 1; f <= fmax; ++f) {
    const int64_t m = (f * k) % T;
    const double term = sinpi(2.0 * (double)m / (double)T);
    const double y = term - comp;                // Kahan: up to T/2 terms of alternating sign
    const double t = s + y;
    comp = (t - s) - y;
    s = t;
  }
  h[k] = 2.0 * s / (double)T;
}

__global__ void circulant_kernel(int64_t T, const double* __restrict__ h, void* __restrict__ H, int dt, int64_t ld) {
  const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= T) return;
  for (int64_t r = blockIdx.y; r < T; r += gridDim.y) {
    int64_t k = r - c;
    if (k < 0) k += T;
    store_from_double(H, dt, r * ld + c, h[k]);
  }
}

// rows x cols block (row0, col0) of the N x N circulant: H[r][c] = h[((row0 + r) - (col0 + c)) mod N]
__global__ void circulant_block_kernel(int64_t N, int64_t rows, int64_t cols, int64_t row0, int64_t col0,
                                       const double* __restrict__ h, void* __restrict__ H, int dt, int64_t ld) {
  const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= cols) return;
  for (int64_t r = blockIdx.y; r < rows; r += gridDim.y) {
    int64_t k = ((row0 + r) - (col0 + c)) % N;
    if (k < 0) k += N;
    store_from_double(H, dt, r * ld + c, h[k]);
  }
}

// rows 0..Tp-1: (w_f / sqrt(T)) cos(2 pi f t / T);  rows Tp..2Tp-1: -(w_f / sqrt(T)) sin(2 pi f t / T),  f = row + 1
__global__ void dft_matrix_kernel(int64_t T, int64_t Tp, void* __restrict__ F, int dt, int64_t ld) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= T) return;
  const double inv = rsqrt((double)T);
  for (int64_t row = blockIdx.y; row < 2 * Tp; row += gridDim.y) {
    const bool im = row >= Tp;
    const int64_t f = (im ? row - Tp : row) + 1;
    const double w = (2 * f == T) ? 1.0 : 2.0;   // Nyquist kept once
    const int64_t m = (f * t) % T;
    double sn, cs;
    sincospi(2.0 * (double)m / (double)T, &sn, &cs);
    store_from_double(F, dt, row * ld + t, (im ? -sn : cs) * w * inv);
  }
}

// E = [[Zr, -Zi], [Zi, Zr]] from stacked Z = [Zr; Zi] (2 Tp x S)
__global__ void embed_complex_kernel(const void* __restrict__ Z, int zdt, int64_t ldz, int64_t Tp, int64_t S,
                                     void* __restrict__ E, int edt, int64_t lde) {
  const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= S) return;
  for (int64_t r = blockIdx.y; r < Tp; r += gridDim.y) {
    const double zr = load_as_double(Z, zdt, r * ldz + c), zi = load_as_double(Z, zdt, (r + Tp) * ldz + c);
    store_from_double(E, edt, r * lde + c, zr);
    store_from_double(E, edt, r * lde + S + c, -zi);
    store_from_double(E, edt, (r + Tp) * lde + c, zi);
    store_from_double(E, edt, (r + Tp) * lde + S + c, zr);
  }
}

static inline unsigned rows_grid(int64_t rows) {
  int64_t g = 8LL * sm_count();
  if (g > rows) g = rows;
  if (g < 1) g = 1;
  if (g > 65535) g = 65535;
  return (unsigned)g;
}

}  // namespace xmca

using namespace xmca;

extern "C" int xmca_hilbert_matrix(int64_t T, void* d_H, int h_dtype, int64_t ldh, double* d_taps, void* stream) {
  XMCA_REQUIRE(T >= 1 && d_H && d_taps && ldh >= T && dtype_ok(h_dtype), "xmca_hilbert_matrix: bad argument");
  cudaStream_t st = (cudaStream_t)stream;
  hilbert_taps_kernel<<<(unsigned)((T + 127) / 128), 128, 0, st>>>(T, d_taps);
  XMCA_LAUNCHED();
  circulant_kernel<<<dim3((unsigned)((T + 255) / 256), rows_grid(T)), 256, 0, st>>>(T, d_taps, d_H, h_dtype, ldh);
  XMCA_LAUNCHED();
  return XMCA_OK;
}

extern "C" int xmca_hilbert_block(int64_t N, int64_t rows, int64_t cols, int64_t row0, int64_t col0,
                                  void* d_H, int h_dtype, int64_t ldh, double* d_taps, void* stream) {
  XMCA_REQUIRE(N >= 1 && rows >= 1 && cols >= 1 && row0 >= 0 && col0 >= 0 && row0 + rows <= N && col0 + cols <= N &&
                   d_H && d_taps && ldh >= cols && dtype_ok(h_dtype), "xmca_hilbert_block: bad argument");
  cudaStream_t st = (cudaStream_t)stream;
  hilbert_taps_kernel<<<(unsigned)((N + 127) / 128), 128, 0, st>>>(N, d_taps);
  XMCA_LAUNCHED();
  circulant_block_kernel<<<dim3((unsigned)((cols + 255) / 256), rows_grid(rows)), 256, 0, st>>>(
      N, rows, cols, row0, col0, d_taps, d_H, h_dtype, ldh);
  XMCA_LAUNCHED();
  return XMCA_OK;
}

extern "C" int64_t xmca_dft_rows(int64_t T) { return 2 * (T / 2); }

extern "C" int xmca_dft_matrix(int64_t T, void* d_F, int f_dtype, int64_t ldf, void* stream) {
  XMCA_REQUIRE(T >= 2 && d_F && ldf >= T && dtype_ok(f_dtype), "xmca_dft_matrix: bad argument");
  const int64_t Tp = T / 2;
  dft_matrix_kernel<<<dim3((unsigned)((T + 255) / 256), rows_grid(2 * Tp)), 256, 0, (cudaStream_t)stream>>>(
      T, Tp, d_F, f_dtype, ldf);
  XMCA_LAUNCHED();
  return XMCA_OK;
}

extern "C" int xmca_embed_complex(const void* d_Z, int z_dtype, int64_t ldz, int64_t rows_half, int64_t cols,
                                  void* d_E, int e_dtype, int64_t lde, void* stream) {
  XMCA_REQUIRE(d_Z && d_E && rows_half >= 1 && cols >= 1 && ldz >= cols && lde >= 2 * cols &&
                   dtype_ok(z_dtype) && dtype_ok(e_dtype), "xmca_embed_complex: bad argument");
  embed_complex_kernel<<<dim3((unsigned)((cols + 255) / 256), rows_grid(rows_half)), 256, 0, (cudaStream_t)stream>>>(
      d_Z, z_dtype, ldz, rows_half, cols, d_E, e_dtype, lde);
  XMCA_LAUNCHED();
  return XMCA_OK;
}
