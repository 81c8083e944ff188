// Streaming (HBM-bound) helpers: scaling/conversion copies, transposes,
// column reductions, centring, TF32 hi/lo operand split, Philox surrogates.
// All are one pass over the data with coalesced accesses along the
// contiguous (column) index; grids are sized in multiples of the SM count and
// grid-stride over rows.
#include "common.cuh"

namespace xmca {

thread_local std::string g_last_error;
std::atomic<long long> g_launches{0};

// ---------------------------------------------------------------- scale_copy
__global__ void scale_copy_kernel(const void* __restrict__ X, int xdt, int64_t ldx,
                                  void* __restrict__ Y, int ydt, int64_t ldy,
                                  int64_t rows, int64_t cols,
                                  const double* __restrict__ cs, const double* __restrict__ rs) {
  const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= cols) return;
  const double sc = cs ? cs[c] : 1.0;
  for (int64_t r = blockIdx.y; r < rows; r += gridDim.y) {
    double v = load_as_double(X, xdt, r * ldx + c) * sc;
    if (rs) v *= rs[r];
    store_from_double(Y, ydt, r * ldy + c, v);
  }
}

// ----------------------------------------------------------------- transpose
__global__ void transpose_kernel(const void* __restrict__ X, int xdt, int64_t rows, int64_t cols,
                                 int64_t ldx, void* __restrict__ Y, int ydt, int64_t ldy) {
  __shared__ double tile[32][33];
  const int64_t c0 = (int64_t)blockIdx.x * 32, r0 = (int64_t)blockIdx.y * 32;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    int64_t r = r0 + i, c = c0 + threadIdx.x;
    tile[i][threadIdx.x] = (r < rows && c < cols) ? load_as_double(X, xdt, r * ldx + c) : 0.0;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    int64_t c = c0 + i, r = r0 + threadIdx.x;     // Y[c][r]
    if (c < cols && r < rows) store_from_double(Y, ydt, c * ldy + r, tile[threadIdx.x][i]);
  }
}

// ------------------------------------------------------- TF32 hi/lo split
// hi = rna_tf32(x), lo = rna_tf32(x - hi): x = hi + lo to ~2^-22 relative, so
// hi*hi + hi*lo + lo*hi on the TF32 tensor pipe reproduces the fp32 product.
__device__ __forceinline__ float to_tf32(float x) {
  uint32_t u;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(x));
  return __uint_as_float(u);
}

__global__ void split_tf32_kernel(const void* __restrict__ X, int xdt, int64_t rows, int64_t cols,
                                  int64_t ldx, float* __restrict__ hi, float* __restrict__ lo,
                                  int64_t ldo) {
  const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= cols) return;
  for (int64_t r = blockIdx.y; r < rows; r += gridDim.y) {
    float x = (float)load_as_double(X, xdt, r * ldx + c);
    float h = to_tf32(x);
    hi[r * ldo + c] = h;
    lo[r * ldo + c] = to_tf32(x - h);
  }
}

__global__ void split_tf32_transpose_kernel(const void* __restrict__ X, int xdt, int64_t rows,
                                            int64_t cols, int64_t ldx, float* __restrict__ hi,
                                            float* __restrict__ lo, int64_t ldo) {
  __shared__ float tile[32][33];
  const int64_t c0 = (int64_t)blockIdx.x * 32, r0 = (int64_t)blockIdx.y * 32;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    int64_t r = r0 + i, c = c0 + threadIdx.x;
    tile[i][threadIdx.x] = (r < rows && c < cols) ? (float)load_as_double(X, xdt, r * ldx + c) : 0.f;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    int64_t c = c0 + i, r = r0 + threadIdx.x;     // out[c][r]
    if (c < cols && r < rows) {
      float x = tile[threadIdx.x][i];
      float h = to_tf32(x);
      hi[c * ldo + r] = h;
      lo[c * ldo + r] = to_tf32(x - h);
    }
  }
}

// --------------------------------------------------------------- col_sumsq
// one block column-strip of 32 columns x all rows; 8 row-lanes per column.
__global__ void col_sumsq_kernel(const void* __restrict__ X, int xdt, int64_t ldx, int64_t row0,
                                 int64_t row1, int64_t cols, double* __restrict__ out) {
  __shared__ double red[8][33];
  const int64_t c = (int64_t)blockIdx.x * 32 + threadIdx.x;
  double s = 0.0;
  if (c < cols)
    for (int64_t r = row0 + threadIdx.y; r < row1; r += 8) {
      double v = load_as_double(X, xdt, r * ldx + c);
      s = fma(v, v, s);
    }
  red[threadIdx.y][threadIdx.x] = s;
  __syncthreads();
  if (threadIdx.y == 0 && c < cols) {
    double t = 0.0;
#pragma unroll
    for (int i = 0; i < 8; ++i) t += red[i][threadIdx.x];
    out[c] = t;
  }
}

// ------------------------------------------------------------ centre columns
__global__ void center_columns_kernel(void* __restrict__ X, int xdt, int64_t rows, int64_t cols,
                                      int64_t ldx, double* __restrict__ mean) {
  __shared__ double red[8][33];
  const int64_t c = (int64_t)blockIdx.x * 32 + threadIdx.x;
  double s = 0.0;
  if (c < cols)
    for (int64_t r = threadIdx.y; r < rows; r += 8) s += load_as_double(X, xdt, r * ldx + c);
  red[threadIdx.y][threadIdx.x] = s;
  __syncthreads();
  double mu = 0.0;
#pragma unroll
  for (int i = 0; i < 8; ++i) mu += red[i][threadIdx.x];
  mu /= (double)rows;
  if (c < cols) {
    if (threadIdx.y == 0 && mean) mean[c] = mu;
    for (int64_t r = threadIdx.y; r < rows; r += 8)
      store_from_double(X, xdt, r * ldx + c, load_as_double(X, xdt, r * ldx + c) - mu);
  }
}

// ------------------------------------------------------------------- Philox
__device__ __forceinline__ void philox_round(uint32_t (&c)[4], uint32_t k0, uint32_t k1) {
  const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u;
  uint32_t hi0 = __umulhi(M0, c[0]), lo0 = M0 * c[0];
  uint32_t hi1 = __umulhi(M1, c[2]), lo1 = M1 * c[2];
  uint32_t n0 = hi1 ^ c[1] ^ k0, n1 = lo1, n2 = hi0 ^ c[3] ^ k1, n3 = lo0;
  c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
}
__device__ __forceinline__ void philox4x32_10(uint32_t (&c)[4], uint32_t k0, uint32_t k1) {
#pragma unroll
  for (int i = 0; i < 10; ++i) {
    philox_round(c, k0, k1);
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
}
// counter = (pair index lo, pair index hi, stream lo, stream hi); key = seed.
// Each Philox call yields 4 x 32 bits -> two 53-bit uniforms -> one Box-Muller
// pair (z0, z1) assigned to elements (2q, 2q+1) of the row-major index space.
__global__ void fill_normal_kernel(void* __restrict__ X, int xdt, int64_t rows, int64_t cols,
                                   int64_t ldx, uint64_t seed, uint64_t stream_id) {
  const int64_t total = rows * cols;
  const int64_t npairs = (total + 1) / 2;
  for (int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; q < npairs;
       q += (int64_t)gridDim.x * blockDim.x) {
    uint32_t c[4] = {(uint32_t)q, (uint32_t)((uint64_t)q >> 32), (uint32_t)stream_id,
                     (uint32_t)(stream_id >> 32)};
    philox4x32_10(c, (uint32_t)seed, (uint32_t)(seed >> 32));
    uint64_t a = ((uint64_t)c[0] << 32) | c[1], b = ((uint64_t)c[2] << 32) | c[3];
    double u1 = ((double)(a >> 11) + 0.5) * (1.0 / 9007199254740992.0);   // (0,1)
    double u2 = ((double)(b >> 11) + 0.5) * (1.0 / 9007199254740992.0);
    double rad = sqrt(-2.0 * log(u1));
    double sn, cs;
    sincospi(2.0 * u2, &sn, &cs);
    int64_t e0 = 2 * q, e1 = 2 * q + 1;
    store_from_double(X, xdt, (e0 / cols) * ldx + (e0 % cols), rad * cs);
    if (e1 < total) store_from_double(X, xdt, (e1 / cols) * ldx + (e1 % cols), rad * sn);
  }
}


// ------------------------------------------------------------- gather rows
// Y[i, :] = X[idx[i], :] * (row_scale ? row_scale[i] : 1)
__global__ void gather_rows_kernel(const void* __restrict__ X, int xdt, int64_t ldx,
                                   const int64_t* __restrict__ idx, int64_t n_out, int64_t cols,
                                   const double* __restrict__ rs, void* __restrict__ Y, int ydt,
                                   int64_t ldy) {
  const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= cols) return;
  for (int64_t i = blockIdx.y; i < n_out; i += gridDim.y) {
    double v = load_as_double(X, xdt, idx[i] * ldx + c);
    if (rs) v *= rs[i];
    store_from_double(Y, ydt, i * ldy + c, v);
  }
}

// --------------------------------------------------------------- row sumsq
// out[r] = sum_c X[r,c]^2   (one warp per row)
__global__ void row_sumsq_kernel(const void* __restrict__ X, int xdt, int64_t ldx, int64_t rows,
                                 int64_t cols, double* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const int64_t warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t r = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; r < rows; r += warps) {
    double s = 0.0;
    for (int64_t c = lane; c < cols; c += 32) {
      double v = load_as_double(X, xdt, r * ldx + c);
      s = fma(v, v, s);
    }
    s = warp_sum(s);
    if (lane == 0) out[r] = s;
  }
}

// -------------------------------------------------------------- col absmax
// out[c] = max_r |X[r,c] * (row_scale ? row_scale[r] : 1)|; out must be zero-initialised
__global__ void col_absmax_kernel(const void* __restrict__ X, int xdt, int64_t ldx, int64_t rows,
                                  int64_t cols, const double* __restrict__ rs,
                                  unsigned long long* __restrict__ out_bits) {
  const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= cols) return;
  double m = 0.0;
  for (int64_t r = blockIdx.y; r < rows; r += gridDim.y) {
    double v = load_as_double(X, xdt, r * ldx + c);
    if (rs) v *= rs[r];
    m = fmax(m, fabs(v));
  }
  atomicMax(out_bits + c, (unsigned long long)__double_as_longlong(m));
}

// ------------------------------------------------------- Promax target
// Xn = X * row_scale[r] / colmax[c];  P = Xn * |Xn|^(power-1)   (rotation.py:121-124)
// also writes the row-normalised X (fp64) when Xout != nullptr.
__global__ void promax_target_kernel(const double* __restrict__ X, int64_t ldx, int64_t rows,
                                     int64_t cols, const double* __restrict__ rs,
                                     const double* __restrict__ colmax, double power,
                                     double* __restrict__ Xout, double* __restrict__ Pout, int64_t ldo) {
  const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= cols) return;
  const double cm = colmax[c];
  for (int64_t r = blockIdx.y; r < rows; r += gridDim.y) {
    double x = X[r * ldx + c] * rs[r];
    double xn = x / cm;
    if (Xout) Xout[r * ldo + c] = x;
    Pout[r * ldo + c] = xn * pow(fabs(xn), power - 1.0);
  }
}

// Complex counterparts (planar re / im, fp64): column maximum of |x| and the Promax target
// P = Xn |Xn|^(power-1) with Xn = X * row_scale / colmax (rotation.py:115-124 with complex dtype).
__global__ void col_absmax_complex_kernel(const double* __restrict__ Xr, const double* __restrict__ Xi, int64_t ldx,
                                          int64_t rows, int64_t cols, const double* __restrict__ rs,
                                          unsigned long long* __restrict__ out_bits) {
  const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= cols) return;
  double m = 0.0;
  for (int64_t r = blockIdx.y; r < rows; r += gridDim.y) {
    const double a = Xr[r * ldx + c], b = Xi[r * ldx + c];
    double v = sqrt(a * a + b * b);
    if (rs) v *= rs[r];
    m = fmax(m, v);
  }
  atomicMax(out_bits + c, (unsigned long long)__double_as_longlong(m));
}

__global__ void promax_target_complex_kernel(const double* __restrict__ Xr, const double* __restrict__ Xi,
                                             int64_t ldx, int64_t rows, int64_t cols,
                                             const double* __restrict__ rs, const double* __restrict__ colmax,
                                             double power, double* __restrict__ Xor, double* __restrict__ Xoi,
                                             double* __restrict__ Por, double* __restrict__ Poi, int64_t ldo) {
  const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= cols) return;
  const double cm = colmax[c];
  for (int64_t r = blockIdx.y; r < rows; r += gridDim.y) {
    const double xr = Xr[r * ldx + c] * rs[r], xi = Xi[r * ldx + c] * rs[r];
    const double nr = xr / cm, ni = xi / cm;
    const double f = pow(sqrt(nr * nr + ni * ni), power - 1.0);
    Xor[r * ldo + c] = xr; Xoi[r * ldo + c] = xi;
    Por[r * ldo + c] = nr * f; Poi[r * ldo + c] = ni * f;
  }
}

// ------------------------------------------------------------ field ingest
// Constructor pre-processing on the device (array.py:191-240): per column the
// NaN flag, mean and standard deviation (ddof = 0, two-pass like numpy);
// one block = a strip of 32 columns x all rows, 8 row lanes per column.
__global__ void field_col_stats_kernel(const void* __restrict__ X, int xdt, int64_t rows, int64_t cols,
                                       int64_t ldx, double* __restrict__ mean, double* __restrict__ stdv,
                                       int* __restrict__ col_nan) {
  __shared__ double red[8][33];
  __shared__ int flag[8][33];
  const int64_t c = (int64_t)blockIdx.x * 32 + threadIdx.x;
  double s = 0.0;
  int bad = 0;
  if (c < cols)
    for (int64_t r = threadIdx.y; r < rows; r += 8) {
      double v = load_as_double(X, xdt, r * ldx + c);
      bad |= (v != v);
      s += v;
    }
  red[threadIdx.y][threadIdx.x] = s;
  flag[threadIdx.y][threadIdx.x] = bad;
  __syncthreads();
  double mu = 0.0;
  int anybad = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) { mu += red[i][threadIdx.x]; anybad |= flag[i][threadIdx.x]; }
  mu /= (double)rows;
  __syncthreads();
  double q = 0.0;
  if (c < cols && !anybad)
    for (int64_t r = threadIdx.y; r < rows; r += 8) {
      double d = load_as_double(X, xdt, r * ldx + c) - mu;
      q = fma(d, d, q);
    }
  red[threadIdx.y][threadIdx.x] = q;
  __syncthreads();
  if (threadIdx.y == 0 && c < cols) {
    double t = 0.0;
#pragma unroll
    for (int i = 0; i < 8; ++i) t += red[i][threadIdx.x];
    mean[c] = mu;
    stdv[c] = sqrt(t / (double)rows);
    col_nan[c] = anybad;
  }
}

// row_valid[r] = 1 if the row holds at least one non-NaN value (tools/array.py:65-73); one warp per row
__global__ void field_row_valid_kernel(const void* __restrict__ X, int xdt, int64_t rows, int64_t cols,
                                       int64_t ldx, int* __restrict__ row_valid) {
  const int lane = threadIdx.x & 31;
  const int64_t warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t r = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; r < rows; r += warps) {
    int ok = 0;
    for (int64_t c = lane; c < cols && !ok; c += 32) {
      double v = load_as_double(X, xdt, r * ldx + c);
      ok |= (v == v);
    }
    ok = __any_sync(0xffffffffu, ok);
    if (lane == 0) row_valid[r] = ok;
  }
}

// Y[r, j] = X[r, idx[j]] - mean[idx[j]], subtraction in the field's own precision (array.py:199-207)
__global__ void compact_center_kernel(const void* __restrict__ X, int xdt, int64_t rows, int64_t ldx,
                                      const int64_t* __restrict__ idx, int64_t n_keep,
                                      const double* __restrict__ mean, void* __restrict__ Y, int ydt,
                                      int64_t ldy) {
  const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n_keep) return;
  const int64_t c = idx[j];
  const double mu = mean[c];
  if (xdt == XMCA_F32) {
    const float muf = (float)mu;
    const float* x = reinterpret_cast<const float*>(X);
    for (int64_t r = blockIdx.y; r < rows; r += gridDim.y)
      store_from_double(Y, ydt, r * ldy + j, (double)(x[r * ldx + c] - muf));
  } else {
    const double* x = reinterpret_cast<const double*>(X);
    for (int64_t r = blockIdx.y; r < rows; r += gridDim.y)
      store_from_double(Y, ydt, r * ldy + j, x[r * ldx + c] - mu);
  }
}

static inline unsigned row_blocks(int64_t rows) {
  int64_t want = 8LL * sm_count();
  int64_t g = rows < want ? rows : want;
  return (unsigned)(g < 1 ? 1 : (g > 65535 ? 65535 : g));
}

}  // namespace xmca

using namespace xmca;

extern "C" const char* xmca_last_error(void) { return g_last_error.c_str(); }
extern "C" int xmca_version(void) { return 100; }
extern "C" long long xmca_launch_count(void) { return g_launches.load(); }

extern "C" int xmca_scale_copy(const void* d_X, int x_dtype, int64_t ldx, void* d_Y, int y_dtype,
                               int64_t ldy, int64_t rows, int64_t cols, const double* d_col_scale,
                               const double* d_row_scale, void* stream) {
  XMCA_REQUIRE(d_X && d_Y && rows > 0 && cols > 0, "xmca_scale_copy: bad argument");
  XMCA_REQUIRE(dtype_ok(x_dtype) && dtype_ok(y_dtype), "xmca_scale_copy: bad dtype");
  XMCA_REQUIRE(ldx >= cols && ldy >= cols, "xmca_scale_copy: leading dimension too small");
  dim3 grid((unsigned)((cols + 255) / 256), row_blocks(rows));
  scale_copy_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(d_X, x_dtype, ldx, d_Y, y_dtype, ldy,
                                                            rows, cols, d_col_scale, d_row_scale);
  XMCA_LAUNCHED();
  return XMCA_OK;
}

extern "C" int xmca_transpose(const void* d_X, int x_dtype, int64_t rows, int64_t cols, int64_t ldx,
                              void* d_Y, int y_dtype, int64_t ldy, void* stream) {
  XMCA_REQUIRE(d_X && d_Y && rows > 0 && cols > 0, "xmca_transpose: bad argument");
  XMCA_REQUIRE(dtype_ok(x_dtype) && dtype_ok(y_dtype), "xmca_transpose: bad dtype");
  XMCA_REQUIRE(ldx >= cols && ldy >= rows, "xmca_transpose: leading dimension too small");
  XMCA_REQUIRE((rows + 31) / 32 <= 65535, "xmca_transpose: too many rows");
  dim3 grid((unsigned)((cols + 31) / 32), (unsigned)((rows + 31) / 32));
  transpose_kernel<<<grid, dim3(32, 8), 0, (cudaStream_t)stream>>>(d_X, x_dtype, rows, cols, ldx,
                                                                   d_Y, y_dtype, ldy);
  XMCA_LAUNCHED();
  return XMCA_OK;
}

extern "C" int xmca_split_tf32(const void* d_X, int x_dtype, int64_t rows, int64_t cols, int64_t ldx,
                               int transpose, float* d_hi, float* d_lo, int64_t ldo, void* stream) {
  XMCA_REQUIRE(d_X && d_hi && d_lo && rows > 0 && cols > 0, "xmca_split_tf32: bad argument");
  XMCA_REQUIRE(dtype_ok(x_dtype), "xmca_split_tf32: bad dtype");
  XMCA_REQUIRE(ldx >= cols && ldo >= (transpose ? rows : cols),
               "xmca_split_tf32: leading dimension too small");
  if (transpose) {
    XMCA_REQUIRE((rows + 31) / 32 <= 65535, "xmca_split_tf32: too many rows");
    dim3 grid((unsigned)((cols + 31) / 32), (unsigned)((rows + 31) / 32));
    split_tf32_transpose_kernel<<<grid, dim3(32, 8), 0, (cudaStream_t)stream>>>(
        d_X, x_dtype, rows, cols, ldx, d_hi, d_lo, ldo);
  } else {
    dim3 grid((unsigned)((cols + 255) / 256), row_blocks(rows));
    split_tf32_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(d_X, x_dtype, rows, cols, ldx, d_hi,
                                                              d_lo, ldo);
  }
  XMCA_LAUNCHED();
  return XMCA_OK;
}

extern "C" int xmca_col_sumsq(const void* d_X, int x_dtype, int64_t ldx, int64_t row0, int64_t row1,
                              int64_t cols, double* d_out, void* stream) {
  XMCA_REQUIRE(d_X && d_out && cols > 0 && row1 >= row0 && row0 >= 0, "xmca_col_sumsq: bad argument");
  XMCA_REQUIRE(dtype_ok(x_dtype) && ldx >= cols, "xmca_col_sumsq: bad dtype / ld");
  col_sumsq_kernel<<<(unsigned)((cols + 31) / 32), dim3(32, 8), 0, (cudaStream_t)stream>>>(
      d_X, x_dtype, ldx, row0, row1, cols, d_out);
  XMCA_LAUNCHED();
  return XMCA_OK;
}

extern "C" int xmca_center_columns(void* d_X, int x_dtype, int64_t rows, int64_t cols, int64_t ldx,
                                   double* d_mean, void* stream) {
  XMCA_REQUIRE(d_X && rows > 0 && cols > 0, "xmca_center_columns: bad argument");
  XMCA_REQUIRE(dtype_ok(x_dtype) && ldx >= cols, "xmca_center_columns: bad dtype / ld");
  center_columns_kernel<<<(unsigned)((cols + 31) / 32), dim3(32, 8), 0, (cudaStream_t)stream>>>(
      d_X, x_dtype, rows, cols, ldx, d_mean);
  XMCA_LAUNCHED();
  return XMCA_OK;
}

extern "C" int xmca_field_stats(const void* d_X, int x_dtype, int64_t rows, int64_t cols, int64_t ldx,
                                double* d_mean, double* d_std, int* d_col_nan, int* d_row_valid,
                                void* stream) {
  XMCA_REQUIRE(d_X && d_mean && d_std && d_col_nan && d_row_valid && rows > 0 && cols > 0,
               "xmca_field_stats: bad argument");
  XMCA_REQUIRE(dtype_ok(x_dtype) && ldx >= cols, "xmca_field_stats: bad dtype / ld");
  field_col_stats_kernel<<<(unsigned)((cols + 31) / 32), dim3(32, 8), 0, (cudaStream_t)stream>>>(
      d_X, x_dtype, rows, cols, ldx, d_mean, d_std, d_col_nan);
  XMCA_LAUNCHED();
  int64_t blocks = (rows + 7) / 8;
  int64_t cap = 16LL * sm_count();
  if (blocks > cap) blocks = cap;
  field_row_valid_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(d_X, x_dtype, rows, cols, ldx,
                                                                             d_row_valid);
  XMCA_LAUNCHED();
  return XMCA_OK;
}

extern "C" int xmca_compact_center(const void* d_X, int x_dtype, int64_t rows, int64_t ldx,
                                   const int64_t* d_idx, int64_t n_keep, const double* d_mean,
                                   void* d_Y, int y_dtype, int64_t ldy, void* stream) {
  XMCA_REQUIRE(d_X && d_idx && d_mean && d_Y && rows > 0 && n_keep > 0, "xmca_compact_center: bad argument");
  XMCA_REQUIRE(dtype_ok(x_dtype) && dtype_ok(y_dtype) && ldy >= n_keep, "xmca_compact_center: bad dtype / ld");
  dim3 grid((unsigned)((n_keep + 255) / 256), row_blocks(rows));
  compact_center_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(d_X, x_dtype, rows, ldx, d_idx, n_keep, d_mean,
                                                                d_Y, y_dtype, ldy);
  XMCA_LAUNCHED();
  return XMCA_OK;
}

extern "C" int xmca_fill_normal(void* d_X, int x_dtype, int64_t rows, int64_t cols, int64_t ldx,
                                uint64_t seed, uint64_t stream_id, void* stream) {
  XMCA_REQUIRE(d_X && rows > 0 && cols > 0, "xmca_fill_normal: bad argument");
  XMCA_REQUIRE(dtype_ok(x_dtype) && ldx >= cols, "xmca_fill_normal: bad dtype / ld");
  int64_t npairs = (rows * cols + 1) / 2;
  int64_t blocks = (npairs + 255) / 256;
  int64_t cap = 16LL * sm_count();
  if (blocks > cap) blocks = cap;
  fill_normal_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(d_X, x_dtype, rows, cols,
                                                                         ldx, seed, stream_id);
  XMCA_LAUNCHED();
  return XMCA_OK;
}

extern "C" int xmca_gather_rows(const void* d_X, int x_dtype, int64_t ldx, const int64_t* d_idx,
                                int64_t n_out, int64_t cols, const double* d_row_scale, void* d_Y,
                                int y_dtype, int64_t ldy, void* stream) {
  XMCA_REQUIRE(d_X && d_idx && d_Y && n_out > 0 && cols > 0, "xmca_gather_rows: bad argument");
  XMCA_REQUIRE(dtype_ok(x_dtype) && dtype_ok(y_dtype) && ldx >= cols && ldy >= cols,
               "xmca_gather_rows: bad dtype / ld");
  dim3 grid((unsigned)((cols + 255) / 256), row_blocks(n_out));
  gather_rows_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(d_X, x_dtype, ldx, d_idx, n_out, cols,
                                                             d_row_scale, d_Y, y_dtype, ldy);
  XMCA_LAUNCHED();
  return XMCA_OK;
}

extern "C" int xmca_row_sumsq(const void* d_X, int x_dtype, int64_t ldx, int64_t rows, int64_t cols,
                              double* d_out, void* stream) {
  XMCA_REQUIRE(d_X && d_out && rows > 0 && cols > 0, "xmca_row_sumsq: bad argument");
  XMCA_REQUIRE(dtype_ok(x_dtype) && ldx >= cols, "xmca_row_sumsq: bad dtype / ld");
  int64_t blocks = (rows + 7) / 8;
  int64_t cap = 16LL * sm_count();
  if (blocks > cap) blocks = cap;
  row_sumsq_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(d_X, x_dtype, ldx, rows, cols, d_out);
  XMCA_LAUNCHED();
  return XMCA_OK;
}

extern "C" int xmca_col_absmax(const void* d_X, int x_dtype, int64_t ldx, int64_t rows, int64_t cols,
                               const double* d_row_scale, double* d_out, void* stream) {
  XMCA_REQUIRE(d_X && d_out && rows > 0 && cols > 0, "xmca_col_absmax: bad argument");
  XMCA_REQUIRE(dtype_ok(x_dtype) && ldx >= cols, "xmca_col_absmax: bad dtype / ld");
  XMCA_CUDA(cudaMemsetAsync(d_out, 0, (size_t)cols * 8, (cudaStream_t)stream));
  dim3 grid((unsigned)((cols + 255) / 256), row_blocks(rows));
  col_absmax_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(
      d_X, x_dtype, ldx, rows, cols, d_row_scale, reinterpret_cast<unsigned long long*>(d_out));
  XMCA_LAUNCHED();
  return XMCA_OK;
}

extern "C" int xmca_col_absmax_complex(const double* d_Xr, const double* d_Xi, int64_t ldx, int64_t rows, int64_t cols,
                                       const double* d_row_scale, double* d_out, void* stream) {
  XMCA_REQUIRE(d_Xr && d_Xi && d_out && rows > 0 && cols > 0 && ldx >= cols, "xmca_col_absmax_complex: bad argument");
  XMCA_CUDA(cudaMemsetAsync(d_out, 0, (size_t)cols * sizeof(double), (cudaStream_t)stream));
  dim3 grid((unsigned)((cols + 255) / 256), row_blocks(rows));
  col_absmax_complex_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(
      d_Xr, d_Xi, ldx, rows, cols, d_row_scale, reinterpret_cast<unsigned long long*>(d_out));
  XMCA_LAUNCHED();
  return XMCA_OK;
}

extern "C" int xmca_promax_target_complex(const double* d_Xr, const double* d_Xi, int64_t ldx, int64_t rows,
                                          int64_t cols, const double* d_row_scale, const double* d_colmax,
                                          double power, double* d_Xor, double* d_Xoi, double* d_Por, double* d_Poi,
                                          int64_t ldo, void* stream) {
  XMCA_REQUIRE(d_Xr && d_Xi && d_row_scale && d_colmax && d_Xor && d_Xoi && d_Por && d_Poi && rows > 0 && cols > 0,
               "xmca_promax_target_complex: bad argument");
  XMCA_REQUIRE(ldx >= cols && ldo >= cols, "xmca_promax_target_complex: leading dimension too small");
  dim3 grid((unsigned)((cols + 255) / 256), row_blocks(rows));
  promax_target_complex_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(d_Xr, d_Xi, ldx, rows, cols, d_row_scale,
                                                                       d_colmax, power, d_Xor, d_Xoi, d_Por, d_Poi, ldo);
  XMCA_LAUNCHED();
  return XMCA_OK;
}

extern "C" int xmca_promax_target(const double* d_X, int64_t ldx, int64_t rows, int64_t cols,
                                  const double* d_row_scale, const double* d_colmax, double power,
                                  double* d_Xout, double* d_Pout, int64_t ldo, void* stream) {
  XMCA_REQUIRE(d_X && d_row_scale && d_colmax && d_Pout && rows > 0 && cols > 0,
               "xmca_promax_target: bad argument");
  XMCA_REQUIRE(ldx >= cols && ldo >= cols, "xmca_promax_target: leading dimension too small");
  dim3 grid((unsigned)((cols + 255) / 256), row_blocks(rows));
  promax_target_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(d_X, ldx, rows, cols, d_row_scale,
                                                               d_colmax, power, d_Xout, d_Pout, ldo);
  XMCA_LAUNCHED();
  return XMCA_OK;
}
