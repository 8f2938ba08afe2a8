// Toeplitz product y = T(c) x, T[i,j] = c[j-i+n-1]  (fft_utils.py:17-172), forward and backward.
// This is the public helper of the reference's module API (tests import it); the KERPLE attention
// path itself never materialises D1/D2 and goes through erv_tileattn.cu instead.  Direct O(n^2 d)
// evaluation with the coefficient window and an x tile staged in shared memory.
#include "erv_common.cuh"

namespace erv {

constexpr int TP_ROWS = 32;   // output rows per CTA
constexpr int TP_COLS = 64;   // feature columns per CTA (threads.x)
constexpr int TP_JT = 32;     // x rows per staged tile

// transpose == 0: y[i] = sum_j c[j-i+n-1] x[j];  transpose != 0: y[j] = sum_i c[j-i+n-1] x[i]
__global__ void __launch_bounds__(256) toeplitz_kernel(const float* __restrict__ c, const float* __restrict__ x,
                                                       float* __restrict__ y, int P, int c_count, int n, int d,
                                                       int transpose) {
  __shared__ float xs[TP_JT][TP_COLS];
  __shared__ float cs[TP_ROWS + TP_JT];
  const int p = blockIdx.z, r0 = blockIdx.y * TP_ROWS, d0 = blockIdx.x * TP_COLS;
  const int tx = threadIdx.x % TP_COLS, ty = threadIdx.x / TP_COLS;  // ty in 0..3, rows ty, ty+4, ...
  const float* crow = c + (size_t)(c_count == P ? p : p % c_count) * (2 * n - 1);
  const float* xp = x + (size_t)p * n * d;
  float acc[TP_ROWS / 4];
#pragma unroll
  for (int k = 0; k < TP_ROWS / 4; ++k) acc[k] = 0.f;
  for (int j0 = 0; j0 < n; j0 += TP_JT) {
    __syncthreads();
    for (int i = threadIdx.x; i < TP_JT * TP_COLS; i += blockDim.x) {
      int jj = i / TP_COLS, dd = i % TP_COLS;
      xs[jj][dd] = (j0 + jj < n && d0 + dd < d) ? xp[(size_t)(j0 + jj) * d + d0 + dd] : 0.f;
    }
    // coefficient window: index w = (jj - rr) + (TP_ROWS - 1) for rr < TP_ROWS, jj < TP_JT
    for (int w = threadIdx.x; w < TP_ROWS + TP_JT - 1; w += blockDim.x) {
      int delta = (j0 - r0) + w - (TP_ROWS - 1);          // (source row) - (output row)
      int idx = transpose ? (n - 1 - delta) : (n - 1 + delta);
      cs[w] = (idx >= 0 && idx < 2 * n - 1) ? crow[idx] : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < TP_ROWS / 4; ++k) {
      const int rr = ty + 4 * k;
      float a = acc[k];
#pragma unroll 8
      for (int jj = 0; jj < TP_JT; ++jj) a = fmaf(cs[jj - rr + TP_ROWS - 1], xs[jj][tx], a);
      acc[k] = a;
    }
  }
  if (d0 + tx < d) {
#pragma unroll
    for (int k = 0; k < TP_ROWS / 4; ++k) {
      const int r = r0 + ty + 4 * k;
      if (r < n) y[((size_t)p * n + r) * d + d0 + tx] = acc[k];
    }
  }
}

// dc[row][delta + n - 1] = sum_{p -> row} sum_i dy[p][i][:] . x[p][i+delta][:];  grid (2n-1, c_count), block 256
__global__ void __launch_bounds__(256) toeplitz_dc_kernel(const float* __restrict__ x, const float* __restrict__ dy,
                                                          float* __restrict__ dc, int P, int c_count, int n, int d) {
  __shared__ float red[256];
  const int idx = blockIdx.x, row = blockIdx.y, delta = idx - (n - 1);
  const int i_lo = max(0, -delta), i_hi = min(n - 1, n - 1 - delta);
  float acc = 0.f;
  for (int p = row; p < P; p += c_count) {
    for (int i = i_lo; i <= i_hi; ++i) {
      const float* a = dy + ((size_t)p * n + i) * d;
      const float* b = x + ((size_t)p * n + i + delta) * d;
      for (int dd = threadIdx.x; dd < d; dd += blockDim.x) acc = fmaf(a[dd], b[dd], acc);
    }
  }
  red[threadIdx.x] = acc;
  __syncthreads();
  for (int s = 128; s > 0; s >>= 1) {
    if (threadIdx.x < s) red[threadIdx.x] += red[threadIdx.x + s];
    __syncthreads();
  }
  if (threadIdx.x == 0) dc[(size_t)row * (2 * n - 1) + idx] = red[0];
}

static int toeplitz_launch(const float* c, const float* x, float* y, int P, int c_count, int n, int d, int transpose,
                           cudaStream_t st) {
  dim3 grid((d + TP_COLS - 1) / TP_COLS, (n + TP_ROWS - 1) / TP_ROWS, P);
  if (grid.z > 65535) { set_error("toeplitz: batch %d too large", P); return ERV_E_UNSUPPORTED; }
  toeplitz_kernel<<<grid, 256, 0, st>>>(c, x, y, P, c_count, n, d, transpose);
  ERV_LAUNCH_CHECK();
  return ERV_OK;
}

}  // namespace erv

using namespace erv;

extern "C" int erv_toeplitz_matmul_fwd(const float* c, const float* x, float* y, int P, int c_count, int n, int d,
                                       void* stream) {
  ERV_CHECK_ARG(c && x && y && P > 0 && n > 0 && d > 0, "erv_toeplitz_matmul_fwd: bad arguments");
  ERV_CHECK_ARG(c_count > 0 && P % c_count == 0, "erv_toeplitz_matmul_fwd: c_count %d must divide P %d", c_count, P);
  return toeplitz_launch(c, x, y, P, c_count, n, d, 0, (cudaStream_t)stream);
}

extern "C" int erv_toeplitz_matmul_bwd(const float* c, const float* x, const float* dy, float* dx, float* dc, int P,
                                       int c_count, int n, int d, void* stream) {
  ERV_CHECK_ARG(c && x && dy && P > 0 && n > 0 && d > 0, "erv_toeplitz_matmul_bwd: bad arguments");
  ERV_CHECK_ARG(c_count > 0 && P % c_count == 0, "erv_toeplitz_matmul_bwd: c_count %d must divide P %d", c_count, P);
  cudaStream_t st = (cudaStream_t)stream;
  if (dx) {
    int rc = toeplitz_launch(c, dy, dx, P, c_count, n, d, 1, st);
    if (rc) return rc;
  }
  if (dc) {
    toeplitz_dc_kernel<<<dim3(2 * n - 1, c_count), 256, 0, st>>>(x, dy, dc, P, c_count, n, d);
    ERV_LAUNCH_CHECK();
  }
  return ERV_OK;
}
