// Position tables and stand-alone rotations (RoPE, Circulant-STRING).
// Tables depend only on parameters, so they are built once per forward by tiny kernels; the heavy
// kernels consume them from L2.  Table math runs in fp64 (a few thousand elements) on top of a
// shared-memory twiddle table, so each (token, frequency) costs one sincos.
#include "erv_common.cuh"

namespace erv {

// ---- RoPE: rope.py:53-68 -------------------------------------------------------------------------
__global__ void rope_table_kernel(float theta, int n_pos, int half, float* __restrict__ cos_out,
                                  float* __restrict__ sin_out) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_pos * half) return;
  int n = i / half, m = i % half;
  float e = (float)(2 * m) / (float)(2 * half);
  float freq = (float)(1.0 / pow((double)theta, (double)e));
  float ang = (float)n * freq;  // the reference multiplies in fp32
  cos_out[i] = (float)cos((double)ang);
  sin_out[i] = (float)sin((double)ang);
}

// ---- Circulant-STRING table: circulant_string.py:207-295 --------------------------------------------
// twiddles tw[r] = (cos, sin)(2 pi r / D)
__device__ __forceinline__ void fill_twiddles(double* tw_c, double* tw_s, int D) {
  for (int r = threadIdx.x; r < D; r += blockDim.x) sincospi(2.0 * r / D, &tw_s[r], &tw_c[r]);
}
// imf[k][d] = Im FFT(c_k)[d] = -sum_t c[k,t] sin(2 pi d t / D)
__device__ __forceinline__ void fill_imfft(double* imf, const float* __restrict__ coeffs_h, const double* tw_s,
                                           int coord_dim, int D) {
  for (int i = threadIdx.x; i < coord_dim * D; i += blockDim.x) {
    int k = i / D, d = i % D;
    double im = 0.0;
    for (int t = 0; t < D; ++t) im -= (double)coeffs_h[k * D + t] * tw_s[(d * t) % D];
    imf[i] = im;
  }
}

// grid (ceil((N)/rows), H), block 256.  g[n][m] = (1/D) sum_d cos(theta_d + 2 pi d m / D),
// theta[n][d] = 2 sum_k pos[n-1][k] imf[k][d]
__global__ void circ_table_fwd_kernel(const float* __restrict__ coeffs, const float* __restrict__ pos, int N, int D,
                                      int coord_dim, int rows, float* __restrict__ g_out) {
  extern __shared__ double sm[];
  double* tw_c = sm;               // [D]
  double* tw_s = tw_c + D;         // [D]
  double* imf = tw_s + D;          // [coord_dim][D]
  double* cth = imf + coord_dim * D;  // [rows][D]
  double* sth = cth + rows * D;       // [rows][D]
  int h = blockIdx.y, n0 = blockIdx.x * rows;
  fill_twiddles(tw_c, tw_s, D);
  __syncthreads();
  fill_imfft(imf, coeffs + (size_t)h * coord_dim * D, tw_s, coord_dim, D);
  __syncthreads();
  for (int i = threadIdx.x; i < rows * D; i += blockDim.x) {
    int n = n0 + i / D, d = i % D;
    double th = 0.0;
    if (n >= 1 && n < N)
      for (int k = 0; k < coord_dim; ++k) th += 2.0 * (double)pos[(size_t)(n - 1) * coord_dim + k] * imf[k * D + d];
    sincos(th, &sth[i], &cth[i]);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < rows * D; i += blockDim.x) {
    int r = i / D, n = n0 + r, m = i % D;
    if (n >= N) continue;
    float* g = g_out + ((size_t)h * N + n) * D;
    if (n == 0) {  // CLS is not rotated (circulant_string.py:322-339)
      g[m] = (m == 0) ? 1.f : 0.f;
      continue;
    }
    double acc = 0.0;
    for (int d = 0; d < D; ++d) {
      int q = (d * m) % D;
      acc += cth[r * D + d] * tw_c[q] - sth[r * D + d] * tw_s[q];
    }
    g[m] = (float)(acc / D);
  }
}

// slot reduction: part[h][0][i] += sum_{s>=1} part[h][s][i]
__global__ void reduce_slots_kernel(float* __restrict__ part, int slots, size_t per_slot, int H) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= per_slot * H) return;
  size_t h = i / per_slot, j = i % per_slot;
  float* base = part + h * slots * per_slot + j;
  float acc = base[0];
  for (int s = 1; s < slots; ++s) acc += base[(size_t)s * per_slot];
  base[0] = acc;
}

// grid (nblk, H), block 256: partial dIm[h][blk][k][d] = 2 sum_{n in blk} pos[n-1][k] dtheta[n][d],
// dtheta[n][d] = -(1/D) sum_m dg[n][m] sin(theta_d + 2 pi d m / D)
__global__ void circ_table_bwd_kernel(const float* __restrict__ coeffs, const float* __restrict__ pos,
                                      const float* __restrict__ dg, size_t dg_head_stride, int N, int D, int coord_dim,
                                      int rows, double* __restrict__ dim_part) {
  extern __shared__ double sm[];
  double* tw_c = sm;
  double* tw_s = tw_c + D;
  double* imf = tw_s + D;
  double* red = imf + coord_dim * D;  // [nslice][coord_dim][D]
  int h = blockIdx.y, n0 = 1 + blockIdx.x * rows;
  fill_twiddles(tw_c, tw_s, D);
  __syncthreads();
  fill_imfft(imf, coeffs + (size_t)h * coord_dim * D, tw_s, coord_dim, D);
  __syncthreads();
  int d = threadIdx.x % D, slice = threadIdx.x / D, nslice = blockDim.x / D;
  double part[4] = {0, 0, 0, 0};
  for (int n = n0 + slice; n < n0 + rows && n < N; n += nslice) {
    const float* pn = pos + (size_t)(n - 1) * coord_dim;
    double th = 0.0;
    for (int k = 0; k < coord_dim; ++k) th += 2.0 * (double)pn[k] * imf[k * D + d];
    double s, c;
    sincos(th, &s, &c);
    const float* dgn = dg + (size_t)h * dg_head_stride + (size_t)n * D;
    double a = 0.0, b = 0.0;  // sum dg cos(phi), sum dg sin(phi)
    for (int m = 0; m < D; ++m) {
      int q = (d * m) % D;
      double v = (double)dgn[m];
      a += v * tw_c[q];
      b += v * tw_s[q];
    }
    double dth = -(s * a + c * b) / D;
    for (int k = 0; k < coord_dim; ++k) part[k] += 2.0 * (double)pn[k] * dth;
  }
  for (int k = 0; k < coord_dim; ++k) red[((size_t)slice * coord_dim + k) * D + d] = part[k];
  __syncthreads();
  for (int i = threadIdx.x; i < coord_dim * D; i += blockDim.x) {
    double acc = 0.0;
    for (int s = 0; s < nslice; ++s) acc += red[(size_t)s * coord_dim * D + i];
    dim_part[((size_t)h * gridDim.x + blockIdx.x) * coord_dim * D + i] = acc;
  }
}

// grid (H), block 128: dc[k][t] = -sum_d dIm[k][d] sin(2 pi d t / D), dIm summed over nblk partials
__global__ void circ_table_bwd_final_kernel(const double* __restrict__ dim_part, int nblk, int D, int coord_dim,
                                            float* __restrict__ dcoeffs) {
  extern __shared__ double sm[];
  double* tw_s = sm;            // [D]
  double* dimf = tw_s + D;      // [coord_dim][D]
  int h = blockIdx.x;
  for (int r = threadIdx.x; r < D; r += blockDim.x) tw_s[r] = sinpi(2.0 * r / D);
  for (int i = threadIdx.x; i < coord_dim * D; i += blockDim.x) {
    double acc = 0.0;
    for (int b = 0; b < nblk; ++b) acc += dim_part[((size_t)h * nblk + b) * coord_dim * D + i];
    dimf[i] = acc;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < coord_dim * D; i += blockDim.x) {
    int k = i / D, t = i % D;
    double acc = 0.0;
    for (int d = 0; d < D; ++d) acc -= dimf[k * D + d] * tw_s[(d * t) % D];
    dcoeffs[((size_t)h * coord_dim + k) * D + t] = (float)acc;
  }
}

// ---- stand-alone rotation of [B,H,N,D] fp32 ----------------------------------------------------------
__global__ void rotate_kernel(const float* __restrict__ x, float* __restrict__ y, int B, int H, int N, int D,
                              int rot, const float* __restrict__ ta, const float* __restrict__ tb, int inverse) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  size_t total = (size_t)B * H * N * D;
  if (i >= total) return;
  int a = (int)(i % D);
  size_t row = i / D;
  int n = (int)(row % N);
  int h = (int)((row / N) % H);
  const float* xr = x + row * D;
  if (rot == ERV_ROT_ROPE) {
    int m = a >> 1;
    float c = ta[(size_t)n * (D / 2) + m], s = tb[(size_t)n * (D / 2) + m];
    if (inverse) s = -s;
    float xe = xr[2 * m], xo = xr[2 * m + 1];
    y[i] = (a & 1) ? (xe * s + xo * c) : (xe * c - xo * s);
  } else if (rot == ERV_ROT_CIRCULANT) {
    const float* g = ta + ((size_t)h * N + n) * D;
    float acc = 0.f;
    if (!inverse) {
      for (int b = 0; b < D; ++b) { int j = a - b; if (j < 0) j += D; acc += g[j] * xr[b]; }
    } else {
      for (int b = 0; b < D; ++b) { int j = b - a; if (j < 0) j += D; acc += g[j] * xr[b]; }
    }
    y[i] = acc;
  } else {
    y[i] = xr[a];
  }
}

// dg[h,n,m] += sum_b sum_a dy[b,h,n,a] * x[b,h,n,(a-m) mod D]
__global__ void rotate_table_grad_kernel(const float* __restrict__ x, const float* __restrict__ dy, int B, int H,
                                         int N, int D, float* __restrict__ dg) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (size_t)H * N * D) return;
  int m = (int)(i % D);
  size_t hn = i / D;  // h*N + n
  int h = (int)(hn / N), n = (int)(hn % N);
  float acc = 0.f;
  for (int b = 0; b < B; ++b) {
    size_t row = (((size_t)b * H + h) * N + n) * D;
    for (int a = 0; a < D; ++a) { int j = a - m; if (j < 0) j += D; acc += dy[row + a] * x[row + j]; }
  }
  dg[i] += acc;
}

}  // namespace erv

using namespace erv;

extern "C" int erv_rope_table(float theta, int num_patches, int head_dim, float* cos_out, float* sin_out,
                              void* stream) {
  ERV_CHECK_ARG(head_dim > 0 && head_dim % 2 == 0, "erv_rope_table: head_dim %d must be even", head_dim);
  ERV_CHECK_ARG(num_patches > 0 && cos_out && sin_out, "erv_rope_table: bad arguments");
  int total = num_patches * (head_dim / 2);
  rope_table_kernel<<<(total + 255) / 256, 256, 0, (cudaStream_t)stream>>>(theta, num_patches, head_dim / 2, cos_out,
                                                                           sin_out);
  ERV_LAUNCH_CHECK();
  return ERV_OK;
}

static int circ_rows(int D) { return D >= 64 ? 8 : 32; }

extern "C" int erv_circulant_table_fwd(const float* coeffs, const float* positions, int H, int N, int head_dim,
                                       int coord_dim, float* g_out, void* stream) {
  ERV_CHECK_ARG(coeffs && g_out && H > 0 && N > 0 && head_dim > 0, "erv_circulant_table_fwd: bad arguments");
  ERV_CHECK_ARG(N == 1 || positions, "erv_circulant_table_fwd: positions missing");
  ERV_CHECK_ARG(coord_dim >= 1 && coord_dim <= 4, "erv_circulant_table_fwd: coord_dim %d unsupported", coord_dim);
  ERV_CHECK_ARG(head_dim <= 256, "erv_circulant_table_fwd: head_dim %d > 256", head_dim);
  int rows = circ_rows(head_dim);
  size_t smem = (size_t)(2 + coord_dim + 2 * rows) * head_dim * sizeof(double);
  circ_table_fwd_kernel<<<dim3((N + rows - 1) / rows, H), 256, smem, (cudaStream_t)stream>>>(
      coeffs, positions, N, head_dim, coord_dim, rows, g_out);
  ERV_LAUNCH_CHECK();
  return ERV_OK;
}

static int circ_bwd_rows() { return 64; }

extern "C" size_t erv_circulant_table_bwd_scratch(int H, int N, int head_dim, int coord_dim) {
  int nblk = (N - 1 + circ_bwd_rows() - 1) / circ_bwd_rows();
  if (nblk < 1) nblk = 1;
  return (size_t)H * nblk * coord_dim * head_dim * sizeof(double);
}

extern "C" int erv_circulant_table_bwd(const float* coeffs, const float* positions, float* dg_part, int slots, int H,
                                       int N, int head_dim, int coord_dim, float* dcoeffs, void* scratch,
                                       size_t scratch_bytes, void* stream) {
  ERV_CHECK_ARG(coeffs && dg_part && dcoeffs && scratch && slots >= 1, "erv_circulant_table_bwd: bad arguments");
  ERV_CHECK_ARG(coord_dim >= 1 && coord_dim <= 4, "erv_circulant_table_bwd: coord_dim %d unsupported", coord_dim);
  ERV_CHECK_ARG(head_dim <= 256, "erv_circulant_table_bwd: head_dim %d > 256", head_dim);
  ERV_CHECK_ARG(N >= 2, "erv_circulant_table_bwd: no patch tokens");
  if (scratch_bytes < erv_circulant_table_bwd_scratch(H, N, head_dim, coord_dim) || ((uintptr_t)scratch & 7)) {
    set_error("erv_circulant_table_bwd: scratch too small or misaligned");
    return ERV_E_WORKSPACE;
  }
  cudaStream_t st = (cudaStream_t)stream;
  const int D = head_dim;
  size_t per_slot = (size_t)N * D;
  if (slots > 1) {
    size_t total = per_slot * H;
    reduce_slots_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(dg_part, slots, per_slot, H);
    ERV_LAUNCH_CHECK();
  }
  int rows = circ_bwd_rows();
  int nblk = (N - 1 + rows - 1) / rows;
  double* dim_part = reinterpret_cast<double*>(scratch);
  int nslice = 256 / D;
  size_t smem = (size_t)(2 + coord_dim + nslice * coord_dim) * D * sizeof(double);
  circ_table_bwd_kernel<<<dim3(nblk, H), nslice * D, smem, st>>>(coeffs, positions, dg_part, (size_t)slots * per_slot, N,
                                                                 D, coord_dim, rows, dim_part);
  ERV_LAUNCH_CHECK();
  circ_table_bwd_final_kernel<<<H, 128, (size_t)(1 + coord_dim) * D * sizeof(double), st>>>(dim_part, nblk, D,
                                                                                           coord_dim, dcoeffs);
  ERV_LAUNCH_CHECK();
  return ERV_OK;
}

extern "C" int erv_rotate(const float* x, float* y, int B, int H, int N, int head_dim, int rot, const float* tab_a,
                          const float* tab_b, int inverse, void* stream) {
  ERV_CHECK_ARG(x && y && B > 0 && H > 0 && N > 0 && head_dim > 0, "erv_rotate: bad arguments");
  ERV_CHECK_ARG(rot == ERV_ROT_NONE || tab_a, "erv_rotate: table missing");
  ERV_CHECK_ARG(rot != ERV_ROT_ROPE || (tab_b && head_dim % 2 == 0), "erv_rotate: rope needs sin table, even head_dim");
  size_t total = (size_t)B * H * N * head_dim;
  rotate_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(x, y, B, H, N, head_dim, rot, tab_a,
                                                                                   tab_b, inverse);
  ERV_LAUNCH_CHECK();
  return ERV_OK;
}

extern "C" int erv_rotate_table_grad(const float* x_raw, const float* dy, int B, int H, int N, int head_dim,
                                     float* dg_accum, void* stream) {
  ERV_CHECK_ARG(x_raw && dy && dg_accum, "erv_rotate_table_grad: bad arguments");
  size_t total = (size_t)H * N * head_dim;
  rotate_table_grad_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(x_raw, dy, B, H, N,
                                                                                              head_dim, dg_accum);
  ERV_LAUNCH_CHECK();
  return ERV_OK;
}
