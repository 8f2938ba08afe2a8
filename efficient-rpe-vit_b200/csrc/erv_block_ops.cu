// The rest of the transformer block around the attention core, for the reference's tiny
// model dims (C = 32..64) where the stock library kernels are far off the HBM roofline:
//   * weight + bias gradient of a Linear layer, dW[O][I] = sum_r dy[r][O] x[r][I], db = sum_r dy[r]: the reduction runs over
//     all B*N tokens (66 560 at config 2) while O*I is at most a few thousand, so the work is split over the token axis
//     (one slab per CTA), partial results are summed by a second kernel in a fixed order;
//   * LayerNorm forward / backward with one warp per row (C = 32 is exactly one element per lane).
#include "erv_common.cuh"

namespace erv {

constexpr int WG_KC = 32;  // token rows per staged chunk

// grid: G slabs; block: (O/4)*(I/4) threads (<= 256), thread (to, ti) owns a 4x4 tile of dW.
template <typename T>
__global__ void __launch_bounds__(256) wgrad_partial_kernel(const T* __restrict__ dy, const T* __restrict__ x,
                                                            float* __restrict__ part, int R, int O, int I,
                                                            int rows_per_cta) {
  extern __shared__ __align__(16) float sm[];
  float* dy_s = sm;                 // [WG_KC][O]
  float* x_s = sm + WG_KC * O;      // [WG_KC][I]
  const int nti = I / 4;
  const int to = threadIdx.x / nti, ti = threadIdx.x % nti;
  const int r_begin = blockIdx.x * rows_per_cta, r_end = min(R, r_begin + rows_per_cta);
  float acc[4][4];
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b) acc[a][b] = 0.f;
  float bsum = 0.f;  // thread t < O also owns bias column t
  for (int r0 = r_begin; r0 < r_end; r0 += WG_KC) {
    const int rows = min(WG_KC, r_end - r0);
    __syncthreads();
    for (int i = threadIdx.x; i < WG_KC * O / 4; i += blockDim.x) {
      const int k = i / (O / 4), c = i % (O / 4);
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (k < rows) v = ld4(dy + (size_t)(r0 + k) * O + 4 * c);
      st4(dy_s + k * O + 4 * c, v);
    }
    for (int i = threadIdx.x; i < WG_KC * I / 4; i += blockDim.x) {
      const int k = i / (I / 4), c = i % (I / 4);
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (k < rows) v = ld4(x + (size_t)(r0 + k) * I + 4 * c);
      st4(x_s + k * I + 4 * c, v);
    }
    __syncthreads();
#pragma unroll 8
    for (int k = 0; k < WG_KC; ++k) {
      const float4 a = ld4(dy_s + k * O + 4 * to);
      const float4 b = ld4(x_s + k * I + 4 * ti);
      const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int p = 0; p < 4; ++p)
#pragma unroll
        for (int q = 0; q < 4; ++q) acc[p][q] = fmaf(av[p], bv[q], acc[p][q]);
    }
    if ((int)threadIdx.x < O) {
#pragma unroll 8
      for (int k = 0; k < WG_KC; ++k) bsum += dy_s[k * O + threadIdx.x];
    }
  }
  float* dst = part + (size_t)blockIdx.x * (O * I + O);
#pragma unroll
  for (int p = 0; p < 4; ++p)
    st4(dst + (size_t)(4 * to + p) * I + 4 * ti, make_float4(acc[p][0], acc[p][1], acc[p][2], acc[p][3]));
  if ((int)threadIdx.x < O) dst[O * I + threadIdx.x] = bsum;
}

// out[i] = sum_g part[g][i]; the first n_w entries go to dw, the rest to db (if non-null).
// Block = 32 outputs x 8 partial slices: each thread sums every 8th partial (coalesced over i), then the 8 slices are
// combined through shared memory in a fixed order.
__global__ void __launch_bounds__(256) partial_sum_kernel(const float* __restrict__ part, int G, int stride,
                                                          float* __restrict__ dw, int n_w, float* __restrict__ db) {
  __shared__ float red[8][32];
  const int lane = threadIdx.x & 31, slice = threadIdx.x >> 5;
  const int i = blockIdx.x * 32 + lane;
  float acc = 0.f;
  if (i < stride)
    for (int g = slice; g < G; g += 8) acc += part[(size_t)g * stride + i];
  red[slice][lane] = acc;
  __syncthreads();
  if (slice == 0 && i < stride) {
    float t = red[0][lane];
#pragma unroll
    for (int k = 1; k < 8; ++k) t += red[k][lane];
    if (i < n_w) dw[i] = t;
    else if (db != nullptr) db[i - n_w] = t;
  }
}

// ---- LayerNorm: one warp per row, lane owns columns lane, lane+32, ... (C <= 32*LN_MAX) -----------------------------
constexpr int LN_MAX = 32;

template <int NV>
__global__ void __launch_bounds__(256) layernorm_fwd_kernel(const float* __restrict__ x, const float* __restrict__ gamma,
                                                            const float* __restrict__ beta, float* __restrict__ y,
                                                            float* __restrict__ mean_out, float* __restrict__ rstd_out,
                                                            int R, int C, float eps) {
  const int lane = threadIdx.x & 31, warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int nwarps = (gridDim.x * blockDim.x) >> 5;
  float g[NV], bt[NV];
#pragma unroll
  for (int j = 0; j < NV; ++j) {
    const int c = lane + 32 * j;
    g[j] = c < C ? gamma[c] : 0.f;
    bt[j] = c < C ? beta[c] : 0.f;
  }
  for (int r = warp; r < R; r += nwarps) {
    float v[NV];
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < NV; ++j) {
      const int c = lane + 32 * j;
      v[j] = c < C ? x[(size_t)r * C + c] : 0.f;
      s += v[j];
    }
    const float mean = warp_sum(s) / C;
    float q = 0.f;
#pragma unroll
    for (int j = 0; j < NV; ++j) {
      const int c = lane + 32 * j;
      const float d = c < C ? v[j] - mean : 0.f;
      q = fmaf(d, d, q);
    }
    const float rstd = rsqrtf(warp_sum(q) / C + eps);
#pragma unroll
    for (int j = 0; j < NV; ++j) {
      const int c = lane + 32 * j;
      if (c < C) y[(size_t)r * C + c] = (v[j] - mean) * rstd * g[j] + bt[j];
    }
    if (lane == 0) {
      mean_out[r] = mean;
      rstd_out[r] = rstd;
    }
  }
}

// dx = rstd * (dy*g - mean(dy*g) - xhat * mean(dy*g*xhat)); per-CTA partial dgamma/dbeta -> part[cta][2][C]
template <int NV>
__global__ void __launch_bounds__(256) layernorm_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ x,
                                                            const float* __restrict__ gamma,
                                                            const float* __restrict__ mean_in,
                                                            const float* __restrict__ rstd_in, float* __restrict__ dx,
                                                            float* __restrict__ part, int R, int C) {
  __shared__ float red[8][2][32 * NV];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = (gridDim.x * blockDim.x) >> 5;
  float g[NV], dg[NV], db[NV];
#pragma unroll
  for (int j = 0; j < NV; ++j) {
    const int c = lane + 32 * j;
    g[j] = c < C ? gamma[c] : 0.f;
    dg[j] = 0.f;
    db[j] = 0.f;
  }
  for (int r = warp; r < R; r += nwarps) {
    const float mean = mean_in[r], rstd = rstd_in[r];
    float xh[NV], dyv[NV];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int j = 0; j < NV; ++j) {
      const int c = lane + 32 * j;
      const bool in = c < C;
      dyv[j] = in ? dy[(size_t)r * C + c] : 0.f;
      xh[j] = in ? (x[(size_t)r * C + c] - mean) * rstd : 0.f;
      const float t = dyv[j] * g[j];
      s1 += t;
      s2 = fmaf(t, xh[j], s2);
      dg[j] = fmaf(dyv[j], xh[j], dg[j]);
      db[j] += dyv[j];
    }
    s1 = warp_sum(s1) / C;
    s2 = warp_sum(s2) / C;
#pragma unroll
    for (int j = 0; j < NV; ++j) {
      const int c = lane + 32 * j;
      if (c < C) dx[(size_t)r * C + c] = rstd * (dyv[j] * g[j] - s1 - xh[j] * s2);
    }
  }
#pragma unroll
  for (int j = 0; j < NV; ++j) {
    red[wib][0][lane + 32 * j] = dg[j];
    red[wib][1][lane + 32 * j] = db[j];
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) {
    const int w = i / C, c = i % C;
    float acc = 0.f;
    for (int k = 0; k < 8; ++k) acc += red[k][w][c];
    part[(size_t)blockIdx.x * 2 * C + i] = acc;
  }
}

static int ln_grid(int R) {
  int g = (R + 7) / 8;  // 8 warps per CTA
  const int cap = kNumSMs * 8;
  return g < cap ? (g < 1 ? 1 : g) : cap;
}

}  // namespace erv

using namespace erv;

static int wgrad_slabs(int R) {
  int g = (R + 2 * WG_KC - 1) / (2 * WG_KC);  // at least two chunks per slab
  return g < 4 * kNumSMs ? (g < 1 ? 1 : g) : 4 * kNumSMs;
}

extern "C" int erv_linear_wgrad_supported(int R, int O, int I) {
  return R >= 1024 && O % 4 == 0 && I % 4 == 0 && I >= 16 && (O / 4) * (I / 4) <= 256 && (O / 4) * (I / 4) >= 32 && O <= 256 &&
         (size_t)WG_KC * (O + I) * sizeof(float) <= 48 * 1024;
}

extern "C" size_t erv_linear_wgrad_workspace(int R, int O, int I) {
  return align_up((size_t)wgrad_slabs(R) * (O * I + O) * sizeof(float), 256);
}

extern "C" int erv_linear_wgrad(const void* dy, const void* x, float* dw, float* db, int R, int O, int I, int dtype,
                                void* workspace, size_t workspace_bytes, void* stream) {
  ERV_CHECK_ARG(dy && x && dw && workspace, "erv_linear_wgrad: null pointer");
  ERV_CHECK_ARG(dtype == ERV_F32 || dtype == ERV_BF16, "erv_linear_wgrad: bad dtype %d", dtype);
  if (!erv_linear_wgrad_supported(R, O, I)) {
    set_error("erv_linear_wgrad: shape R=%d O=%d I=%d unsupported (use the library GEMM)", R, O, I);
    return ERV_E_UNSUPPORTED;
  }
  if (workspace_bytes < erv_linear_wgrad_workspace(R, O, I)) { set_error("erv_linear_wgrad: workspace too small"); return ERV_E_WORKSPACE; }
  cudaStream_t st = (cudaStream_t)stream;
  const int G = wgrad_slabs(R);
  int rows_per_cta = (R + G - 1) / G;
  rows_per_cta = (rows_per_cta + WG_KC - 1) / WG_KC * WG_KC;
  const int threads = (O / 4) * (I / 4);
  const size_t smem = (size_t)WG_KC * (O + I) * sizeof(float);
  float* part = static_cast<float*>(workspace);
  if (dtype == ERV_F32)
    wgrad_partial_kernel<float><<<G, threads, smem, st>>>((const float*)dy, (const float*)x, part, R, O, I, rows_per_cta);
  else
    wgrad_partial_kernel<__nv_bfloat16><<<G, threads, smem, st>>>((const __nv_bfloat16*)dy, (const __nv_bfloat16*)x, part,
                                                                  R, O, I, rows_per_cta);
  ERV_LAUNCH_CHECK();
  const int stride = O * I + O;
  partial_sum_kernel<<<(stride + 31) / 32, 256, 0, st>>>(part, G, stride, dw, O * I, db);
  ERV_LAUNCH_CHECK();
  return ERV_OK;
}

extern "C" int erv_layernorm_fwd(const float* x, const float* gamma, const float* beta, float* y, float* mean,
                                 float* rstd, int R, int C, float eps, void* stream) {
  ERV_CHECK_ARG(x && gamma && beta && y && mean && rstd && R > 0, "erv_layernorm_fwd: bad arguments");
  if (C < 1 || C > 32 * LN_MAX) { set_error("erv_layernorm_fwd: C=%d unsupported (max %d)", C, 32 * LN_MAX); return ERV_E_UNSUPPORTED; }
  cudaStream_t st = (cudaStream_t)stream;
  const int grid = ln_grid(R);
#define LN_FWD(NV_) layernorm_fwd_kernel<NV_><<<grid, 256, 0, st>>>(x, gamma, beta, y, mean, rstd, R, C, eps)
  if (C <= 32) LN_FWD(1); else if (C <= 64) LN_FWD(2); else if (C <= 128) LN_FWD(4); else if (C <= 256) LN_FWD(8);
  else if (C <= 512) LN_FWD(16); else LN_FWD(32);
#undef LN_FWD
  ERV_LAUNCH_CHECK();
  return ERV_OK;
}

extern "C" size_t erv_layernorm_bwd_workspace(int R, int C) { return align_up((size_t)ln_grid(R) * 2 * C * sizeof(float), 256); }

extern "C" int erv_layernorm_bwd(const float* dy, const float* x, const float* gamma, const float* mean,
                                 const float* rstd, float* dx, float* dgamma, float* dbeta, int R, int C,
                                 void* workspace, size_t workspace_bytes, void* stream) {
  ERV_CHECK_ARG(dy && x && gamma && mean && rstd && dx && dgamma && dbeta && workspace && R > 0, "erv_layernorm_bwd: bad arguments");
  if (C < 1 || C > 256) { set_error("erv_layernorm_bwd: C=%d unsupported (max 256)", C); return ERV_E_UNSUPPORTED; }
  if (workspace_bytes < erv_layernorm_bwd_workspace(R, C)) { set_error("erv_layernorm_bwd: workspace too small"); return ERV_E_WORKSPACE; }
  cudaStream_t st = (cudaStream_t)stream;
  const int grid = ln_grid(R);
  float* part = static_cast<float*>(workspace);
#define LN_BWD(NV_) layernorm_bwd_kernel<NV_><<<grid, 256, 0, st>>>(dy, x, gamma, mean, rstd, dx, part, R, C)
  if (C <= 32) LN_BWD(1); else if (C <= 64) LN_BWD(2); else if (C <= 128) LN_BWD(4); else LN_BWD(8);
#undef LN_BWD
  ERV_LAUNCH_CHECK();
  // dgamma = first C entries, dbeta = next C
  partial_sum_kernel<<<(2 * C + 31) / 32, 256, 0, st>>>(part, grid, 2 * C, dgamma, C, dbeta);
  ERV_LAUNCH_CHECK();
  return ERV_OK;
}
