// The two ends of the ViT around the transformer blocks, for model dim 32:
//   embed      x[b, 0] = cls + pos[0] ; x[b, 1+p] = patch(b, p) W^T + bias + pos[1+p]        base_vit.py:190-223
//   head_loss  loss = mean_b CrossEntropy(Linear(LayerNorm(x[b, 0])), label_b)                base_vit.py:230-233, training.py:57-60
// Each replaces a chain of ~6 (forward) / ~12 (backward) library kernels that run for 3-20 us apiece on a few thousand
// elements.  fp32 FMAs, one warp per token (lane = channel); parameter gradients are deterministic (fixed-order sums).
#include "erv_block_common.cuh"

namespace erv {
namespace eh {

constexpr int C = 32, MAXPD = 192, TILE = 64;

struct EmbedArgs {
  const float* img; const float* w; const float* b; const float* cls; const float* pos;
  float* out;
  const float* dout; float* part;
  int B, Cin, S, P, G, N, PD;
};

// element k of the patch of token (b, n >= 1): k = c P^2 + i P + j (base_vit.py:190-196)
__device__ __forceinline__ float patch_elem(const EmbedArgs& p, int b, int pidx, int k) {
  const int pp = p.P * p.P, c = k / pp, ij = k - c * pp, i = ij / p.P, j = ij - i * p.P;
  const int gy = pidx / p.G, gx = pidx - gy * p.G;
  return __ldg(p.img + (((size_t)b * p.Cin + c) * p.S + gy * p.P + i) * p.S + gx * p.P + j);
}

__global__ void __launch_bounds__(256) embed_fwd_kernel(const EmbedArgs p) {
  extern __shared__ __align__(16) float sm[];
  float* Wt = sm;                 // [PD][32]
  float* pb = sm + p.PD * C;      // [8 warps][PD]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < C * p.PD; i += blockDim.x) {
    const int o = i / p.PD, k = i - o * p.PD;
    Wt[k * C + o] = __ldg(p.w + i);
  }
  __syncthreads();
  const float bias = __ldg(p.b + lane), cls = __ldg(p.cls + lane);
  float* mine = pb + warp * p.PD;
  const long total = (long)p.B * p.N;
  for (long t = (long)blockIdx.x * 8 + warp; t < total; t += (long)gridDim.x * 8) {
    const int b = (int)(t / p.N), n = (int)(t - (long)b * p.N);
    const float pos = __ldg(p.pos + n * C + lane);
    if (n == 0) {
      p.out[t * C + lane] = cls + pos;
      continue;
    }
    for (int k = lane; k < p.PD; k += 32) mine[k] = patch_elem(p, b, n - 1, k);
    __syncwarp();
    float acc = bias;
    for (int k = 0; k < p.PD; ++k) acc = fmaf(mine[k], Wt[k * C + lane], acc);
    p.out[t * C + lane] = acc + pos;
    __syncwarp();
  }
}

// dW[o][k] partials: thread (o = tid / 8, ks = tid % 8) owns k = ks + 8 q; the CTA walks 64-token tiles of the patch tokens
__global__ void __launch_bounds__(256) embed_wgrad_kernel(const EmbedArgs p) {
  extern __shared__ __align__(16) float sm[];
  float* sd = sm;                       // [TILE][32]
  float* sp = sm + TILE * C;            // [TILE][PD]
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int o = tid >> 3, ks = tid & 7;
  float acc[MAXPD / 8];
#pragma unroll
  for (int q = 0; q < MAXPD / 8; ++q) acc[q] = 0.f;
  const int np = p.N - 1;
  const long total = (long)p.B * np;
  const long ntiles = (total + TILE - 1) / TILE;
  for (long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    __syncthreads();
    for (int tt = warp; tt < TILE; tt += 8) {
      const long u = tile * TILE + tt;
      const bool ok = u < total;
      const int b = ok ? (int)(u / np) : 0, pidx = ok ? (int)(u - (long)b * np) : 0;
      sd[tt * C + lane] = ok ? __ldg(p.dout + ((size_t)b * p.N + 1 + pidx) * C + lane) : 0.f;
      for (int k = lane; k < p.PD; k += 32) sp[tt * p.PD + k] = ok ? patch_elem(p, b, pidx, k) : 0.f;
    }
    __syncthreads();
    for (int tt = 0; tt < TILE; ++tt) {
      const float d = sd[tt * C + o];
      const float* pr = sp + tt * p.PD + ks;
#pragma unroll
      for (int q = 0; q < MAXPD / 8; ++q)
        if (ks + 8 * q < p.PD) acc[q] = fmaf(d, pr[8 * q], acc[q]);
    }
  }
  float* part = p.part + (size_t)blockIdx.x * C * p.PD;
#pragma unroll
  for (int q = 0; q < MAXPD / 8; ++q)
    if (ks + 8 * q < p.PD) part[o * p.PD + ks + 8 * q] = acc[q];
}

// tmp [N][32] = sum_b dout[b] -> dpos (+)= tmp ; dcls (+)= tmp[0] ; db (+)= sum_{n >= 1} tmp[n]
__global__ void __launch_bounds__(256) embed_small_grads_kernel(const float* __restrict__ tmp, float* dpos, float* dcls, float* db,
                                                                int N, int accumulate) {
  for (int i = threadIdx.x; i < N * C; i += blockDim.x) dpos[i] = (accumulate ? dpos[i] : 0.f) + tmp[i];
  if (threadIdx.x < C) {
    const int c = threadIdx.x;
    float s = 0.f;
    for (int n = 1; n < N; ++n) s += tmp[n * C + c];
    db[c] = (accumulate ? db[c] : 0.f) + s;
    dcls[c] = (accumulate ? dcls[c] : 0.f) + tmp[c];
  }
}

struct HeadArgs {
  const float* x; const float* ln_w; const float* ln_b; const float* w; const float* b; const long long* labels;
  float* rows;                        // fwd: per-sample losses
  const float* dloss; float* dx; float* part;  // bwd: per-CTA partials [grid][K*32 + K + 32 + 32]
  int B, N, K; float eps;
};

// CTAs of 8 warps, one sample per warp (lane = channel): LayerNorm, K logits, log-softmax, NLL -> rows[b]
__global__ void __launch_bounds__(256) head_loss_fwd_kernel(const HeadArgs p) {
  __shared__ float Ws[32 * C], bs[32], nrm[8][C];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < p.K * C; i += blockDim.x) Ws[i] = __ldg(p.w + i);
  if (threadIdx.x < p.K) bs[threadIdx.x] = __ldg(p.b + threadIdx.x);
  __syncthreads();
  const int s = blockIdx.x * 8 + warp;
  if (s >= p.B) return;
  const float g = __ldg(p.ln_w + lane), bt = __ldg(p.ln_b + lane);
  const float x = __ldg(p.x + (size_t)s * p.N * C + lane);
  const float mean = warp_sum(x) * (1.0f / C);
  const float d = x - mean;
  const float rstd = rsqrtf(warp_sum(d * d) * (1.0f / C) + p.eps);
  nrm[warp][lane] = d * rstd * g + bt;
  __syncwarp();
  float logit = -INFINITY;
  if (lane < p.K) {
    float a = bs[lane];
#pragma unroll
    for (int c = 0; c < C; ++c) a = fmaf(nrm[warp][c], Ws[lane * C + c], a);
    logit = a;
  }
  const float mx = warp_max(logit);
  const float se = warp_sum(lane < p.K ? expf(logit - mx) : 0.f);
  // a label outside [0, K) (F.cross_entropy raises for it) makes this sample's loss NaN instead of reading another lane
  const long long lab = p.labels[s];
  const bool lab_ok = lab >= 0 && lab < (long long)p.K;
  const float picked = __shfl_sync(0xffffffffu, logit, lab_ok ? (int)lab : 0);
  if (lane == 0) p.rows[s] = lab_ok ? (mx + logf(se)) - picked : __int_as_float(0x7fc00000);
}
// loss = mean of rows, fixed order (one warp)
__global__ void __launch_bounds__(32) head_loss_mean_kernel(const float* __restrict__ rows, float* loss, int B) {
  float s = 0.f;
  for (int i = threadIdx.x; i < B; i += 32) s += rows[i];
  s = warp_sum(s);
  if (threadIdx.x == 0) *loss = s / (float)B;
}

// backward of the same: dx[b, 0] (the rest of dx is zeroed by the caller); per-CTA partials of dW | db | dln_w | dln_b
__global__ void __launch_bounds__(256) head_loss_bwd_kernel(const HeadArgs p) {
  __shared__ float Ws[32 * C], bs[32], nrm[8][C], dlg[8][32];
  __shared__ float red[8][35 * C];  // per warp: dW rows (K x 32) | db | dgam | dbet  (K <= 32)
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, P = p.K * C + p.K + 2 * C;
  for (int i = threadIdx.x; i < p.K * C; i += blockDim.x) Ws[i] = __ldg(p.w + i);
  if (threadIdx.x < p.K) bs[threadIdx.x] = __ldg(p.b + threadIdx.x);
  for (int i = threadIdx.x; i < 8 * 35 * C; i += blockDim.x) (&red[0][0])[i] = 0.f;
  __syncthreads();
  const int s = blockIdx.x * 8 + warp;
  if (s < p.B) {
    const float g = __ldg(p.ln_w + lane), bt = __ldg(p.ln_b + lane);
    const float scale = *p.dloss / (float)p.B;
    const float x = __ldg(p.x + (size_t)s * p.N * C + lane);
    const float mean = warp_sum(x) * (1.0f / C);
    const float d = x - mean;
    const float rstd = rsqrtf(warp_sum(d * d) * (1.0f / C) + p.eps);
    const float xh = d * rstd, n = xh * g + bt;
    nrm[warp][lane] = n;
    __syncwarp();
    float logit = -INFINITY;
    if (lane < p.K) {
      float a = bs[lane];
#pragma unroll
      for (int c = 0; c < C; ++c) a = fmaf(nrm[warp][c], Ws[lane * C + c], a);
      logit = a;
    }
    const float mx = warp_max(logit);
    const float e = lane < p.K ? expf(logit - mx) : 0.f;
    const float se = warp_sum(e);
    const int lab = (int)p.labels[s];
    const float dl = lane < p.K ? (e / se - (lane == lab ? 1.f : 0.f)) * scale : 0.f;  // d loss / d logit[lane]
    dlg[warp][lane] = dl;
    __syncwarp();
    float dn = 0.f;
    for (int k = 0; k < p.K; ++k) {
      const float v = dlg[warp][k];
      dn = fmaf(v, Ws[k * C + lane], dn);
      red[warp][k * C + lane] = v * n;  // dW[k][c] of this sample
    }
    if (lane < p.K) red[warp][p.K * C + lane] = dl;
    red[warp][p.K * C + p.K + lane] = dn * xh;
    red[warp][p.K * C + p.K + C + lane] = dn;
    const float dxh = dn * g;
    const float a1 = warp_sum(dxh) * (1.0f / C), a2 = warp_sum(dxh * xh) * (1.0f / C);
    p.dx[(size_t)s * p.N * C + lane] = rstd * (dxh - a1 - xh * a2);
  }
  __syncthreads();
  float* part = p.part + (size_t)blockIdx.x * P;
  for (int i = threadIdx.x; i < P; i += blockDim.x) {
    float v = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) v += red[w][i];
    part[i] = v;
  }
}

static int fill_embed(EmbedArgs& a, const char* fn, int B, int Cin, int S, int P) {
  ERV_CHECK_ARG(B > 0 && Cin > 0 && S > 0 && P > 0 && S % P == 0, "%s: bad image geometry", fn);
  a.B = B; a.Cin = Cin; a.S = S; a.P = P; a.G = S / P; a.N = a.G * a.G + 1; a.PD = Cin * P * P;
  if (a.PD > MAXPD) { set_error("%s: patch dimension %d not supported (<= %d)", fn, a.PD, MAXPD); return ERV_E_UNSUPPORTED; }
  return ERV_OK;
}

}  // namespace eh
}  // namespace erv

namespace erv {
namespace blk {  // tcgen05 versions for 4x4 patches (erv_block_tc.cu)
bool embed_tc_eligible(int Cin, int P, int N);
size_t embed_tc_bwd_workspace(int B, int Cin, int N);
int launch_embed_fwd_tc(const float* images, const float* w, const float* b, const float* cls, const float* pos, float* out,
                        int B, int Cin, int S, cudaStream_t st);
int launch_embed_bwd_tc(const float* images, const float* dout, float* dw, float* db, float* dcls, float* dpos, int B, int Cin,
                        int S, float* workspace, cudaStream_t st);
}  // namespace blk
}  // namespace erv

using namespace erv;
using namespace erv::eh;

extern "C" int erv_embed_supported(int dim, int patch_dim) { return dim == C && patch_dim <= MAXPD; }

extern "C" int erv_embed_fwd(const float* images, const float* w, const float* b, const float* cls, const float* pos, float* out,
                             int B, int Cin, int S, int P, void* stream) {
  ERV_CHECK_ARG(images && w && b && cls && pos && out, "erv_embed_fwd: null pointer");
  EmbedArgs a{};
  int rc = fill_embed(a, "erv_embed_fwd", B, Cin, S, P);
  if (rc) return rc;
  if (blk::embed_tc_eligible(Cin, P, a.N)) return blk::launch_embed_fwd_tc(images, w, b, cls, pos, out, B, Cin, S, (cudaStream_t)stream);
  a.img = images; a.w = w; a.b = b; a.cls = cls; a.pos = pos; a.out = out;
  const size_t smem = (size_t)(a.PD * C + 8 * a.PD) * sizeof(float);
  ERV_CUDA(allow_smem(embed_fwd_kernel, smem));
  const long units = ((long)B * a.N + 7) / 8;
  const int grid = (int)(units < 4L * kNumSMs ? units : 4L * kNumSMs);
  embed_fwd_kernel<<<grid, 256, smem, (cudaStream_t)stream>>>(a);
  ERV_LAUNCH_CHECK();
  return ERV_OK;
}

static int embed_grid(int B, int N) {
  const long tiles = ((long)B * (N - 1) + TILE - 1) / TILE;
  return (int)(tiles < 2L * kNumSMs ? (tiles < 1 ? 1 : tiles) : 2L * kNumSMs);
}

extern "C" size_t erv_embed_bwd_workspace(int B, int Cin, int S, int P) {
  if (B <= 0 || P <= 0 || S % P) return 0;
  const int G = S / P, N = G * G + 1, PD = Cin * P * P;
  const size_t fma = align_up((size_t)embed_grid(B, N) * C * PD * sizeof(float), 256) + align_up((size_t)N * C * sizeof(float), 256);
  const size_t tc = blk::embed_tc_eligible(Cin, P, N) ? blk::embed_tc_bwd_workspace(B, Cin, N) : 0;
  return fma > tc ? fma : tc;
}

// dw [32][PD], db [32], dcls [32], dpos [N][32]; accumulate != 0 adds to the existing contents (fused accumulation into .grad)
extern "C" int erv_embed_bwd(const float* images, const float* dout, float* dw, float* db, float* dcls, float* dpos,
                             int accumulate, int B, int Cin, int S, int P, void* workspace, size_t workspace_bytes,
                             void* stream) {
  ERV_CHECK_ARG(images && dout && dw && db && dcls && dpos && workspace, "erv_embed_bwd: null pointer");
  EmbedArgs a{};
  int rc = fill_embed(a, "erv_embed_bwd", B, Cin, S, P);
  if (rc) return rc;
  if (workspace_bytes < erv_embed_bwd_workspace(B, Cin, S, P)) { set_error("erv_embed_bwd: workspace too small"); return ERV_E_WORKSPACE; }
  cudaStream_t st = (cudaStream_t)stream;
  if (blk::embed_tc_eligible(Cin, P, a.N)) {
    if (!accumulate) {
      ERV_CUDA(cudaMemsetAsync(dw, 0, (size_t)C * a.PD * sizeof(float), st));
      ERV_CUDA(cudaMemsetAsync(db, 0, C * sizeof(float), st));
      ERV_CUDA(cudaMemsetAsync(dcls, 0, C * sizeof(float), st));
      ERV_CUDA(cudaMemsetAsync(dpos, 0, (size_t)a.N * C * sizeof(float), st));
    }
    return blk::launch_embed_bwd_tc(images, dout, dw, db, dcls, dpos, B, Cin, S, (float*)workspace, st);
  }
  const int grid = embed_grid(B, a.N);
  a.img = images; a.dout = dout; a.part = (float*)workspace;
  float* tmp = reinterpret_cast<float*>(static_cast<char*>(workspace) + align_up((size_t)grid * C * a.PD * sizeof(float), 256));
  const size_t smem = (size_t)(TILE * C + TILE * a.PD) * sizeof(float);
  ERV_CUDA(allow_smem(embed_wgrad_kernel, smem));
  embed_wgrad_kernel<<<grid, 256, smem, st>>>(a);
  ERV_LAUNCH_CHECK();
  {  // dW = sum of the CTA partials (added to dw when accumulating)
    float* dst[1] = {dw};
    const int seg[2] = {0, C * a.PD};
    rc = blk::launch_sum(a.part, dw, grid, C * a.PD, accumulate ? dst : nullptr, seg, 1, st);
    if (rc) return rc;
  }
  rc = blk::launch_sum(dout, tmp, B, a.N * C, nullptr, nullptr, 0, st);  // tmp[n][c] = sum_b dout[b][n][c]
  if (rc) return rc;
  embed_small_grads_kernel<<<1, 256, 0, st>>>(tmp, dpos, dcls, db, a.N, accumulate);
  ERV_LAUNCH_CHECK();
  return ERV_OK;
}

extern "C" size_t erv_head_loss_workspace(int B, int K) {
  if (B <= 0 || K <= 0) return 0;
  const size_t rows = align_up((size_t)B * sizeof(float), 256);
  const size_t parts = align_up((size_t)((B + 7) / 8) * (K * C + K + 2 * C) * sizeof(float), 256);
  return rows > parts ? rows : parts;
}

extern "C" int erv_head_loss_fwd(const float* x, const float* ln_w, const float* ln_b, const float* w, const float* b,
                                 const long long* labels, float* loss, int B, int N, int dim, int K, float eps, void* workspace,
                                 size_t workspace_bytes, void* stream) {
  ERV_CHECK_ARG(x && ln_w && ln_b && w && b && labels && loss && workspace && B > 0 && N > 0, "erv_head_loss_fwd: bad arguments");
  if (dim != C || K < 1 || K > 32) { set_error("erv_head_loss_fwd: dim %d / classes %d not supported (32, <= 32)", dim, K); return ERV_E_UNSUPPORTED; }
  if (workspace_bytes < erv_head_loss_workspace(B, K)) { set_error("erv_head_loss_fwd: workspace too small"); return ERV_E_WORKSPACE; }
  HeadArgs a{};
  a.x = x; a.ln_w = ln_w; a.ln_b = ln_b; a.w = w; a.b = b; a.labels = labels; a.rows = (float*)workspace; a.B = B; a.N = N; a.K = K;
  a.eps = eps;
  cudaStream_t st = (cudaStream_t)stream;
  head_loss_fwd_kernel<<<(B + 7) / 8, 256, 0, st>>>(a);
  ERV_LAUNCH_CHECK();
  head_loss_mean_kernel<<<1, 32, 0, st>>>(a.rows, loss, B);
  ERV_LAUNCH_CHECK();
  return ERV_OK;
}

// dx [B, N, dim] (fully written: zero except token 0); dparams [K*dim | K | dim | dim] or, with grad_accum (4 pointers), added there
extern "C" int erv_head_loss_bwd(const float* x, const float* ln_w, const float* ln_b, const float* w, const float* b,
                                 const long long* labels, const float* dloss, float* dx, float* dparams,
                                 float* const* grad_accum, int B, int N, int dim, int K, float eps, void* workspace,
                                 size_t workspace_bytes, void* stream) {
  ERV_CHECK_ARG(x && ln_w && ln_b && w && b && labels && dloss && dx && (dparams || grad_accum) && workspace && B > 0 && N > 0,
                "erv_head_loss_bwd: bad arguments");
  if (dim != C || K < 1 || K > 32) { set_error("erv_head_loss_bwd: dim %d / classes %d not supported (32, <= 32)", dim, K); return ERV_E_UNSUPPORTED; }
  if (workspace_bytes < erv_head_loss_workspace(B, K)) { set_error("erv_head_loss_bwd: workspace too small"); return ERV_E_WORKSPACE; }
  cudaStream_t st = (cudaStream_t)stream;
  ERV_CUDA(cudaMemsetAsync(dx, 0, (size_t)B * N * C * sizeof(float), st));
  HeadArgs a{};
  a.x = x; a.ln_w = ln_w; a.ln_b = ln_b; a.w = w; a.b = b; a.labels = labels; a.dloss = dloss; a.dx = dx; a.part = (float*)workspace;
  a.B = B; a.N = N; a.K = K; a.eps = eps;
  const int grid = (B + 7) / 8;
  head_loss_bwd_kernel<<<grid, 256, 0, st>>>(a);
  ERV_LAUNCH_CHECK();
  const int seg[5] = {0, K * C, K * C + K, K * C + K + C, K * C + K + 2 * C};
  return blk::launch_sum(a.part, dparams, grid, seg[4], grad_accum, seg, 4, st);
}
