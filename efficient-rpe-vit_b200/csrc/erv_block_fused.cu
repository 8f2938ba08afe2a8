// Fused transformer-block kernels around the attention core, specialised to the reference's
// model dims: dim C = 32 (one channel per lane), MLP width 64 (configs/datasets/*.py:20-24).
//
//   ln_qkv      qkv = LayerNorm1(x) W_qkv^T (+ b)                                   unified_transformer.py:75-83
//   mlp         x1 = x + drop(a W_proj^T + b_proj) ; y = x1 + drop(fc2(drop(gelu(fc1(LayerNorm2(x1))))))
//                                                                                  favor_plus.py:263-265, unified_transformer.py:85-88
// and their backward kernels, which recompute the forward from (x) resp. (a, x) and the dropout seed, so nothing but
// the block input and the attention output is kept between forward and backward.
//
// Mapping: a warp owns 8 token rows at a time; lane = channel, so LayerNorm statistics are warp reductions and every
// global access is a full 128-byte line.  The tiny GEMMs run as 8-token x (OUT/32 per lane) register tiles on packed
// fp32 FMAs (FFMA2: the pair is two consecutive reduction indices), weights staged once per CTA in shared memory in a
// pair-interleaved layout that is conflict-free for 8-byte lane reads, activations broadcast from shared memory.
// Weight gradients: the CTA's 64-token tile of activations / gradients stays in shared memory; every thread owns a few
// entries of each dW in registers and accumulates over the tile; per-CTA partials are summed by a second kernel in a
// fixed order (deterministic, no atomics).
// Dropout masks come from a counter hash of (seed, stream, element) so the backward regenerates them.
#include <initializer_list>

#include "erv_block_common.cuh"

namespace erv {
namespace blk {

typedef unsigned long long u64;
__device__ __forceinline__ u64 pk(float lo, float hi) {
  return ((u64)__float_as_uint(hi) << 32) | (u64)__float_as_uint(lo);
}
__device__ __forceinline__ void ffma2v(u64& acc, u64 a, u64 b) { asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(acc) : "l"(a), "l"(b)); }
__device__ __forceinline__ float sum2(u64 v) { return __uint_as_float((uint32_t)v) + __uint_as_float((uint32_t)(v >> 32)); }

// Stage W [OUT][IN] (nn.Linear layout) for y = x W^T: reduction index c = input; dst[(c/2)*OUT*2 + j*2 + (c&1)] = W[j][c]
__device__ __forceinline__ void stage_fwd(float* dst, const float* __restrict__ W, int OUT, int IN) {
  for (int i = threadIdx.x; i < OUT * IN; i += blockDim.x) {
    const int j = i / IN, c = i % IN;
    dst[((c >> 1) * OUT + j) * 2 + (c & 1)] = __ldg(W + i);
  }
}
// ... for dx = dy W: reduction index j = output; dst[(j/2)*IN*2 + c*2 + (j&1)] = W[j][c]
__device__ __forceinline__ void stage_bwd(float* dst, const float* __restrict__ W, int OUT, int IN) {
  for (int i = threadIdx.x; i < OUT * IN; i += blockDim.x) {
    const int j = i / IN, c = i % IN;
    dst[((j >> 1) * IN + c) * 2 + (j & 1)] = __ldg(W + i);
  }
}

// out[m][t] = sum_k act[t][k] * Wp(k, lane + 32 m): act rows in shared memory (stride AS floats, broadcast reads),
// Wp pair-interleaved over k with NOUT columns.  K reduction length, NOUT outputs (NOUT/32 per lane).
template <int K, int NOUT, int AS>
__device__ __forceinline__ void gemv(const float* act, const float* Wp, int lane, float (&out)[NOUT / 32][T]) {
  u64 acc[NOUT / 32][T];
#pragma unroll
  for (int m = 0; m < NOUT / 32; ++m)
#pragma unroll
    for (int t = 0; t < T; ++t) acc[m][t] = 0ull;
#pragma unroll 2
  for (int k = 0; k < K; k += 4) {
    u64 x01[T], x23[T];  // 64-bit views of the loaded pairs: no repacking moves
#pragma unroll
    for (int t = 0; t < T; ++t) {
      const ulonglong2 v = *reinterpret_cast<const ulonglong2*>(act + t * AS + k);
      x01[t] = v.x;
      x23[t] = v.y;
    }
#pragma unroll
    for (int m = 0; m < NOUT / 32; ++m) {
      const u64 wa2 = *reinterpret_cast<const u64*>(Wp + ((k >> 1) * NOUT + lane + 32 * m) * 2);
      const u64 wb2 = *reinterpret_cast<const u64*>(Wp + (((k >> 1) + 1) * NOUT + lane + 32 * m) * 2);
#pragma unroll
      for (int t = 0; t < T; ++t) {
        ffma2v(acc[m][t], x01[t], wa2);
        ffma2v(acc[m][t], x23[t], wb2);
      }
    }
  }
#pragma unroll
  for (int m = 0; m < NOUT / 32; ++m)
#pragma unroll
    for (int t = 0; t < T; ++t) out[m][t] = sum2(acc[m][t]);
}

// LayerNorm statistics of T rows held one channel per lane (two-pass, population variance: matches at::layer_norm)
__device__ __forceinline__ void ln_stats(const float (&x)[T], float (&mean)[T], float (&rstd)[T], float eps) {
#pragma unroll
  for (int t = 0; t < T; ++t) mean[t] = warp_sum(x[t]) * (1.0f / C);
#pragma unroll
  for (int t = 0; t < T; ++t) {
    const float d = x[t] - mean[t];
    rstd[t] = rsqrtf(warp_sum(d * d) * (1.0f / C) + eps);
  }
}

struct LnQkvArgs {
  const float* x; const float* ln_w; const float* ln_b; const float* w; const float* b;  // b may be null
  float* qkv;           // fwd out
  const float* dqkv; const float* dres;  // bwd in (dres may be null)
  float* dx; float* part;                // bwd out: dx [R][C], per-CTA partial parameter gradients [grid][P_QKV]
  int R; float eps;
};

__global__ void __launch_bounds__(THREADS) ln_qkv_fwd_kernel(const LnQkvArgs p) {
  extern __shared__ __align__(16) float sm[];
  float* Wf = sm;                       // [C/2][QKV][2]
  float* act = sm + QKV * C;            // [WARPS][T][C]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  stage_fwd(Wf, p.w, QKV, C);
  __syncthreads();
  const float g = __ldg(p.ln_w + lane), bt = __ldg(p.ln_b + lane);
  float bias[QKV / 32];
#pragma unroll
  for (int m = 0; m < QKV / 32; ++m) bias[m] = p.b ? __ldg(p.b + lane + 32 * m) : 0.f;
  float* a = act + warp * T * C;
  const int ntiles = (p.R + T - 1) / T;
  float nx[T];  // rows of the next tile, in flight while this one is computed
  {
    const int r0 = (blockIdx.x * WARPS + warp) * T;
#pragma unroll
    for (int t = 0; t < T; ++t) nx[t] = (r0 + t < p.R) ? __ldg(p.x + (size_t)(r0 + t) * C + lane) : 0.f;
  }
  for (int tile = blockIdx.x * WARPS + warp; tile < ntiles; tile += gridDim.x * WARPS) {
    const int r0 = tile * T, rn = (tile + gridDim.x * WARPS) * T;
    float x[T], mean[T], rstd[T];
#pragma unroll
    for (int t = 0; t < T; ++t) {
      x[t] = nx[t];
      nx[t] = (rn + t < p.R) ? __ldg(p.x + (size_t)(rn + t) * C + lane) : 0.f;
    }
    ln_stats(x, mean, rstd, p.eps);
#pragma unroll
    for (int t = 0; t < T; ++t) a[t * C + lane] = (x[t] - mean[t]) * rstd[t] * g + bt;
    __syncwarp();
    float out[QKV / 32][T];
    gemv<C, QKV, C>(a, Wf, lane, out);
#pragma unroll
    for (int t = 0; t < T; ++t)
      if (r0 + t < p.R) {
#pragma unroll
        for (int m = 0; m < QKV / 32; ++m) p.qkv[(size_t)(r0 + t) * QKV + lane + 32 * m] = out[m][t] + bias[m];
      }
    __syncwarp();
  }
}

// dx = dres + LN1_bwd(dqkv W_qkv) ; dW_qkv += dqkv^T LN1(x) ; db_qkv ; dln_w ; dln_b
__global__ void __launch_bounds__(THREADS) ln_qkv_bwd_kernel(const LnQkvArgs p) {
  extern __shared__ __align__(16) float sm[];
  float* Wb = sm;                        // [QKV/2][C][2]
  float* sxn = Wb + QKV * C;             // [TILE][C]   LN1 output
  float* sdq = sxn + TILE * C;           // [TILE][QKV] dqkv
  float* red = sdq + TILE * QKV;         // [WARPS][QKV + 2 C]
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  stage_bwd(Wb, p.w, QKV, C);
  __syncthreads();
  const float g = __ldg(p.ln_w + lane), bt = __ldg(p.ln_b + lane);
  u64 dw[3][2];  // thread (o = tid/8 + 32 k, i0 = (tid%8)*4): dW[o][i0..i0+3] as packed pairs (FFMA2)
#pragma unroll
  for (int k = 0; k < 3; ++k) dw[k][0] = dw[k][1] = 0ull;
  float dgam = 0.f, dbet = 0.f, dbq[QKV / 32] = {0.f, 0.f, 0.f};
  const int wo = tid >> 3, wi = (tid & 7) * 4;
  const int nct = (p.R + TILE - 1) / TILE;
  for (int ct = blockIdx.x; ct < nct; ct += gridDim.x) {
    const int r0 = ct * TILE + warp * T;
    float* xn_w = sxn + warp * T * C;
    float* dq_w = sdq + warp * T * QKV;
    float x[T], mean[T], rstd[T], xh[T];
#pragma unroll
    for (int t = 0; t < T; ++t) x[t] = (r0 + t < p.R) ? __ldg(p.x + (size_t)(r0 + t) * C + lane) : 0.f;
#pragma unroll
    for (int t = 0; t < T; ++t)
#pragma unroll
      for (int m = 0; m < QKV / 32; ++m) {
        const float v = (r0 + t < p.R) ? __ldg(p.dqkv + (size_t)(r0 + t) * QKV + lane + 32 * m) : 0.f;
        dq_w[t * QKV + lane + 32 * m] = v;
        dbq[m] += v;
      }
    ln_stats(x, mean, rstd, p.eps);
#pragma unroll
    for (int t = 0; t < T; ++t) {
      xh[t] = (x[t] - mean[t]) * rstd[t];
      xn_w[t * C + lane] = (r0 + t < p.R) ? xh[t] * g + bt : 0.f;
    }
    __syncwarp();
    float dn[1][T];
    gemv<QKV, C, QKV>(dq_w, Wb, lane, dn);
#pragma unroll
    for (int t = 0; t < T; ++t) {
      const float d = dn[0][t];
      dgam += d * xh[t];
      dbet += d;
      const float dxh = d * g;
      const float m1 = warp_sum(dxh) * (1.0f / C), m2 = warp_sum(dxh * xh[t]) * (1.0f / C);
      if (r0 + t < p.R) {
        const float dres = p.dres ? __ldg(p.dres + (size_t)(r0 + t) * C + lane) : 0.f;
        p.dx[(size_t)(r0 + t) * C + lane] = dres + rstd[t] * (dxh - m1 - xh[t] * m2);
      }
    }
    __syncthreads();
    // weight gradient over the CTA's 64-token tile
#pragma unroll 4
    for (int t = 0; t < TILE; ++t) {
      const ulonglong2 xv = *reinterpret_cast<const ulonglong2*>(sxn + t * C + wi);
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        const float d = sdq[t * QKV + wo + 32 * k];
        const u64 dd = pk(d, d);
        ffma2v(dw[k][0], dd, xv.x);
        ffma2v(dw[k][1], dd, xv.y);
      }
    }
    __syncthreads();
  }
  float* part = p.part + (size_t)blockIdx.x * P_QKV;
#pragma unroll
  for (int k = 0; k < 3; ++k) *reinterpret_cast<ulonglong2*>(part + (wo + 32 * k) * C + wi) = make_ulonglong2(dw[k][0], dw[k][1]);
  float* rw = red + warp * (QKV + 2 * C);
#pragma unroll
  for (int m = 0; m < QKV / 32; ++m) rw[lane + 32 * m] = dbq[m];
  rw[QKV + lane] = dgam;
  rw[QKV + C + lane] = dbet;
  __syncthreads();
  if (tid < QKV + 2 * C) {
    float s = 0.f;
    for (int w = 0; w < WARPS; ++w) s += red[w * (QKV + 2 * C) + tid];
    part[QKV * C + tid] = s;
  }
}

__global__ void __launch_bounds__(THREADS) mlp_fwd_kernel(const MlpArgs p) {
  extern __shared__ __align__(16) float sm[];
  float* Wp = sm;                    // proj fwd  [C/2][C][2]
  float* W1 = Wp + C * C;            // fc1 fwd   [C/2][MLP][2]
  float* W2 = W1 + MLP * C;          // fc2 fwd   [MLP/2][C][2]
  float* act = W2 + C * MLP;         // [WARPS][T][MLP]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  stage_fwd(Wp, p.w_proj, C, C);
  stage_fwd(W1, p.w1, MLP, C);
  stage_fwd(W2, p.w2, C, MLP);
  __syncthreads();
  const float bp = __ldg(p.b_proj + lane), g = __ldg(p.ln_w + lane), bt = __ldg(p.ln_b + lane), b2 = __ldg(p.b2 + lane);
  const float b1[2] = {__ldg(p.b1 + lane), __ldg(p.b1 + lane + 32)};
  const bool drop = p.p_drop > 0.f;
  const unsigned long long seed = drop ? (unsigned long long)*p.seed : 0ull;
  const uint32_t thresh = drop ? (uint32_t)fminf(p.p_drop * 4294967296.0f, 4294967295.0f) : 0u;
  const float inv_keep = drop ? 1.0f / (1.0f - p.p_drop) : 1.0f;
  const uint32_t s0 = (uint32_t)p.salt * 4u;
  float* aw = act + warp * T * MLP;
  const int ntiles = (p.R + T - 1) / T;
  float na[T], nx[T];  // rows of the next tile, in flight while this one is computed
  {
    const int r0 = (blockIdx.x * WARPS + warp) * T;
#pragma unroll
    for (int t = 0; t < T; ++t) {
      const bool ok = r0 + t < p.R;
      na[t] = ok ? __ldg(p.a + (size_t)(r0 + t) * C + lane) : 0.f;
      nx[t] = ok ? __ldg(p.x + (size_t)(r0 + t) * C + lane) : 0.f;
    }
  }
  for (int tile = blockIdx.x * WARPS + warp; tile < ntiles; tile += gridDim.x * WARPS) {
    const int r0 = tile * T, rn = (tile + gridDim.x * WARPS) * T;
    float x[T];
#pragma unroll
    for (int t = 0; t < T; ++t) {
      aw[t * MLP + lane] = na[t];
      x[t] = nx[t];
      const bool ok = rn + t < p.R;
      na[t] = ok ? __ldg(p.a + (size_t)(rn + t) * C + lane) : 0.f;
      nx[t] = ok ? __ldg(p.x + (size_t)(rn + t) * C + lane) : 0.f;
    }
    __syncwarp();
    float pr[1][T];
    gemv<C, C, MLP>(aw, Wp, lane, pr);
    float x1[T], mean[T], rstd[T];
#pragma unroll
    for (int t = 0; t < T; ++t) {
      float v = pr[0][t] + bp;
      if (drop) v *= drop_scale(seed, s0, (uint32_t)(r0 + t) * C + lane, thresh, inv_keep);
      x1[t] = x[t] + v;
    }
    ln_stats(x1, mean, rstd, p.eps);
    __syncwarp();
#pragma unroll
    for (int t = 0; t < T; ++t) aw[t * MLP + lane] = (x1[t] - mean[t]) * rstd[t] * g + bt;
    __syncwarp();
    float h[2][T];
    gemv<C, MLP, MLP>(aw, W1, lane, h);
    __syncwarp();
#pragma unroll
    for (int t = 0; t < T; ++t)
#pragma unroll
      for (int m = 0; m < 2; ++m) {
        float v = gelu_f(h[m][t] + b1[m]);
        if (drop) v *= drop_scale(seed, s0 + 1, (uint32_t)(r0 + t) * MLP + lane + 32 * m, thresh, inv_keep);
        aw[t * MLP + lane + 32 * m] = v;
      }
    __syncwarp();
    float o[1][T];
    gemv<MLP, C, MLP>(aw, W2, lane, o);
#pragma unroll
    for (int t = 0; t < T; ++t) {
      float v = o[0][t] + b2;
      if (drop) v *= drop_scale(seed, s0 + 2, (uint32_t)(r0 + t) * C + lane, thresh, inv_keep);
      if (r0 + t < p.R) p.y[(size_t)(r0 + t) * C + lane] = x1[t] + v;
    }
    __syncwarp();
  }
}

__global__ void __launch_bounds__(THREADS, 2) mlp_bwd_kernel(const MlpArgs p) {
  extern __shared__ __align__(16) float sm[];
  float* Wp = sm;                   // proj fwd
  float* W1 = Wp + C * C;           // fc1 fwd
  float* W2b = W1 + MLP * C;        // fc2 bwd  [C/2][MLP][2]   (reduction over fc2 outputs)
  float* W1b = W2b + C * MLP;       // fc1 bwd  [MLP/2][C][2]
  float* Wpb = W1b + MLP * C;       // proj bwd [C/2][C][2]
  float* sa = Wpb + C * C;          // [TILE][C]   attention output rows
  float* sdp = sa + TILE * C;       // [TILE][C]   d(proj output)
  float* sn2 = sdp + TILE * C;      // [TILE][C]   LayerNorm2 output
  float* sdo = sn2 + TILE * C;      // [TILE][C]   d(fc2 output)
  float* shd = sdo + TILE * C;      // [TILE][MLP] dropped gelu(fc1)
  float* sdh = shd + TILE * MLP;    // [TILE][MLP] d(fc1 pre-activation)
  float* red = sdh + TILE * MLP;    // [WARPS][192]
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  stage_fwd(Wp, p.w_proj, C, C);
  stage_fwd(W1, p.w1, MLP, C);
  stage_bwd(W2b, p.w2, C, MLP);
  stage_bwd(W1b, p.w1, MLP, C);
  stage_bwd(Wpb, p.w_proj, C, C);
  __syncthreads();
  const float bp = __ldg(p.b_proj + lane), g = __ldg(p.ln_w + lane), bt = __ldg(p.ln_b + lane);
  const float b1[2] = {__ldg(p.b1 + lane), __ldg(p.b1 + lane + 32)};
  const bool drop = p.p_drop > 0.f;
  const unsigned long long seed = drop ? (unsigned long long)*p.seed : 0ull;
  const uint32_t thresh = drop ? (uint32_t)fminf(p.p_drop * 4294967296.0f, 4294967295.0f) : 0u;
  const float inv_keep = drop ? 1.0f / (1.0f - p.p_drop) : 1.0f;
  const uint32_t s0 = (uint32_t)p.salt * 4u;
  // thread-owned weight-gradient entries
  u64 dwp[2], dw1[4], dw2[4];  // packed pairs of consecutive input indices (FFMA2)
  dwp[0] = dwp[1] = 0ull;
#pragma unroll
  for (int i = 0; i < 4; ++i) dw1[i] = dw2[i] = 0ull;
  const int po = tid >> 3, pi = (tid & 7) * 4;    // dW_proj[po][pi..+3]        (32 x 32)
  const int o1 = tid >> 2, i1 = (tid & 3) * 8;    // dW1[o1][i1..+7]            (64 x 32)
  const int o2 = tid >> 3, i2 = (tid & 7) * 8;    // dW2[o2][i2..+7]            (32 x 64)
  float dbp = 0.f, dgam = 0.f, dbet = 0.f, db1[2] = {0.f, 0.f}, db2 = 0.f;  // per-lane (channel) sums
  const int nct = (p.R + TILE - 1) / TILE;
  for (int ct = blockIdx.x; ct < nct; ct += gridDim.x) {
    const int tr = warp * T, r0 = ct * TILE + tr;  // this warp's rows of the tile
    float x[T], dy[T];
#pragma unroll
    for (int t = 0; t < T; ++t) {
      const bool ok = r0 + t < p.R;
      sa[(tr + t) * C + lane] = ok ? __ldg(p.a + (size_t)(r0 + t) * C + lane) : 0.f;
      x[t] = ok ? __ldg(p.x + (size_t)(r0 + t) * C + lane) : 0.f;
      dy[t] = ok ? __ldg(p.dy + (size_t)(r0 + t) * C + lane) : 0.f;
    }
    __syncwarp();
    // ---- recompute the forward
    float pr[1][T];
    gemv<C, C, C>(sa + tr * C, Wp, lane, pr);
    float x1[T], mean[T], rstd[T], xh[T], m1[T];
#pragma unroll
    for (int t = 0; t < T; ++t) {
      m1[t] = drop ? drop_scale(seed, s0, (uint32_t)(r0 + t) * C + lane, thresh, inv_keep) : 1.0f;
      x1[t] = x[t] + (pr[0][t] + bp) * m1[t];
    }
    ln_stats(x1, mean, rstd, p.eps);
#pragma unroll
    for (int t = 0; t < T; ++t) {
      xh[t] = (x1[t] - mean[t]) * rstd[t];
      sn2[(tr + t) * C + lane] = (r0 + t < p.R) ? xh[t] * g + bt : 0.f;
    }
    __syncwarp();
    float h[2][T];
    gemv<C, MLP, C>(sn2 + tr * C, W1, lane, h);
    float m2[2][T];  // dropout scale, then dropout scale * gelu'(h)
#pragma unroll
    for (int t = 0; t < T; ++t)
#pragma unroll
      for (int m = 0; m < 2; ++m) {
        const float ms = drop ? drop_scale(seed, s0 + 1, (uint32_t)(r0 + t) * MLP + lane + 32 * m, thresh, inv_keep) : 1.0f;
        float gy, gd;
        gelu_both(h[m][t] + b1[m], gy, gd);
        shd[(tr + t) * MLP + lane + 32 * m] = (r0 + t < p.R) ? gy * ms : 0.f;
        m2[m][t] = ms * gd;
      }
    // ---- backward through fc2
#pragma unroll
    for (int t = 0; t < T; ++t) {
      const float m3 = drop ? drop_scale(seed, s0 + 2, (uint32_t)(r0 + t) * C + lane, thresh, inv_keep) : 1.0f;
      const float d = dy[t] * m3;
      sdo[(tr + t) * C + lane] = d;
      db2 += d;
    }
    __syncwarp();
    float dh[2][T];
    gemv<C, MLP, C>(sdo + tr * C, W2b, lane, dh);
#pragma unroll
    for (int t = 0; t < T; ++t)
#pragma unroll
      for (int m = 0; m < 2; ++m) {
        const float d = dh[m][t] * m2[m][t];
        sdh[(tr + t) * MLP + lane + 32 * m] = d;
        db1[m] += d;
      }
    __syncwarp();
    // ---- backward through fc1 and LayerNorm2
    float dn[1][T];
    gemv<MLP, C, MLP>(sdh + tr * MLP, W1b, lane, dn);
    float dx1[T];
#pragma unroll
    for (int t = 0; t < T; ++t) {
      const float d = dn[0][t];
      dgam += d * xh[t];
      dbet += d;
      const float dxh = d * g;
      const float a1 = warp_sum(dxh) * (1.0f / C), a2 = warp_sum(dxh * xh[t]) * (1.0f / C);
      dx1[t] = dy[t] + rstd[t] * (dxh - a1 - xh[t] * a2);
      const float dp = dx1[t] * m1[t];
      sdp[(tr + t) * C + lane] = dp;
      dbp += dp;
      if (r0 + t < p.R) p.dx1[(size_t)(r0 + t) * C + lane] = dx1[t];
    }
    __syncwarp();
    // ---- backward through the projection
    float da[1][T];
    gemv<C, C, C>(sdp + tr * C, Wpb, lane, da);
#pragma unroll
    for (int t = 0; t < T; ++t)
      if (r0 + t < p.R) p.da[(size_t)(r0 + t) * C + lane] = da[0][t];
    __syncthreads();
    // ---- weight gradients over the CTA's 64-token tile
#pragma unroll 2
    for (int t = 0; t < TILE; ++t) {
      {
        const float d = sdp[t * C + po];
        const ulonglong2 v = *reinterpret_cast<const ulonglong2*>(sa + t * C + pi);
        const u64 dd = pk(d, d);
        ffma2v(dwp[0], dd, v.x); ffma2v(dwp[1], dd, v.y);
      }
      {
        const float d = sdh[t * MLP + o1];
        const ulonglong2 v = *reinterpret_cast<const ulonglong2*>(sn2 + t * C + i1);
        const ulonglong2 w = *reinterpret_cast<const ulonglong2*>(sn2 + t * C + i1 + 4);
        const u64 dd = pk(d, d);
        ffma2v(dw1[0], dd, v.x); ffma2v(dw1[1], dd, v.y); ffma2v(dw1[2], dd, w.x); ffma2v(dw1[3], dd, w.y);
      }
      {
        const float d = sdo[t * C + o2];
        const ulonglong2 v = *reinterpret_cast<const ulonglong2*>(shd + t * MLP + i2);
        const ulonglong2 w = *reinterpret_cast<const ulonglong2*>(shd + t * MLP + i2 + 4);
        const u64 dd = pk(d, d);
        ffma2v(dw2[0], dd, v.x); ffma2v(dw2[1], dd, v.y); ffma2v(dw2[2], dd, w.x); ffma2v(dw2[3], dd, w.y);
      }
    }
    __syncthreads();
  }
  float* part = p.part + (size_t)blockIdx.x * P_MLP;
  *reinterpret_cast<ulonglong2*>(part + O_PROJ + po * C + pi) = make_ulonglong2(dwp[0], dwp[1]);
  *reinterpret_cast<ulonglong2*>(part + O_W1 + o1 * C + i1) = make_ulonglong2(dw1[0], dw1[1]);
  *reinterpret_cast<ulonglong2*>(part + O_W1 + o1 * C + i1 + 4) = make_ulonglong2(dw1[2], dw1[3]);
  *reinterpret_cast<ulonglong2*>(part + O_W2 + o2 * MLP + i2) = make_ulonglong2(dw2[0], dw2[1]);
  *reinterpret_cast<ulonglong2*>(part + O_W2 + o2 * MLP + i2 + 4) = make_ulonglong2(dw2[2], dw2[3]);
  float* rw = red + warp * 192;  // dbp | dgam | dbet | db1 (64) | db2
  rw[lane] = dbp; rw[32 + lane] = dgam; rw[64 + lane] = dbet; rw[96 + lane] = db1[0]; rw[128 + lane] = db1[1]; rw[160 + lane] = db2;
  __syncthreads();
  if (tid < 192) {
    float s = 0.f;
    for (int w = 0; w < WARPS; ++w) s += red[w * 192 + tid];
    const int dst = tid < 32 ? O_BPROJ + tid : tid < 64 ? O_LNW + tid - 32 : tid < 96 ? O_LNB + tid - 64
                    : tid < 160 ? O_B1 + tid - 96 : O_B2 + tid - 160;
    part[dst] = s;
  }
}

// out[k] = sum over CTAs of part[cta][k] in a fixed order: 32 entries per CTA, 32 slices of the partial list per entry.
// With `dst` (one pointer per parameter segment, segment s = entries [seg[s], seg[s+1])) the sums are ADDED to the
// parameters' gradient buffers instead (fused gradient accumulation); a null segment pointer is skipped.
struct SumArgs {
  const float* part; float* out; int n, P, nseg;
  float* dst[8]; int seg[9];
};
__global__ void __launch_bounds__(1024) sum_partials_kernel(const SumArgs a) {
  // 32 entries x 32 slices of the partial list per CTA: every thread has at most ceil(n / 32) independent loads in flight
  // (n <= 296 partial rows), the slices are combined in a fixed order
  __shared__ float sh[32][33];
  const int kx = threadIdx.x & 31, sl = threadIdx.x >> 5, k = blockIdx.x * 32 + kx;
  float s0 = 0.f, s1 = 0.f;
  if (k < a.P) {
    int c = sl;
    for (; c + 32 < a.n; c += 64) {
      s0 += a.part[(size_t)c * a.P + k];
      s1 += a.part[(size_t)(c + 32) * a.P + k];
    }
    if (c < a.n) s0 += a.part[(size_t)c * a.P + k];
  }
  sh[sl][kx] = s0 + s1;
  __syncthreads();
  if (sl == 0 && k < a.P) {
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 32; ++i) s += sh[i][kx];
    if (a.nseg == 0) {
      a.out[k] = s;
    } else {
      for (int g = 0; g < a.nseg; ++g)
        if (k >= a.seg[g] && k < a.seg[g + 1] && a.dst[g] != nullptr) a.dst[g][k - a.seg[g]] += s;
    }
  }
}

int launch_sum(const float* part, float* out, int n, int P, float* const* dst, const int* seg, int nseg, cudaStream_t st) {
  SumArgs a{};
  a.part = part; a.out = out; a.n = n; a.P = P; a.nseg = dst ? nseg : 0;
  if (dst) {
    for (int g = 0; g < nseg; ++g) a.dst[g] = dst[g];
    for (int g = 0; g <= nseg; ++g) a.seg[g] = seg[g];
  }
  sum_partials_kernel<<<(P + 31) / 32, 1024, 0, st>>>(a);
  ERV_LAUNCH_CHECK();
  return ERV_OK;
}

static int grid_for(int R, int rows_per_cta_iter) {
  const int units = (R + rows_per_cta_iter - 1) / rows_per_cta_iter;
  const int cap = 2 * kNumSMs;
  return units < cap ? (units < 1 ? 1 : units) : cap;
}

}  // namespace blk
}  // namespace erv

namespace erv {
namespace blk {
bool mlp_bwd_tc_enabled();
void set_block_tc(int v);
int launch_mlp_bwd_tc(const MlpArgs& a, int max_ctas, cudaStream_t st, int* grid_out);
int launch_mlp_fwd_tc(const MlpArgs& a, cudaStream_t st);
int launch_ln_qkv_tc(bool bwd, const float* x, const float* ln_w, const float* ln_b, const float* w, const float* b, float* qkv,
                     const float* dqkv, const float* dres, float* dx, float* part, int rows, float eps, int max_ctas,
                     cudaStream_t st, int* grid_out, int act_bf16);
}  // namespace blk
}  // namespace erv

using namespace erv;
using namespace erv::blk;

extern "C" int erv_block_supported(int dim, int mlp_dim) { return dim == C && mlp_dim == MLP; }
extern "C" void erv_block_set_tensor_core(int mode) { set_block_tc(mode); }
extern "C" int erv_block_ln_qkv_params(void) { return P_QKV; }
extern "C" int erv_block_mlp_params(void) { return P_MLP; }
extern "C" size_t erv_block_ln_qkv_bwd_workspace(int rows) { return align_up((size_t)grid_for(rows, TILE) * P_QKV * sizeof(float), 256); }
extern "C" size_t erv_block_mlp_bwd_workspace(int rows) { return align_up((size_t)grid_for(rows, TILE) * P_MLP * sizeof(float), 256); }

static bool aligned16(std::initializer_list<const void*> ptrs) {
  for (const void* q : ptrs)
    if (reinterpret_cast<uintptr_t>(q) & 15) return false;
  return true;
}

// act_dtype: ERV_F32 or ERV_BF16 (tcgen05 family only) = element type of the activations exchanged with the attention core
static int check_act(const char* fn, int act_dtype) {
  if (act_dtype == ERV_F32) return ERV_OK;
  if (act_dtype == ERV_BF16 && mlp_bwd_tc_enabled()) return ERV_OK;
  set_error("%s: activation dtype %d unsupported (bf16 needs the tensor-core block kernels)", fn, act_dtype);
  return ERV_E_UNSUPPORTED;
}
extern "C" int erv_block_act_bf16_supported(void) { return mlp_bwd_tc_enabled() ? 1 : 0; }

extern "C" int erv_block_ln_qkv_fwd(const float* x, const float* ln_w, const float* ln_b, const float* w_qkv,
                                    const float* b_qkv, void* qkv_out, int act_dtype, int rows, int dim, float eps,
                                    void* stream) {
  float* qkv = static_cast<float*>(qkv_out);
  ERV_CHECK_ARG(x && ln_w && ln_b && w_qkv && qkv && rows > 0, "erv_block_ln_qkv_fwd: bad arguments");
  if (int rc = check_act("erv_block_ln_qkv_fwd", act_dtype)) return rc;
  ERV_CHECK_ARG(aligned16({x, ln_w, ln_b, w_qkv, b_qkv, qkv}), "erv_block_ln_qkv_fwd: pointers must be 16-byte aligned");
  if (dim != C) { set_error("erv_block_ln_qkv_fwd: dim %d not supported (32)", dim); return ERV_E_UNSUPPORTED; }
  if (mlp_bwd_tc_enabled())  // tcgen05 tiles (erv_block_tc.cu)
    return launch_ln_qkv_tc(false, x, ln_w, ln_b, w_qkv, b_qkv, qkv, nullptr, nullptr, nullptr, nullptr, rows, eps, 0,
                            (cudaStream_t)stream, nullptr, act_dtype == ERV_BF16);
  LnQkvArgs a{};
  a.x = x; a.ln_w = ln_w; a.ln_b = ln_b; a.w = w_qkv; a.b = b_qkv; a.qkv = qkv; a.R = rows; a.eps = eps;
  const size_t smem = (size_t)(QKV * C + WARPS * T * C) * sizeof(float);
  ERV_CUDA(allow_smem(ln_qkv_fwd_kernel, smem));
  ln_qkv_fwd_kernel<<<grid_for(rows, TILE), THREADS, smem, (cudaStream_t)stream>>>(a);
  ERV_LAUNCH_CHECK();
  return ERV_OK;
}

extern "C" int erv_block_ln_qkv_bwd(const float* x, const void* dqkv_in, int act_dtype, const float* dres, const float* ln_w,
                                    const float* ln_b, const float* w_qkv, float* dx, float* dparams,
                                    float* const* grad_accum, int rows, int dim, float eps, void* workspace,
                                    size_t workspace_bytes, void* stream) {
  const float* dqkv = static_cast<const float*>(dqkv_in);
  if (int rc = check_act("erv_block_ln_qkv_bwd", act_dtype)) return rc;
  ERV_CHECK_ARG(x && dqkv && ln_w && ln_b && w_qkv && dx && (dparams || grad_accum) && workspace && rows > 0,
                "erv_block_ln_qkv_bwd: bad arguments");
  ERV_CHECK_ARG(aligned16({x, dqkv, dres, ln_w, ln_b, w_qkv, dx, workspace}), "erv_block_ln_qkv_bwd: pointers must be 16-byte aligned");
  if (dim != C) { set_error("erv_block_ln_qkv_bwd: dim %d not supported (32)", dim); return ERV_E_UNSUPPORTED; }
  if (workspace_bytes < erv_block_ln_qkv_bwd_workspace(rows)) { set_error("erv_block_ln_qkv_bwd: workspace too small"); return ERV_E_WORKSPACE; }
  const int seg[5] = {0, QKV * C, QKV * C + QKV, QKV * C + QKV + C, P_QKV};
  if (mlp_bwd_tc_enabled()) {  // tcgen05 tiles (erv_block_tc.cu)
    int tc_grid = 0;
    int rc = launch_ln_qkv_tc(true, x, ln_w, ln_b, w_qkv, nullptr, nullptr, dqkv, dres, dx, (float*)workspace, rows, eps,
                              grid_for(rows, TILE), (cudaStream_t)stream, &tc_grid, act_dtype == ERV_BF16);
    if (rc) return rc;
    return launch_sum((const float*)workspace, dparams, tc_grid, P_QKV, grad_accum, seg, 4, (cudaStream_t)stream);
  }
  LnQkvArgs a{};
  a.x = x; a.ln_w = ln_w; a.ln_b = ln_b; a.w = w_qkv; a.dqkv = dqkv; a.dres = dres; a.dx = dx; a.part = (float*)workspace;
  a.R = rows; a.eps = eps;
  const int grid = grid_for(rows, TILE);
  const size_t smem = (size_t)(QKV * C + TILE * C + TILE * QKV + WARPS * (QKV + 2 * C)) * sizeof(float);
  ERV_CUDA(allow_smem(ln_qkv_bwd_kernel, smem));
  cudaStream_t st = (cudaStream_t)stream;
  ln_qkv_bwd_kernel<<<grid, THREADS, smem, st>>>(a);
  ERV_LAUNCH_CHECK();
  return launch_sum((const float*)workspace, dparams, grid, P_QKV, grad_accum, seg, 4, st);
}

static int fill_mlp(MlpArgs& a, const char* fn, const float* attn_out, const float* x, const float* const* params, int rows,
                    int dim, int mlp_dim, float eps, float p_drop, const long long* seed, int salt, int act_dtype) {
  if (int rc = check_act(fn, act_dtype)) return rc;
  a.act_bf16 = act_dtype == ERV_BF16;
  ERV_CHECK_ARG(attn_out && x && params && rows > 0, "%s: bad arguments", fn);
  for (int i = 0; i < 8; ++i) ERV_CHECK_ARG(params[i], "%s: parameter %d is null", fn, i);
  for (int i = 0; i < 8; ++i) ERV_CHECK_ARG(aligned16({params[i]}), "%s: parameter %d must be 16-byte aligned", fn, i);
  ERV_CHECK_ARG(aligned16({attn_out, x}), "%s: pointers must be 16-byte aligned", fn);
  ERV_CHECK_ARG(p_drop >= 0.f && p_drop < 1.f && (p_drop == 0.f || seed), "%s: dropout %g needs 0 <= p < 1 and a seed", fn, (double)p_drop);
  if (dim != C || mlp_dim != MLP) { set_error("%s: dims (%d, %d) not supported (32, 64)", fn, dim, mlp_dim); return ERV_E_UNSUPPORTED; }
  a.a = attn_out; a.x = x;
  a.w_proj = params[0]; a.b_proj = params[1]; a.ln_w = params[2]; a.ln_b = params[3];
  a.w1 = params[4]; a.b1 = params[5]; a.w2 = params[6]; a.b2 = params[7];
  a.seed = seed; a.salt = salt; a.R = rows; a.eps = eps; a.p_drop = p_drop;
  return ERV_OK;
}

extern "C" int erv_block_mlp_fwd(const void* attn_out, int act_dtype, const float* x, const float* const* params, float* y,
                                 int rows, int dim, int mlp_dim, float eps, float p_drop, const long long* seed, int salt,
                                 void* stream) {
  MlpArgs a{};
  int rc = fill_mlp(a, "erv_block_mlp_fwd", static_cast<const float*>(attn_out), x, params, rows, dim, mlp_dim, eps, p_drop, seed,
                    salt, act_dtype);
  if (rc) return rc;
  ERV_CHECK_ARG(y && aligned16({y}), "erv_block_mlp_fwd: null or misaligned output");
  a.y = y;
  if (mlp_bwd_tc_enabled()) return launch_mlp_fwd_tc(a, (cudaStream_t)stream);  // tcgen05 tiles (erv_block_tc.cu)
  const size_t smem = (size_t)(C * C + 2 * MLP * C + WARPS * T * MLP) * sizeof(float);
  ERV_CUDA(allow_smem(mlp_fwd_kernel, smem));
  mlp_fwd_kernel<<<grid_for(rows, TILE), THREADS, smem, (cudaStream_t)stream>>>(a);
  ERV_LAUNCH_CHECK();
  return ERV_OK;
}

extern "C" int erv_block_mlp_bwd(const void* attn_out, int act_dtype, const float* x, const float* dy,
                                 const float* const* params, void* d_attn_out_v, float* dx1, float* dparams,
                                 float* const* grad_accum, int rows, int dim, int mlp_dim, float eps, float p_drop,
                                 const long long* seed, int salt, void* workspace, size_t workspace_bytes, void* stream) {
  MlpArgs a{};
  float* d_attn_out = static_cast<float*>(d_attn_out_v);
  int rc = fill_mlp(a, "erv_block_mlp_bwd", static_cast<const float*>(attn_out), x, params, rows, dim, mlp_dim, eps, p_drop, seed,
                    salt, act_dtype);
  if (rc) return rc;
  ERV_CHECK_ARG(dy && d_attn_out && dx1 && (dparams || grad_accum) && workspace, "erv_block_mlp_bwd: null pointer");
  ERV_CHECK_ARG(aligned16({dy, d_attn_out, dx1, workspace}), "erv_block_mlp_bwd: pointers must be 16-byte aligned");
  if (workspace_bytes < erv_block_mlp_bwd_workspace(rows)) { set_error("erv_block_mlp_bwd: workspace too small"); return ERV_E_WORKSPACE; }
  a.dy = dy; a.da = d_attn_out; a.dx1 = dx1; a.part = (float*)workspace;
  const int seg[9] = {O_PROJ, O_BPROJ, O_LNW, O_LNB, O_W1, O_B1, O_W2, O_B2, P_MLP};
  if (mlp_bwd_tc_enabled()) {  // tcgen05 tiles (erv_block_tc.cu)
    int tc_grid = 0;
    rc = launch_mlp_bwd_tc(a, grid_for(rows, TILE), (cudaStream_t)stream, &tc_grid);
    if (rc) return rc;
    return launch_sum((const float*)workspace, dparams, tc_grid, P_MLP, grad_accum, seg, 8, (cudaStream_t)stream);
  }
  const int grid = grid_for(rows, TILE);
  const size_t smem = (size_t)(2 * C * C + 3 * MLP * C + 4 * TILE * C + 2 * TILE * MLP + WARPS * 192) * sizeof(float);
  ERV_CUDA(allow_smem(mlp_bwd_kernel, smem));
  cudaStream_t st = (cudaStream_t)stream;
  mlp_bwd_kernel<<<grid, THREADS, smem, st>>>(a);
  ERV_LAUNCH_CHECK();
  return launch_sum((const float*)workspace, dparams, grid, P_MLP, grad_accum, seg, 8, st);
}
