// Shared by the fused block kernels (erv_block_fused.cu: FFMA2 register tiles; erv_block_tc.cu: tcgen05 tiles).
#pragma once
#include "erv_common.cuh"

namespace erv {
namespace blk {

constexpr int C = 32, QKV = 96, MLP = 64, T = 8, WARPS = 8, THREADS = 256, TILE = T * WARPS;
constexpr int P_QKV = QKV * C + QKV + C + C;                              // dW_qkv | db_qkv | dln_w | dln_b
constexpr int P_MLP = C * C + C + C + C + MLP * C + MLP + C * MLP + C;   // dW_proj | db_proj | dln_w | dln_b | dW1 | db1 | dW2 | db2
constexpr int O_PROJ = 0, O_BPROJ = C * C, O_LNW = O_BPROJ + C, O_LNB = O_LNW + C, O_W1 = O_LNB + C,
              O_B1 = O_W1 + MLP * C, O_W2 = O_B1 + MLP, O_B2 = O_W2 + C * MLP;

// counter-based dropout: keep-scale of element idx of stream `stream` (1/(1-p) or 0)
__device__ __forceinline__ float drop_scale(unsigned long long seed, uint32_t stream, uint32_t idx, uint32_t thresh, float inv_keep) {
  uint32_t h = idx * 0x9E3779B1u + (uint32_t)seed;
  h ^= h >> 16; h *= 0x85EBCA6Bu;
  h += stream * 0xC2B2AE35u + (uint32_t)(seed >> 32);
  h ^= h >> 13; h *= 0xC2B2AE35u;
  h ^= h >> 16; h *= 0x27D4EB2Fu;
  h ^= h >> 15;
  return h >= thresh ? inv_keep : 0.f;
}

__device__ __forceinline__ float gelu_f(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752f)); }
// gelu(x) and gelu'(x) from one erf evaluation
__device__ __forceinline__ void gelu_both(float x, float& y, float& dy) {
  const float cdf = 0.5f * (1.0f + erff(x * 0.70710678118654752f));
  y = x * cdf;
  dy = cdf + x * 0.3989422804014327f * __expf(-0.5f * x * x);
}

struct MlpArgs {
  const float* a; const float* x;  // attention output (pre-projection) and block input, [R][C]
  const float* w_proj; const float* b_proj; const float* ln_w; const float* ln_b;
  const float* w1; const float* b1; const float* w2; const float* b2;
  float* y;                         // fwd out
  const float* dy;                  // bwd in
  float* da; float* dx1; float* part;  // bwd out
  const long long* seed; int salt;
  int R; float eps, p_drop;
  int act_bf16;                     // a / da are bf16 (autocast) instead of fp32: tcgen05 family only
};

// out[k] = sum over CTAs of part[cta][k] (fixed order); with dst the sums are added to the segment buffers instead
int launch_sum(const float* part, float* out, int n, int P, float* const* dst, const int* seg, int nseg, cudaStream_t st);

}  // namespace blk
}  // namespace erv
