// Shared device/host helpers for the erv_b200 kernels (sm_100a only).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include <atomic>

#include "../../include/erv_b200.h"

namespace erv {

constexpr int kThreads = 256;      // default CTA size for the tile kernels
constexpr float kEps = 1e-6f;      // favor_plus.py:260
constexpr int kNumSMs = 148;       // B200

// ---- host-side status plumbing ---------------------------------------------------------------
void set_error(const char* fmt, ...);
extern std::atomic<uint64_t> g_launches;
inline void count_launch(int n = 1) { g_launches.fetch_add((uint64_t)n, std::memory_order_relaxed); }

#define ERV_CHECK_ARG(cond, ...)            \
  do {                                      \
    if (!(cond)) {                          \
      erv::set_error(__VA_ARGS__);          \
      return ERV_E_INVALID;                 \
    }                                       \
  } while (0)

#define ERV_CUDA(call)                                                              \
  do {                                                                              \
    cudaError_t e_ = (call);                                                        \
    if (e_ != cudaSuccess) {                                                        \
      erv::set_error("%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
      return ERV_E_CUDA;                                                            \
    }                                                                               \
  } while (0)

#define ERV_LAUNCH_CHECK()                                                          \
  do {                                                                              \
    erv::count_launch();                                                            \
    cudaError_t e_ = cudaGetLastError();                                            \
    if (e_ != cudaSuccess) {                                                        \
      erv::set_error("kernel launch failed: %s (%s:%d)", cudaGetErrorString(e_), __FILE__, __LINE__); \
      return ERV_E_CUDA;                                                            \
    }                                                                               \
  } while (0)

inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// Opt a kernel in to > 48 KB of dynamic shared memory.  The attribute is set once per (function, device) and
// only raised, so steady-state launches (and CUDA-graph capture) make no driver call here.
cudaError_t ensure_smem(const void* fn, size_t bytes);
template <typename K>
inline cudaError_t allow_smem(K kernel, size_t bytes) {
  return ensure_smem((const void*)kernel, bytes);
}
constexpr size_t kMaxSmem = 227 * 1024;

// ---- element access ----------------------------------------------------------------------------
__device__ __forceinline__ float to_f(float v) { return v; }
__device__ __forceinline__ float to_f(__nv_bfloat16 v) { return __bfloat162float(v); }

// load 4 consecutive elements (16-byte aligned for fp32, 8-byte for bf16) as float4
__device__ __forceinline__ float4 ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ float4 ld4(const __nv_bfloat16* p) {
  uint2 raw = *reinterpret_cast<const uint2*>(p);
  __nv_bfloat162 a = *reinterpret_cast<__nv_bfloat162*>(&raw.x);
  __nv_bfloat162 b = *reinterpret_cast<__nv_bfloat162*>(&raw.y);
  float2 fa = __bfloat1622float2(a), fb = __bfloat1622float2(b);
  return make_float4(fa.x, fa.y, fb.x, fb.y);
}
__device__ __forceinline__ void st4(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }
__device__ __forceinline__ void st4(__nv_bfloat16* p, float4 v) {
  __nv_bfloat162 a = __floats2bfloat162_rn(v.x, v.y), b = __floats2bfloat162_rn(v.z, v.w);
  uint2 raw;
  raw.x = *reinterpret_cast<uint32_t*>(&a);
  raw.y = *reinterpret_cast<uint32_t*>(&b);
  *reinterpret_cast<uint2*>(p) = raw;
}

__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// packed qkv addressing: element (b, n, which, h, 0) of a [B, N, 3, H, Dh] buffer
__device__ __forceinline__ size_t qkv_off(int b, int n, int which, int h, int N, int H, int DH) {
  return ((((size_t)b * N + n) * 3 + which) * H + h) * DH;
}
// element (b, n, h, 0) of a [B, N, H, Dh] buffer
__device__ __forceinline__ size_t out_off(int b, int n, int h, int N, int H, int DH) {
  return (((size_t)b * N + n) * H + h) * DH;
}

}  // namespace erv
