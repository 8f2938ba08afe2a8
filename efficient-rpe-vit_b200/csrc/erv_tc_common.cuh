// Building blocks shared by the tensor-core linear-attention kernels (forward: erv_linattn_tc.cu, backward:
// erv_linattn_tc_bwd.cu): argument block, operand-image geometry, register-level prologue, bf16 hi/lo splitting.
#pragma once
#include <cuda_bf16.h>

#include "erv_common.cuh"
#include "erv_umma.cuh"

namespace erv {
using namespace umma;

struct LaTcArgs {
  const void* qkv;
  void* out;
  const float* omega;  // [H][DH][M]
  const float* ta;
  const float* tb;
  int B, N, H, M, Mp16, kind, rot;
  float prescale, inv_sqrt_m;
  float* state;  // optional [B*H][DH+1][Mp]: the finished [S|z] of every pair, saved for the backward (short-sequence kernel)
  long long* trace = nullptr;  // optional phase trace (erv_debug_set_trace, builds with -DERV_TRACE only)
};

template <int DH>
struct TcCfg {
  static constexpr int ND = (DH + 1 + 15) / 16 * 16;        // columns of [v|1] and [S|z], padded for the MMA N dim
  static constexpr uint32_t X_LBO = 128, X_SBO = (DH / 4) * 128;  // x / w images: rows x Dh
  static constexpr uint32_t T_LBO = 144, T_SBO = 32 * 144;  // images whose K index is the 128 tokens (padded chunks)
  static constexpr uint32_t P_LBO = 128, P_SBO = 32 * 128;  // phi image [128 tokens x 128 features]
  static constexpr uint32_t X_BYTES = 16 * X_SBO;           // 128 rows
  static constexpr uint32_t PHI_BYTES = 16 * T_SBO;         // >= 16 * P_SBO
  static constexpr uint32_t V_BYTES = (ND / 8) * T_SBO;     // one [ND x 128 tokens] image
  static constexpr int COL_S = 256, COL_O = 256 + 2 * ND;   // TMEM columns
};

__host__ __device__ inline uint32_t tc_w_bytes(int DH, int Mp16) { return (uint32_t)(Mp16 / 8) * (DH / 4) * 128; }
__host__ __device__ inline uint32_t tc_s_bytes(int ND, int Mp16) { return (uint32_t)(ND / 8) * (Mp16 / 4) * 144; }

template <typename T, int DH>
__device__ __forceinline__ void load_row(const T* __restrict__ p, float (&x)[DH]) {
#pragma unroll
  for (int c = 0; c < DH / 4; ++c) {
    float4 v = ld4(p + 4 * c);
    x[4 * c] = v.x; x[4 * c + 1] = v.y; x[4 * c + 2] = v.z; x[4 * c + 3] = v.w;
  }
}

// RoPE / Circulant-STRING rotation of one token row.  Deliberately out of line: the rotations are 0.3-0.6 k instructions per call
// site, the short-sequence kernels inline their caller at three to four sites, and those kernels are instruction-fetch bound
// (profiles/r01_final_ncu_attention.md: no_inst is their top stall reason), so the un-rotated path should not have to jump over
// that code.  The row goes through local memory only on the rotated path.
template <int DH>
__device__ __forceinline__ void rotate_row_impl(float* __restrict__ x, int rot, const float* __restrict__ ta,
                                        const float* __restrict__ tb, int h, int n, int N) {
  if (rot == ERV_ROT_ROPE) {
#pragma unroll
    for (int m = 0; m < DH / 2; ++m) {
      const float c = __ldg(ta + (size_t)n * (DH / 2) + m), s = __ldg(tb + (size_t)n * (DH / 2) + m);
      const float xe = x[2 * m], xo = x[2 * m + 1];
      x[2 * m] = xe * c - xo * s;
      x[2 * m + 1] = xe * s + xo * c;
    }
  } else if (rot == ERV_ROT_CIRCULANT) {
    float g[DH], xin[DH];
    load_row<float, DH>(ta + ((size_t)h * N + n) * DH, g);
#pragma unroll
    for (int a = 0; a < DH; ++a) xin[a] = x[a];
#pragma unroll
    for (int a = 0; a < DH; ++a) {
      float acc = 0.f;
#pragma unroll
      for (int b = 0; b < DH; ++b) acc = fmaf(g[(a - b) & (DH - 1)], xin[b], acc);
      x[a] = acc;
    }
  }
}

template <int DH>
__device__ __noinline__ void rotate_row(float* __restrict__ x, int rot, const float* __restrict__ ta,
                                        const float* __restrict__ tb, int h, int n, int N) {
  rotate_row_impl<DH>(x, rot, ta, tb, h, n, N);
}

// rotation (RoPE / Circulant-STRING) + Dh^-1/4 scale of one token row held in registers
// INLINE_ROT: the long-sequence kernels (one call site, rotation on their hot path at BASELINE config 4) keep it inline.
template <int DH, bool INLINE_ROT = false>
__device__ __forceinline__ void prologue_row(float (&x)[DH], int rot, const float* __restrict__ ta,
                                             const float* __restrict__ tb, int h, int n, int N, float prescale) {
  if (INLINE_ROT) {
    rotate_row_impl<DH>(x, rot, ta, tb, h, n, N);
  } else if (rot != ERV_ROT_NONE) {
    float t[DH];
#pragma unroll
    for (int a = 0; a < DH; ++a) t[a] = x[a];
    rotate_row<DH>(t, rot, ta, tb, h, n, N);
#pragma unroll
    for (int a = 0; a < DH; ++a) x[a] = t[a];
  }
#pragma unroll
  for (int a = 0; a < DH; ++a) x[a] *= prescale;
}

// inverse of the rotation for a gradient row dy: rotate back, and for the Circulant rotation accumulate this token's share of
// dL/dg into the CTA-private slot.  Out of line for the same reason as rotate_row (erv_tc_common.cuh).
template <typename T, int DH>
__device__ __noinline__ void rotate_row_bwd(const float* __restrict__ dy, float* __restrict__ dxr, int rot,
                                            const float* __restrict__ ta, const float* __restrict__ tb, int h, int n, int N,
                                            float* dg_slot, const T* x_raw_row) {
  if (rot == ERV_ROT_ROPE) {
#pragma unroll
    for (int m = 0; m < DH / 2; ++m) {
      const float c = __ldg(ta + (size_t)n * (DH / 2) + m), s = __ldg(tb + (size_t)n * (DH / 2) + m);
      dxr[2 * m] = dy[2 * m] * c + dy[2 * m + 1] * s;
      dxr[2 * m + 1] = dy[2 * m + 1] * c - dy[2 * m] * s;
    }
  } else {
    float g[DH], d[DH];
    load_row<float, DH>(ta + ((size_t)h * N + n) * DH, g);
#pragma unroll
    for (int aa = 0; aa < DH; ++aa) d[aa] = dy[aa];
#pragma unroll
    for (int bq = 0; bq < DH; ++bq) {
      float a = 0.f;
#pragma unroll
      for (int aa = 0; aa < DH; ++aa) a = fmaf(g[(aa - bq) & (DH - 1)], d[aa], a);
      dxr[bq] = a;
    }
    if (dg_slot != nullptr && n >= 1) {
      float xr[DH];
      load_row<T, DH>(x_raw_row, xr);
#pragma unroll
      for (int m = 0; m < DH; ++m) {
        float a = 0.f;
#pragma unroll
        for (int aa = 0; aa < DH; ++aa) a = fmaf(d[aa], xr[(aa - m) & (DH - 1)], a);
        dg_slot[(size_t)n * DH + m] += a;  // slot private to this (CTA, pair side), row private to this thread
      }
    }
  }
}

// inverse of prologue_row for a gradient row dy (already multiplied by the Dh^-1/4 scale)
template <typename T, int DH>
__device__ __forceinline__ void prologue_row_bwd(const float (&dy)[DH], float (&dxr)[DH], int rot, const float* ta,
                                                 const float* tb, int h, int n, int N, float* dg_slot,
                                                 const T* x_raw_row) {
  if (rot == ERV_ROT_ROPE || rot == ERV_ROT_CIRCULANT) {
    float tin[DH], tout[DH];
#pragma unroll
    for (int d = 0; d < DH; ++d) tin[d] = dy[d];
    rotate_row_bwd<T, DH>(tin, tout, rot, ta, tb, h, n, N, dg_slot, x_raw_row);
#pragma unroll
    for (int d = 0; d < DH; ++d) dxr[d] = tout[d];
  } else {
#pragma unroll
    for (int d = 0; d < DH; ++d) dxr[d] = dy[d];
  }
}

// write one token row as the hi / lo TF32 images of a K-major [128 x DH] operand
template <int DH>
__device__ __forceinline__ void store_x_images(uint8_t* xh, uint8_t* xl, const float (&x)[DH], int t) {
  using C = TcCfg<DH>;
#pragma unroll
  for (int c = 0; c < DH / 4; ++c) {
    float hi[4], lo[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      hi[e] = to_tf32(x[4 * c + e]);
      lo[e] = to_tf32(x[4 * c + e] - hi[e]);
    }
    const uint32_t off = (uint32_t)(t >> 3) * C::X_SBO + c * C::X_LBO + (t & 7) * 16;
    st4(reinterpret_cast<float*>(xh + off), make_float4(hi[0], hi[1], hi[2], hi[3]));
    st4(reinterpret_cast<float*>(xl + off), make_float4(lo[0], lo[1], lo[2], lo[3]));
  }
}

// ---- bf16 hi/lo splitting ---------------------------------------------------------------------------------------
// hi = bf16(v) and lo = bf16(v - hi), both round-to-nearest, two values per 32-bit word: hi + lo carries >= 16 significant bits
// (|v - hi| <= 2^-9 |v| is exact in fp32, its own rounding error is <= 2^-18 |v|).  Six instructions per pair of values:
// two F2FP.BF16.F32.PACK_AB, a shift, a mask and two subtractions.  F2FP runs at 32 lanes/clk/SM on a pipe of its own: it does
// not compete with MUFU.EX2 (profiles/r02_tcgen05_mma_cost.md), unlike the round-1 integer version (eight ALU instructions).
__device__ __forceinline__ uint32_t pack_bf16x2_rn(float lo_half, float hi_half) {
  uint32_t r;
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi_half), "f"(lo_half));
  return r;
}
__device__ __forceinline__ void split_pack2(float a, float b, uint32_t& hi, uint32_t& lo) {
  hi = pack_bf16x2_rn(a, b);
  const float ra = a - __uint_as_float(hi << 16), rb = b - __uint_as_float(hi & 0xffff0000u);
  lo = pack_bf16x2_rn(ra, rb);
}
__device__ __forceinline__ void store_split8(uint8_t* img_hi, uint8_t* img_lo, uint32_t off, const float (&v)[8]) {
  uint4 h, l;
  split_pack2(v[0], v[1], h.x, l.x);
  split_pack2(v[2], v[3], h.y, l.y);
  split_pack2(v[4], v[5], h.z, l.z);
  split_pack2(v[6], v[7], h.w, l.w);
  *reinterpret_cast<uint4*>(img_hi + off) = h;
  *reinterpret_cast<uint4*>(img_lo + off) = l;
}
__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// Shared-memory images (bytes).  bf16 images use 16-byte chunks of 8 elements along the fast index f (or d) and 128
// contiguous bytes per group of 8 tokens:  byte(t, f) = (f/8)*CH + (t/8)*128 + (t%8)*16 + (f%8)*2.
// Read as an MN-major operand (rows f, K = t): SBO = CH, LBO = 128; as a K-major operand (rows t, K = f): SBO = 128,
// LBO = CH.  The same image therefore feeds phi^T [v|1] and phi [S|z].
// Sums of 32 per-lane values over the warp in 31 shuffles (a plain butterfly per value needs 160): each step halves the
// values a lane still carries.  Afterwards lane l holds, in the return value, the warp total of the value with index l.
__device__ __forceinline__ float warp_sum32(float (&v)[32]) {
  const int lane = threadIdx.x & 31;
#pragma unroll
  for (int o = 16; o >= 1; o >>= 1) {
    const bool up = (lane & o) != 0;
#pragma unroll
    for (int i = 0; i < o; ++i) {
      const float send = up ? v[i] : v[i + o];
      const float keep = up ? v[i + o] : v[i];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, o);
    }
  }
  return v[0];
}

// 16-byte asynchronous global -> shared copies (LDGSTS): no registers are held while the data is in flight
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(smem_u32(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;\n" ::: "memory"); }

// packed fp32 pair FMA (FFMA2 on sm_100): acc.{lo,hi} += g * w.{lo,hi}; one issue slot for two FMAs
__device__ __forceinline__ unsigned long long pack2(float lo, float hi) {
  return ((unsigned long long)__float_as_uint(hi) << 32) | (unsigned long long)__float_as_uint(lo);
}
__device__ __forceinline__ void ffma2(unsigned long long& acc, float g, float w_lo, float w_hi) {
  asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(acc) : "l"(pack2(g, g)), "l"(pack2(w_lo, w_hi)));
}
__device__ __forceinline__ float lo_of(unsigned long long v) { return __uint_as_float((uint32_t)v); }
__device__ __forceinline__ float hi_of(unsigned long long v) { return __uint_as_float((uint32_t)(v >> 32)); }

constexpr uint32_t kTokCh = 16 * 128;  // chunk stride of images with 128 token rows
constexpr int kTcThreads = 512;        // 4 threads per token row, each owns a quarter of the features

}  // namespace erv
