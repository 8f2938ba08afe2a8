// Pieces shared by the tcgen05 tile kernels (erv_ktile_tc.cu: KERPLE, erv_stile_tc.cu: softmax): tile constants and the
// register-row -> operand-image stores.  Tiles are 128 tokens of the flattened (batch, token) axis of one head; 512 threads,
// thread (row, quarter) owns 32 of the 128 columns of a score tile.
#pragma once
#include "erv_tc_common.cuh"

namespace erv {

constexpr int KT = 128;              // tile rows
constexpr int KTHREADS = 512;
constexpr uint32_t VI_CH = 16 * 128;          // MN-major [K = 128 tokens][N] image: 8-column chunks 2048 B apart
constexpr uint32_t XD_SBO = 256;              // K-major [128 x 16] bf16 image (q, k, dO, v): 8-row groups 256 B apart
constexpr uint32_t XD_BYTES = 16 * XD_SBO;    // 4 KB per level

// one token row (registers) -> K-major bf16 hi/lo image [128 x 16]
__device__ __forceinline__ void kt_store_row_kmajor(uint8_t* img, int row, const float (&x)[16]) {
#pragma unroll
  for (int cc = 0; cc < 2; ++cc) {
    float v[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) v[e] = x[8 * cc + e];
    store_split8(img, img + XD_BYTES, (uint32_t)(row >> 3) * XD_SBO + (uint32_t)cc * 128 + (row & 7) * 16, v);
  }
}
// one token row -> MN-major image [K = token][N = 48]: chunks 0,1 hi | 2 = [extra,0..] | 3 = 0 | 4,5 lo
__device__ __forceinline__ void kt_store_row_mnmajor(uint8_t* img, int row, const float (&x)[16], float extra) {
  const uint32_t off = (uint32_t)(row >> 3) * 128 + (row & 7) * 16;
#pragma unroll
  for (int cc = 0; cc < 2; ++cc) {
    float v[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) v[e] = x[8 * cc + e];
    store_split8(img + (uint32_t)cc * VI_CH, img + (uint32_t)(4 + cc) * VI_CH, off, v);
  }
  *reinterpret_cast<uint4*>(img + 2 * VI_CH + off) = make_uint4((uint32_t)__bfloat16_as_ushort(__float2bfloat16_rn(extra)), 0u, 0u, 0u);
  *reinterpret_cast<uint4*>(img + 3 * VI_CH + off) = make_uint4(0u, 0u, 0u, 0u);
}

// number of CTAs along x: packed tiles (floor(128/N) pairs each) or 128-row tiles of single pairs
inline int tile_tc_grid_x(int B, int N) { return N <= KT ? (B + KT / N - 1) / (KT / N) : B * ((N + KT - 1) / KT); }

}  // namespace erv
