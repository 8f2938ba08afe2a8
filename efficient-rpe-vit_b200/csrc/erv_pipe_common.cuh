// Helpers shared by the warp-specialised, pipelined short-sequence linear-attention kernels (erv_linattn_pipe.cu forward,
// erv_linattn_pipe_bwd.cu backward): mbarrier arrivals of the compute warps, 32-bit shared-window loads / stores, bf16 level
// splits, multi-value warp sums.
#pragma once
#include "erv_tc_common.cuh"

namespace erv {

extern long long* g_trace;  // erv_debug_set_trace (erv_linattn_tc_bwd.cu)

constexpr int kPipeThreads = 608;  // 16 compute warps + the MMA-issue warp + 2 lone-token warps

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(smem_u32(bar)) : "memory");
}
// One arrival per compute warp: every lane first orders its shared-memory writes (read by the tensor core through the
// async proxy) and its tensor-memory reads before the arrival.
__device__ __forceinline__ void warp_arrive(uint64_t* bar) {
  fence_smem_to_async();
  fence_before_sync();
  __syncwarp();
  if ((threadIdx.x & 31) == 0) mbar_arrive(bar);
}
__device__ __forceinline__ void bar_compute() { asm volatile("bar.sync 1, 512;\n" ::: "memory"); }

__device__ __forceinline__ void tmem_ld4(uint32_t taddr, float (&v)[4]) {
  uint32_t r[4];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0,%1,%2,%3}, [%4];\n"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(taddr)
               : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;\n" : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3])::"memory");
#pragma unroll
  for (int i = 0; i < 4; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_wait_ld32(uint32_t (&r)[32]) {
  asm volatile("tcgen05.wait::ld.sync.aligned;\n"
               : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]),
                 "+r"(r[8]), "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15]),
                 "+r"(r[16]), "+r"(r[17]), "+r"(r[18]), "+r"(r[19]), "+r"(r[20]), "+r"(r[21]), "+r"(r[22]), "+r"(r[23]),
                 "+r"(r[24]), "+r"(r[25]), "+r"(r[26]), "+r"(r[27]), "+r"(r[28]), "+r"(r[29]), "+r"(r[30]), "+r"(r[31])
               :
               : "memory");
}

// three-level bf16 split of 8 values (24 significant bits), 16-byte stores into three images
__device__ __forceinline__ void store_split8_l3(uint8_t* i0, uint8_t* i1, uint8_t* i2, uint32_t off, const float (&v)[8]) {
  uint32_t a[4], b[4], c[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float x = v[2 * i], y = v[2 * i + 1];
    a[i] = pack_bf16x2_rn(x, y);
    const float rx = x - __uint_as_float(a[i] << 16), ry = y - __uint_as_float(a[i] & 0xffff0000u);
    b[i] = pack_bf16x2_rn(rx, ry);
    const float sx = rx - __uint_as_float(b[i] << 16), sy = ry - __uint_as_float(b[i] & 0xffff0000u);
    c[i] = pack_bf16x2_rn(sx, sy);
  }
  *reinterpret_cast<uint4*>(i0 + off) = make_uint4(a[0], a[1], a[2], a[3]);
  *reinterpret_cast<uint4*>(i1 + off) = make_uint4(b[0], b[1], b[2], b[3]);
  *reinterpret_cast<uint4*>(i2 + off) = make_uint4(c[0], c[1], c[2], c[3]);
}

// sums of 16 per-lane values over the warp in 16 shuffles; afterwards lane l holds the total of value (l >> 1) & 15
__device__ __forceinline__ float warp_sum16(float (&v)[16]) {
  const int lane = threadIdx.x & 31;
#pragma unroll
  for (int o = 16; o >= 2; o >>= 1) {
    const bool up = (lane & o) != 0;
#pragma unroll
    for (int i = 0; i < o / 2; ++i) {
      const float send = up ? v[i] : v[i + o / 2];
      const float keep = up ? v[i + o / 2] : v[i];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, o);
    }
  }
  return v[0] + __shfl_xor_sync(0xffffffffu, v[0], 1);
}

// sums of 8 per-lane values over the warp in 9 shuffles; afterwards lane l holds the total of value (l >> 2) & 7
__device__ __forceinline__ float warp_sum8(float (&v)[8]) {
  const int lane = threadIdx.x & 31;
#pragma unroll
  for (int o = 16; o >= 4; o >>= 1) {
    const bool up = (lane & o) != 0;
#pragma unroll
    for (int i = 0; i < o / 4; ++i) {
      const float send = up ? v[i] : v[i + o / 4];
      const float keep = up ? v[i + o / 4] : v[i];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, o);
    }
  }
  float t = v[0] + __shfl_xor_sync(0xffffffffu, v[0], 2);
  return t + __shfl_xor_sync(0xffffffffu, t, 1);
}

__device__ __forceinline__ void sts128(uint32_t a, uint32_t x, uint32_t y, uint32_t z, uint32_t w) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};\n" ::"r"(a), "r"(x), "r"(y), "r"(z), "r"(w) : "memory");
}
__device__ __forceinline__ uint4 lds128(uint32_t a) {
  uint4 v;
  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];\n" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a) : "memory");
  return v;
}
__device__ __forceinline__ uint2 lds64(uint32_t a) {
  uint2 v;
  asm volatile("ld.shared.v2.b32 {%0, %1}, [%2];\n" : "=r"(v.x), "=r"(v.y) : "r"(a) : "memory");
  return v;
}
__device__ __forceinline__ void sts16(uint32_t a, uint32_t v) {
  asm volatile("st.shared.b16 [%0], %1;\n" ::"r"(a), "h"((unsigned short)v) : "memory");
}
// hi/lo bf16 split of 8 values, 16-byte stores at two shared-memory addresses (32-bit shared window addresses)
__device__ __forceinline__ void store_split8_sa(uint32_t a_hi, uint32_t a_lo, const float (&v)[8]) {
  uint32_t h[4], l[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) split_pack2(v[2 * i], v[2 * i + 1], h[i], l[i]);
  sts128(a_hi, h[0], h[1], h[2], h[3]);
  sts128(a_lo, l[0], l[1], l[2], l[3]);
}
__device__ __forceinline__ float bf_lo(uint32_t u) { return __uint_as_float(u << 16); }
__device__ __forceinline__ float bf_hi(uint32_t u) { return __uint_as_float(u & 0xffff0000u); }


// registers -> tensor memory, 16 consecutive 32-bit columns of the caller's lane
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};\n" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]),
      "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}

}  // namespace erv
