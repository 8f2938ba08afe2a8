// Diagnostic: D[128 x N] = A[128 x K] * B[N x K]^T on tcgen05 with the no-swizzle canonical layouts of erv_umma.cuh,
// in every K-major / MN-major combination, TF32 or BF16 operands.  tests/test_umma_gpu.py checks it against torch;
// it pins the descriptor encodings the fused kernels rely on.
#include <cuda_bf16.h>

#include "erv_common.cuh"
#include "erv_umma.cuh"

namespace erv {

template <bool BF16>
__global__ void __launch_bounds__(128) umma_probe_kernel(const float* __restrict__ A, const float* __restrict__ B,
                                                         float* __restrict__ D, int N, int K, int a_mn, int b_mn) {
  using namespace umma;
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  constexpr int ES = BF16 ? 2 : 4, T = 16 / ES, KI = BF16 ? 16 : 8;
  const int tid = threadIdx.x, warp = tid >> 5;
  uint8_t* a_s = smem;
  uint8_t* b_s = smem + (size_t)128 * K * ES;
  // K-major: chunks along k adjacent (LBO = 128), row groups SBO apart.  MN-major: chunks along rows adjacent
  // (SBO = 128), groups of 8 k's LBO apart.
  const uint32_t a_lbo = a_mn ? (128 / T) * 128 : 128, a_sbo = a_mn ? 128 : (K / T) * 128;
  const uint32_t b_lbo = b_mn ? (N / T) * 128 : 128, b_sbo = b_mn ? 128 : (K / T) * 128;

  auto put = [&](uint8_t* base, uint32_t off, float v) {
    if (BF16) *reinterpret_cast<__nv_bfloat16*>(base + off) = __float2bfloat16(v);
    else *reinterpret_cast<float*>(base + off) = to_tf32(v);
  };
  for (int k = 0; k < K; ++k) {
    float v = A[(size_t)tid * K + k];
    put(a_s, a_mn ? off_mnmajor(tid, k, T, ES, a_lbo, a_sbo) : off_kmajor(tid, k, T, ES, a_lbo, a_sbo), v);
  }
  for (int r = tid; r < N; r += 128)
    for (int k = 0; k < K; ++k) {
      float v = B[(size_t)r * K + k];
      put(b_s, b_mn ? off_mnmajor(r, k, T, ES, b_lbo, b_sbo) : off_kmajor(r, k, T, ES, b_lbo, b_sbo), v);
    }
  uint32_t cols = 32;
  while (cols < (uint32_t)N) cols <<= 1;
  if (warp == 0) tmem_alloc(&tmem_base_s, cols);
  if (tid == 0) {
    mbar_init(&bar, 1);
    mbar_init_fence();
  }
  fence_smem_to_async();
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem_base = tmem_base_s;
  if (tid == 0) {
    const uint32_t idesc = make_idesc(BF16 ? FMT_BF16 : FMT_TF32, 128, N, a_mn != 0, b_mn != 0);
    for (int s = 0; s < K / KI; ++s) {
      // K-major: one MMA consumes KI/T = 2 chunks along k; MN-major: KI/8 groups of 8 k's
      const uint32_t a_off = a_mn ? s * (KI / 8) * a_lbo : s * 2 * a_lbo;
      const uint32_t b_off = b_mn ? s * (KI / 8) * b_lbo : s * 2 * b_lbo;
      const uint64_t ad = make_desc(smem_u32(a_s) + a_off, a_lbo, a_sbo);
      const uint64_t bd = make_desc(smem_u32(b_s) + b_off, b_lbo, b_sbo);
      if (BF16) mma_f16(tmem_base, ad, bd, idesc, s > 0);
      else mma_tf32(tmem_base, ad, bd, idesc, s > 0);
    }
    commit(&bar);
  }
  mbar_wait(&bar, 0);
  fence_after_sync();
  for (int c0 = 0; c0 < N; c0 += 16) {
    float v[16];
    tmem_ld16(tmem_base + ((uint32_t)(warp * 32) << 16) + c0, v);
#pragma unroll
    for (int i = 0; i < 16; ++i) D[(size_t)tid * N + c0 + i] = v[i];
  }
  fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, cols);
}

// Timing probe: one thread issues `iters` back-to-back MMAs (M=128) of the given N / kind and reports the elapsed
// SM clocks from first issue to completion.  Operands are whatever is in shared memory (zero-filled).
__global__ void __launch_bounds__(128) umma_timing_kernel(int N, int bf16, int a_mn, int b_mn, int iters,
                                                          long long* __restrict__ cycles) {
  using namespace umma;
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < 32 * 1024 / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (warp == 0) tmem_alloc(&tmem_base_s, 512);
  if (tid == 0) {
    mbar_init(&bar, 1);
    mbar_init_fence();
  }
  fence_smem_to_async();
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tm = tmem_base_s;
  if (tid == 0) {
    const uint32_t idesc = make_idesc(bf16 ? FMT_BF16 : FMT_TF32, 128, N, a_mn != 0, b_mn != 0);
    const uint64_t ad = make_desc(smem_u32(smem), 128, 256);         // 128 rows: 16 groups x 256 B
    const uint64_t bd = make_desc(smem_u32(smem) + 8192, 128, 256);  // up to 256 rows: 32 groups x 256 B
    const long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
      if (bf16) mma_f16(tm, ad, bd, idesc, i > 0);
      else mma_tf32(tm, ad, bd, idesc, i > 0);
    }
    commit(&bar);
    mbar_wait(&bar, 0);
    cycles[0] = clock64() - t0;
  }
  fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tm, 512);
}


// Same product with the A operand in tensor memory (written by the threads with tcgen05.st, lane = row): pins the TMEM A layout
// the fused kernels rely on (bf16: column c holds k = 2c in the low and k = 2c+1 in the high half; tf32: one k per column).
template <bool BF16>
__global__ void __launch_bounds__(128) umma_probe_ts_kernel(const float* __restrict__ A, const float* __restrict__ B,
                                                            float* __restrict__ D, int N, int K, int b_mn) {
  using namespace umma;
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  constexpr int ES = BF16 ? 2 : 4, T = 16 / ES, KI = BF16 ? 16 : 8;
  const int tid = threadIdx.x, warp = tid >> 5;
  uint8_t* b_s = smem;
  const uint32_t b_lbo = b_mn ? (N / T) * 128 : 128, b_sbo = b_mn ? 128 : (K / T) * 128;
  for (int r = tid; r < N; r += 128)
    for (int k = 0; k < K; ++k) {
      const float v = B[(size_t)r * K + k];
      const uint32_t off = b_mn ? off_mnmajor(r, k, T, ES, b_lbo, b_sbo) : off_kmajor(r, k, T, ES, b_lbo, b_sbo);
      if (BF16) *reinterpret_cast<__nv_bfloat16*>(b_s + off) = __float2bfloat16(v);
      else *reinterpret_cast<float*>(b_s + off) = to_tf32(v);
    }
  if (warp == 0) tmem_alloc(&tmem_base_s, 512);
  if (tid == 0) {
    mbar_init(&bar, 1);
    mbar_init_fence();
  }
  fence_smem_to_async();
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tm = tmem_base_s;
  const uint32_t lane_off = (uint32_t)(warp * 32) << 16;
  const uint32_t COL_A = 256;  // D: columns [0, N), A: columns [256, 256 + K * ES / 4)
  for (int c0 = 0; c0 < K * ES / 4; c0 += 8) {
    uint32_t r[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (BF16) {
        const __nv_bfloat16 lo = __float2bfloat16(A[(size_t)tid * K + 2 * (c0 + i)]);
        const __nv_bfloat16 hi = __float2bfloat16(A[(size_t)tid * K + 2 * (c0 + i) + 1]);
        r[i] = (uint32_t)__bfloat16_as_ushort(lo) | ((uint32_t)__bfloat16_as_ushort(hi) << 16);
      } else {
        r[i] = __float_as_uint(to_tf32(A[(size_t)tid * K + c0 + i]));
      }
    }
    tmem_st8(tm + lane_off + COL_A + c0, r);
  }
  tmem_wait_st();
  fence_before_sync();
  __syncthreads();
  if (tid == 0) {
    fence_after_sync();
    const uint32_t idesc = make_idesc(BF16 ? FMT_BF16 : FMT_TF32, 128, N, false, b_mn != 0);
    for (int s = 0; s < K / KI; ++s) {
      const uint32_t b_off = b_mn ? s * (KI / 8) * b_lbo : s * 2 * b_lbo;
      const uint64_t bd = make_desc(smem_u32(b_s) + b_off, b_lbo, b_sbo);
      if (BF16) mma_f16_ts(tm, tm + COL_A + s * 8, bd, idesc, s > 0);
      else mma_tf32_ts(tm, tm + COL_A + s * 8, bd, idesc, s > 0);
    }
    commit(&bar);
  }
  mbar_wait(&bar, 0);
  fence_after_sync();
  for (int c0 = 0; c0 < N; c0 += 16) {
    float v[16];
    tmem_ld16(tm + lane_off + c0, v);
#pragma unroll
    for (int i = 0; i < 16; ++i) D[(size_t)tid * N + c0 + i] = v[i];
  }
  fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tm, 512);
}

// Timing probe 2: `iters` MMAs rotating over `nacc` independent accumulators (column blocks of N), M = 64 or 128, the A operand
// from shared memory or tensor memory, optionally `nchain` k-steps per accumulator visit.  cycles[0] = first issue -> completion,
// cycles[1] = issue loop only (the issuing thread's own cost).
__global__ void __launch_bounds__(128) umma_timing2_kernel(int N, int M, int bf16, int a_tmem, int b_mn, int nacc, int iters,
                                                           int elected, long long* __restrict__ cycles) {
  using namespace umma;
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < 32 * 1024 / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (warp == 0) tmem_alloc(&tmem_base_s, 512);
  if (tid == 0) {
    mbar_init(&bar, 1);
    mbar_init_fence();
  }
  fence_smem_to_async();
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tm = tmem_base_s;
  {  // defined A operand in tensor memory
    uint32_t z[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    tmem_st8(tm + ((uint32_t)(warp * 32) << 16) + 480, z);
    tmem_wait_st();
  }
  fence_before_sync();
  __syncthreads();
  const uint32_t idesc = make_idesc(bf16 ? FMT_BF16 : FMT_TF32, M, N, false, b_mn != 0);
  const uint64_t ad = make_desc(smem_u32(smem), 128, 256);
  const uint64_t bd = b_mn ? make_desc(smem_u32(smem) + 8192, (uint32_t)(N / 8) * 128, 128)
                           : make_desc(smem_u32(smem) + 8192, 128, 256);
  auto issue_all = [&]() {
    int a = 0;
    for (int i = 0; i < iters; ++i) {
      const uint32_t d = tm + (uint32_t)a * N;
      if (a_tmem) {
        if (bf16) mma_f16_ts(d, tm + 480, bd, idesc, i >= nacc);
        else mma_tf32_ts(d, tm + 480, bd, idesc, i >= nacc);
      } else {
        if (bf16) mma_f16(d, ad, bd, idesc, i >= nacc);
        else mma_tf32(d, ad, bd, idesc, i >= nacc);
      }
      if (++a == nacc) a = 0;
    }
  };
  if (elected) {  // issue from a converged warp under elect.sync (the CUTLASS pattern)
    if (warp == 0) {
      fence_after_sync();
      const long long t0 = clock64();
      if (elect_one()) {
        issue_all();
        commit(&bar);
      }
      __syncwarp();
      const long long t1 = clock64();
      mbar_wait(&bar, 0);
      if (tid == 0) {
        cycles[0] = clock64() - t0;
        cycles[1] = t1 - t0;
      }
    }
  } else if (tid == 0) {  // divergent single thread (the round-1 kernels)
    fence_after_sync();
    const long long t0 = clock64();
    issue_all();
    const long long t1 = clock64();
    commit(&bar);
    mbar_wait(&bar, 0);
    cycles[0] = clock64() - t0;
    cycles[1] = t1 - t0;
  }
  fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tm, 512);
}

// TMEM <-> register bandwidth and issue-slot probes on one SM with `threads` threads.
//   mode 0: tcgen05.ld 32x32b.x32 (128 B per thread per instruction), mode 1: tcgen05.st 32x32b.x8 (32 B per thread)
//   mode 2: cvt.rn.bf16x2.f32, mode 3: ex2.approx, mode 4: both interleaved 1:1, mode 5: FFMA chain (reference)
__global__ void __launch_bounds__(512) unit_probe_kernel(int mode, int iters, long long* __restrict__ cycles, float* sink) {
  using namespace umma;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5;
  if (warp == 0) tmem_alloc(&tmem_base_s, 512);
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tm = tmem_base_s + ((uint32_t)((warp & 3) * 32) << 16);
  float acc = (float)tid * 1e-3f, acc2 = 1.0f + (float)tid * 1e-4f;
  uint32_t pk = 0;
  __syncthreads();
  const long long t0 = clock64();
  if (mode == 0) {
    uint32_t r[32];
    for (int i = 0; i < iters; ++i) {
      tmem_ld32_nowait(tm + (uint32_t)((i & 7) * 32), r);
      if ((i & 3) == 3) tmem_wait_ld();
    }
    tmem_wait_ld();
    pk = r[0] ^ r[31];
  } else if (mode == 1) {
    uint32_t r[8] = {1, 2, 3, 4, 5, 6, 7, 8};
    for (int i = 0; i < iters; ++i) {
      tmem_st8(tm + (uint32_t)((i & 31) * 8), r);
      if ((i & 7) == 7) tmem_wait_st();
    }
    tmem_wait_st();
  } else if (mode == 2) {  // 4 F2FP per iteration
    float x0 = acc, x1 = acc2, x2 = acc + 1.f, x3 = acc2 + 1.f;
#pragma unroll 4
    for (int i = 0; i < iters; ++i) {
      uint32_t o0, o1, o2, o3;
      asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(o0) : "f"(x0), "f"(x1));
      asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(o1) : "f"(x1), "f"(x2));
      asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(o2) : "f"(x2), "f"(x3));
      asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(o3) : "f"(x3), "f"(x0));
      x0 = __uint_as_float(o0 & 0x3fffffffu); x1 = __uint_as_float(o1 & 0x3fffffffu);
      x2 = __uint_as_float(o2 & 0x3fffffffu); x3 = __uint_as_float(o3 & 0x3fffffffu);
    }
    acc = x0 + x1 + x2 + x3;
  } else if (mode == 3) {  // 4 dependent-chain MUFU.EX2 (+ 4 FADD) per iteration
    float x0 = acc, x1 = acc2, x2 = acc + 1.f, x3 = acc2 + 1.f;
#pragma unroll 4
    for (int i = 0; i < iters; ++i) {
      asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(x0));
      asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(x1));
      asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(x2));
      asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(x3));
      x0 -= 1.f; x1 -= 1.f; x2 -= 1.f; x3 -= 1.f;
    }
    acc = x0 + x1 + x2 + x3;
  } else if (mode == 4) {  // 4 MUFU.EX2 + 4 FADD + 2 F2FP per iteration (the feature phase's mix)
    float x0 = acc, x1 = acc2, x2 = acc + 1.f, x3 = acc2 + 1.f;
#pragma unroll 4
    for (int i = 0; i < iters; ++i) {
      asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(x0));
      asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(x1));
      asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(x2));
      asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(x3));
      uint32_t o0, o1;
      asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(o0) : "f"(x0), "f"(x1));
      asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(o1) : "f"(x2), "f"(x3));
      pk ^= o0 ^ o1;
      x0 -= 1.f; x1 -= 1.f; x2 -= 1.f; x3 -= 1.f;
    }
    acc = x0 + x1 + x2 + x3;
  } else {  // 4 independent FFMA chains per iteration
    float x0 = acc, x1 = acc2, x2 = acc + 1.f, x3 = acc2 + 1.f;
#pragma unroll 4
    for (int i = 0; i < iters; ++i) {
      x0 = fmaf(x0, 1.0001f, acc2); x1 = fmaf(x1, 1.0001f, acc2); x2 = fmaf(x2, 1.0001f, acc2); x3 = fmaf(x3, 1.0001f, acc2);
    }
    acc = x0 + x1 + x2 + x3;
  }
  __syncthreads();
  const long long t1 = clock64();
  if (tid == 0) cycles[0] = t1 - t0;
  if (sink != nullptr) sink[tid] = acc + acc2 + __uint_as_float(pk);
  fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base_s, 512);
}

}  // namespace erv

using namespace erv;

namespace erv { void set_tc_trace(long long* p); }
extern "C" void erv_debug_set_trace(long long* device_buffer) { erv::set_tc_trace(device_buffer); }

extern "C" int erv_debug_umma_timing(int N, int bf16, int a_mn_major, int b_mn_major, int iters, long long* cycles,
                                     void* stream) {
  ERV_CHECK_ARG(cycles && N >= 16 && N <= 256 && N % 16 == 0 && iters > 0, "erv_debug_umma_timing: bad arguments");
  umma_timing_kernel<<<1, 128, 32 * 1024, (cudaStream_t)stream>>>(N, bf16, a_mn_major, b_mn_major, iters, cycles);
  ERV_LAUNCH_CHECK();
  return ERV_OK;
}

extern "C" int erv_debug_umma_gemm(const float* A, const float* B, float* D, int N, int K, int a_mn_major,
                                   int b_mn_major, int bf16, void* stream) {
  ERV_CHECK_ARG(A && B && D, "erv_debug_umma_gemm: null pointer");
  ERV_CHECK_ARG(N >= 16 && N <= 256 && N % 16 == 0, "erv_debug_umma_gemm: N %d must be a multiple of 16 in [16,256]", N);
  ERV_CHECK_ARG(K >= 16 && K % 16 == 0 && K <= 128, "erv_debug_umma_gemm: K %d must be a multiple of 16 in [16,128]", K);
  const size_t es = bf16 ? 2 : 4;
  const size_t smem = (size_t)(128 + N) * K * es;
  cudaStream_t st = (cudaStream_t)stream;
  if (bf16) {
    ERV_CUDA(allow_smem(umma_probe_kernel<true>, smem));
    umma_probe_kernel<true><<<1, 128, smem, st>>>(A, B, D, N, K, a_mn_major, b_mn_major);
  } else {
    ERV_CUDA(allow_smem(umma_probe_kernel<false>, smem));
    umma_probe_kernel<false><<<1, 128, smem, st>>>(A, B, D, N, K, a_mn_major, b_mn_major);
  }
  ERV_LAUNCH_CHECK();
  return ERV_OK;
}

extern "C" int erv_debug_umma_gemm_ts(const float* A, const float* B, float* D, int N, int K, int b_mn_major, int bf16,
                                      void* stream) {
  ERV_CHECK_ARG(A && B && D, "erv_debug_umma_gemm_ts: null pointer");
  ERV_CHECK_ARG(N >= 16 && N <= 256 && N % 16 == 0, "erv_debug_umma_gemm_ts: N %d must be a multiple of 16 in [16,256]", N);
  ERV_CHECK_ARG(K >= 16 && K % 16 == 0 && K <= 256, "erv_debug_umma_gemm_ts: K %d must be a multiple of 16 in [16,256]", K);
  ERV_CHECK_ARG(bf16 || K <= 128, "erv_debug_umma_gemm_ts: tf32 K %d > 128", K);
  const size_t smem = (size_t)N * K * (bf16 ? 2 : 4);
  cudaStream_t st = (cudaStream_t)stream;
  if (bf16) {
    ERV_CUDA(allow_smem(umma_probe_ts_kernel<true>, smem));
    umma_probe_ts_kernel<true><<<1, 128, smem, st>>>(A, B, D, N, K, b_mn_major);
  } else {
    ERV_CUDA(allow_smem(umma_probe_ts_kernel<false>, smem));
    umma_probe_ts_kernel<false><<<1, 128, smem, st>>>(A, B, D, N, K, b_mn_major);
  }
  ERV_LAUNCH_CHECK();
  return ERV_OK;
}

extern "C" int erv_debug_umma_timing2(int N, int M, int bf16, int a_tmem, int b_mn_major, int nacc, int iters,
                                      int elected, long long* cycles, void* stream) {
  ERV_CHECK_ARG(cycles && N >= 16 && N <= 256 && N % 16 == 0 && iters > 0 && (M == 64 || M == 128) && nacc >= 1 &&
                    nacc * N <= 448,
                "erv_debug_umma_timing2: bad arguments");
  umma_timing2_kernel<<<1, 128, 32 * 1024, (cudaStream_t)stream>>>(N, M, bf16, a_tmem, b_mn_major, nacc, iters, elected,
                                                                         cycles);
  ERV_LAUNCH_CHECK();
  return ERV_OK;
}

extern "C" int erv_debug_unit_probe(int mode, int threads, int iters, long long* cycles, float* sink, void* stream) {
  ERV_CHECK_ARG(cycles && mode >= 0 && mode <= 5 && threads >= 32 && threads <= 512 && threads % 32 == 0 && iters > 0,
                "erv_debug_unit_probe: bad arguments");
  unit_probe_kernel<<<1, threads, 0, (cudaStream_t)stream>>>(mode, iters, cycles, sink);
  ERV_LAUNCH_CHECK();
  return ERV_OK;
}
