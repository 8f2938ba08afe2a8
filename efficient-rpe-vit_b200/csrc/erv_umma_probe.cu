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
