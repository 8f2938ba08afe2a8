// One-shot sum all-reduce of the flat gradient over peer memory (NVLink 5 / NVSwitch), for the data-parallel step of
// erv_b200.train (one all-reduce of 28-57 k floats per step).
//
// At this size (~230 KB) an NCCL all-reduce is pure latency (ring / tree steps, several kernel phases): it cost 0.15 ms of
// a 1.75 ms step on 8 GPUs in round 1.  Here every rank's gradient buffer lives in symmetric memory (one allocation per
// rank, every peer mapped into every process by torch.distributed._symmetric_memory); ONE kernel of 8 CTAs per rank, CTA c owning slice c of the
// buffer and running the protocol with CTA c of every peer:
//   1. publishes "my gradient is complete" into every peer's flag word and waits for all peers,
//   2. reads all `world` buffers through NVLink and adds them IN RANK ORDER (every rank computes bit-identical sums, so the
//      replicas stay identical), writing the result to a local scratch buffer,
//   3. publishes "I have finished reading" and waits for all peers, then copies the sums over its own gradient slice, so that
//      the gradient buffer holds the reduced gradient exactly as after an NCCL all-reduce.
// Flags carry a monotonically increasing epoch (no resets); waits are bounded and trap instead of hanging the GPU.
// Every rank must launch this kernel the same number of times.
#include "erv_common.cuh"

namespace erv {

constexpr int kArMaxWorld = 8;
constexpr int kArCtas = 8;  // CTA c of every rank reduces slice c; the flag protocol runs per slice

struct ArArgs {
  float* peer[kArMaxWorld];      // every rank's symmetric buffer as mapped in this process (peer[rank] is the local one)
  uint32_t* flags[kArMaxWorld];  // every rank's flag block: [2 phases][kArCtas][kArMaxWorld] words
  float* out;                    // local scratch, n floats
  const uint32_t* epoch;         // this rank's launch counter (device memory), bumped by a second kernel
  int n4, rank, world;
};

__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ float4 ld_peer4(const float* p) {  // bypasses the (non-coherent) L1 for peer data
  float4 v;
  asm volatile("ld.volatile.global.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
  return v;
}

// slice-level barrier over all ranks: publish `ep` into slot (phase, cta, my rank) of every peer, wait for every peer's slot
__device__ __forceinline__ void ar_barrier(const ArArgs& a, int phase, uint32_t ep) {
  const int t = threadIdx.x;
  __syncthreads();
  if (t < a.world) {
    __threadfence_system();
    const int slot = (phase * kArCtas + blockIdx.x) * kArMaxWorld;
    st_release_sys(a.flags[t] + slot + a.rank, ep);
    const uint32_t* mine = a.flags[a.rank] + slot + t;
    const long long t0 = clock64();
    while ((int32_t)(ld_acquire_sys(mine) - ep) < 0) {
      if (clock64() - t0 > 4000000000ll) __trap();  // ~2 s: a peer never arrived
    }
  }
  __syncthreads();
}

__global__ void __launch_bounds__(512) allreduce_oneshot_kernel(const ArArgs a) {
  const uint32_t ep = *a.epoch + 1;
  const int per = (a.n4 + kArCtas - 1) / kArCtas, beg = blockIdx.x * per, end = min(a.n4, beg + per);
  ar_barrier(a, 0, ep);  // every rank's gradient is complete
  // all peer loads of two slice elements are issued before the first add: an NVLink round trip is ~2 us, and a load that is
  // consumed right behind its issue would serialise `world` of them per element
  for (int i = beg + threadIdx.x; i < end; i += 2 * blockDim.x) {
    const int i2 = i + blockDim.x;
    const bool two = i2 < end;
    float4 v[2][kArMaxWorld];
#pragma unroll
    for (int r = 0; r < kArMaxWorld; ++r)
      if (r < a.world) {
        v[0][r] = ld_peer4(a.peer[r] + 4 * (size_t)i);
        if (two) v[1][r] = ld_peer4(a.peer[r] + 4 * (size_t)i2);
      }
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      if (e == 1 && !two) break;
      float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int r = 0; r < kArMaxWorld; ++r)
        if (r < a.world) {  // rank order: every rank computes bit-identical sums
          s.x += v[e][r].x; s.y += v[e][r].y; s.z += v[e][r].z; s.w += v[e][r].w;
        }
      reinterpret_cast<float4*>(a.out)[e ? i2 : i] = s;
    }
  }
  ar_barrier(a, 1, ep);  // every rank has read this slice of every buffer: the local one may now take the sums
  float* mine = a.peer[a.rank];
  for (int i = beg + threadIdx.x; i < end; i += blockDim.x)
    reinterpret_cast<float4*>(mine)[i] = reinterpret_cast<const float4*>(a.out)[i];
}

__global__ void bump_epoch_kernel(uint32_t* epoch) { *epoch += 1; }

}  // namespace erv

using namespace erv;

extern "C" int erv_allreduce_flag_floats(void) { return 2 * kArCtas * kArMaxWorld; }

extern "C" int erv_allreduce_oneshot(const void* const* peer_bufs, size_t n, size_t flag_offset_floats, float* scratch, int rank,
                                     int world, uint32_t* epoch_dev, void* stream) {
  ERV_CHECK_ARG(peer_bufs && scratch && epoch_dev, "erv_allreduce_oneshot: null pointer");
  ERV_CHECK_ARG(world >= 1 && world <= kArMaxWorld && rank >= 0 && rank < world, "erv_allreduce_oneshot: rank %d / world %d",
                rank, world);
  ERV_CHECK_ARG(n % 4 == 0 && flag_offset_floats >= n && flag_offset_floats % 4 == 0,
                "erv_allreduce_oneshot: n and the flag offset must be multiples of 4 floats, flags behind the data");
  ArArgs a{};
  for (int r = 0; r < world; ++r) {
    ERV_CHECK_ARG(peer_bufs[r] != nullptr, "erv_allreduce_oneshot: peer %d not mapped", r);
    a.peer[r] = static_cast<float*>(const_cast<void*>(peer_bufs[r]));
    a.flags[r] = reinterpret_cast<uint32_t*>(a.peer[r] + flag_offset_floats);
  }
  a.out = scratch; a.epoch = epoch_dev; a.n4 = (int)(n / 4); a.rank = rank; a.world = world;
  allreduce_oneshot_kernel<<<kArCtas, 512, 0, (cudaStream_t)stream>>>(a);
  ERV_LAUNCH_CHECK();
  bump_epoch_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(epoch_dev);
  ERV_LAUNCH_CHECK();
  return ERV_OK;
}
