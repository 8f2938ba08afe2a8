// Library-wide state (thread-local error text, launch counter) and the flat fused Adam step.
#include <stdarg.h>

#include <map>
#include <mutex>
#include <utility>

#include "erv_common.cuh"

namespace erv {

static thread_local char g_err[512] = "";
std::atomic<uint64_t> g_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

cudaError_t ensure_smem(const void* fn, size_t bytes) {
  static std::mutex mu;
  static std::map<std::pair<const void*, int>, size_t> seen;
  if (bytes <= 48 * 1024) return cudaSuccess;
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  std::lock_guard<std::mutex> lock(mu);
  auto key = std::make_pair(fn, dev);
  auto it = seen.find(key);
  if (it != seen.end() && it->second >= bytes) return cudaSuccess;
  e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
  if (e == cudaSuccess) seen[key] = bytes;
  return e;
}

// torch.optim.Adam / AdamW semantics (no amsgrad, no maximize):
//   g = grad * grad_scale (+ wd * p when not decoupled);  p *= 1 - lr*wd when decoupled
//   m = b1 m + (1-b1) g ; v = b2 v + (1-b2) g^2
//   p -= lr / (1 - b1^t) * m / (sqrt(v) / sqrt(1 - b2^t) + eps)
// hyper (optional, device): [lr, beta1, beta2, eps, weight_decay] read at run time, so a captured CUDA graph follows a
// learning-rate schedule (the by-value arguments are baked into the graph at capture).
__global__ void adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                            float* __restrict__ v, size_t n, float lr, float b1, float b2, float eps, float wd,
                            int decoupled, float gscale, int64_t step, const int64_t* __restrict__ step_dev,
                            const float* __restrict__ hyper) {
  if (hyper != nullptr) {
    lr = hyper[0]; b1 = hyper[1]; b2 = hyper[2]; eps = hyper[3]; wd = hyper[4];
  }
  const int64_t t = step_dev ? *step_dev : step;
  const float bc1 = 1.f - powf(b1, (float)t);
  const float bc2s = sqrtf(1.f - powf(b2, (float)t));
  const float step_size = lr / bc1;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    float pi = p[i];
    float gi = g[i] * gscale;
    if (wd != 0.f) {
      if (decoupled) pi *= 1.f - lr * wd;
      else gi += wd * pi;
    }
    float mi = b1 * m[i] + (1.f - b1) * gi;
    float vi = b2 * v[i] + (1.f - b2) * gi * gi;
    m[i] = mi;
    v[i] = vi;
    p[i] = pi - step_size * (mi / (sqrtf(vi) / bc2s + eps));
  }
}

}  // namespace erv

using namespace erv;

extern "C" int erv_abi_version(void) { return ERV_ABI_VERSION; }
extern "C" const char* erv_last_error(void) { return g_err; }
extern "C" uint64_t erv_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }
extern "C" void erv_reset_launch_count(void) { g_launches.store(0, std::memory_order_relaxed); }

extern "C" int erv_adam_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, size_t n, float lr,
                             float beta1, float beta2, float eps, float weight_decay, int decoupled_wd,
                             float grad_scale, int64_t step, const int64_t* step_dev, void* stream) {
  ERV_CHECK_ARG(param && grad && exp_avg && exp_avg_sq, "erv_adam_step: null pointer");
  ERV_CHECK_ARG(step_dev || step >= 1, "erv_adam_step: step must be >= 1");
  if (n == 0) return ERV_OK;
  size_t blocks = (n + 255) / 256;
  if (blocks > (size_t)kNumSMs * 8) blocks = (size_t)kNumSMs * 8;
  adam_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(param, grad, exp_avg, exp_avg_sq, n, lr, beta1,
                                                                  beta2, eps, weight_decay, decoupled_wd, grad_scale,
                                                                  step, step_dev, nullptr);
  ERV_LAUNCH_CHECK();
  return ERV_OK;
}

extern "C" int erv_adam_step_dev(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, size_t n,
                                 const float* hyper, int decoupled_wd, float grad_scale, const int64_t* step_dev,
                                 void* stream) {
  ERV_CHECK_ARG(param && grad && exp_avg && exp_avg_sq && hyper && step_dev, "erv_adam_step_dev: null pointer");
  if (n == 0) return ERV_OK;
  size_t blocks = (n + 255) / 256;
  if (blocks > (size_t)kNumSMs * 8) blocks = (size_t)kNumSMs * 8;
  adam_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(param, grad, exp_avg, exp_avg_sq, n, 0.f, 0.f, 0.f, 0.f,
                                                                  0.f, decoupled_wd, grad_scale, 0, step_dev, hyper);
  ERV_LAUNCH_CHECK();
  return ERV_OK;
}
