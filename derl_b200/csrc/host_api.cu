// Host-buffer convenience entry points of the C ABI (NumPy-style callers).
#include <mutex>

#include "common.cuh"

namespace derl {
namespace {

// Grow-only device scratch shared by the *_host entry points (one caller at a time).
std::mutex g_scratch_mutex;
void* g_scratch = nullptr;
size_t g_scratch_bytes = 0;

int scratch(size_t bytes, void** out) {
  if (bytes > g_scratch_bytes) {
    if (g_scratch != nullptr) DERL_CUDA(cudaFree(g_scratch));
    g_scratch = nullptr;
    g_scratch_bytes = 0;
    DERL_CUDA(cudaMalloc(&g_scratch, bytes));
    g_scratch_bytes = bytes;
  }
  *out = g_scratch;
  return DERL_OK;
}

inline size_t up256(size_t b) { return (b + 255) & ~(size_t)255; }

}  // namespace
}  // namespace derl

using namespace derl;

extern "C" int derl_b200_gae_host(const void* rewards, int rewards_f64, const float* values,
                                  const uint8_t* resets, const float* last_value, int64_t T,
                                  int64_t N, double gamma, double lambda, int normalize,
                                  double epsilon, float* adv, float* vt, void* stream) {
  DERL_REQUIRE(T >= 1 && N >= 1, "gae_host: need T >= 1 and N >= 1");
  DERL_REQUIRE(rewards && values && resets && last_value && adv && vt, "gae_host: null pointer");
  int rc = require_device();
  if (rc != DERL_OK) return rc;
  std::lock_guard<std::mutex> lock(g_scratch_mutex);
  const size_t n = (size_t)T * (size_t)N;
  const size_t rb = n * (rewards_f64 ? 8 : 4);
  const size_t ws = derl_b200_gae_workspace_bytes(T, N);
  const size_t total = up256(rb) + 3 * up256(n * 4) + up256(n) + up256((size_t)N * 4) +
                       up256(ws) + 256;
  void* base = nullptr;
  if ((rc = scratch(total, &base)) != DERL_OK) return rc;
  uint8_t* p = reinterpret_cast<uint8_t*>(base);
  void* d_r = p;            p += up256(rb);
  float* d_v = (float*)p;   p += up256(n * 4);
  float* d_a = (float*)p;   p += up256(n * 4);
  float* d_vt = (float*)p;  p += up256(n * 4);
  uint8_t* d_z = p;         p += up256(n);
  float* d_lv = (float*)p;  p += up256((size_t)N * 4);
  void* d_ws = p;           p += up256(ws);
  double* d_stats = (double*)p;
  cudaStream_t st = as_stream(stream);
  DERL_CUDA(cudaMemcpyAsync(d_r, rewards, rb, cudaMemcpyHostToDevice, st));
  DERL_CUDA(cudaMemcpyAsync(d_v, values, n * 4, cudaMemcpyHostToDevice, st));
  DERL_CUDA(cudaMemcpyAsync(d_z, resets, n, cudaMemcpyHostToDevice, st));
  DERL_CUDA(cudaMemcpyAsync(d_lv, last_value, (size_t)N * 4, cudaMemcpyHostToDevice, st));
  rc = derl_b200_gae(d_r, rewards_f64, d_v, d_z, d_lv, T, N, gamma, lambda, d_a, d_vt,
                     normalize ? d_stats : nullptr, normalize ? d_ws : nullptr, ws,
                     DERL_GAE_AUTO, stream);
  if (rc != DERL_OK) return rc;
  if (normalize) {
    rc = derl_b200_normalize(d_a, d_a, (int64_t)n, d_stats, epsilon, stream);
    if (rc != DERL_OK) return rc;
  }
  DERL_CUDA(cudaMemcpyAsync(adv, d_a, n * 4, cudaMemcpyDeviceToHost, st));
  DERL_CUDA(cudaMemcpyAsync(vt, d_vt, n * 4, cudaMemcpyDeviceToHost, st));
  DERL_CUDA(cudaStreamSynchronize(st));
  return DERL_OK;
}
