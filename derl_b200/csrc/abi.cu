// Library plumbing for the derl_b200 C ABI: error strings, device check, launch counter.
#include <atomic>
#include <stdarg.h>
#include <string.h>

#include "common.cuh"

namespace derl {

static thread_local char g_error[512] = "";
static std::atomic<uint64_t> g_launches{0};
static std::atomic<int> g_device_state{0};  // 0 unknown, 1 ok
static std::atomic<int> g_sm_count{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_error, sizeof(g_error), fmt, ap);
  va_end(ap);
}

int cuda_fail(cudaError_t err, const char* what) {
  set_error("CUDA error %d (%s) in %s", (int)err, cudaGetErrorString(err), what);
  return DERL_E_CUDA;
}

void count_launch(unsigned n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

int require_device() {
  if (g_device_state.load(std::memory_order_acquire) == 1) return DERL_OK;
  int ndev = 0;
  cudaError_t err = cudaGetDeviceCount(&ndev);
  if (err != cudaSuccess || ndev == 0) {
    cudaGetLastError();
    set_error("derl_b200 needs a CUDA device (cudaGetDeviceCount: %s); there is no CPU fallback",
              err == cudaSuccess ? "0 devices" : cudaGetErrorString(err));
    return DERL_E_NO_DEVICE;
  }
  int dev = 0;
  cudaDeviceProp prop;
  if ((err = cudaGetDevice(&dev)) != cudaSuccess ||
      (err = cudaGetDeviceProperties(&prop, dev)) != cudaSuccess) {
    cudaGetLastError();
    set_error("cannot query CUDA device: %s", cudaGetErrorString(err));
    return DERL_E_NO_DEVICE;
  }
  if (prop.major != 10) {
    set_error("derl_b200 is built for sm_100a only; device %d is %s (sm_%d%d)", dev, prop.name,
              prop.major, prop.minor);
    return DERL_E_NO_DEVICE;
  }
  g_sm_count.store(prop.multiProcessorCount, std::memory_order_release);
  g_device_state.store(1, std::memory_order_release);
  return DERL_OK;
}

int sm_count() {
  int n = g_sm_count.load(std::memory_order_acquire);
  return n > 0 ? n : 148;
}

}  // namespace derl

extern "C" {

int derl_b200_abi_version(void) { return DERL_B200_ABI_VERSION; }

const char* derl_b200_last_error(void) { return derl::g_error; }

int derl_b200_device_ok(void) { return derl::require_device(); }

uint64_t derl_b200_launch_count(void) {
  return derl::g_launches.load(std::memory_order_relaxed);
}

}  // extern "C"
