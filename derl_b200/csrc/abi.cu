// Library plumbing for the derl_b200 C ABI: error strings, device check, launch counter.
#include <atomic>
#include <mutex>
#include <vector>
#include <stdarg.h>
#include <string.h>

#include "common.cuh"
#include "tcgen05.cuh"

namespace derl {

static thread_local char g_error[512] = "";
static std::atomic<uint64_t> g_launches{0};
constexpr int kMaxDevices = 64;
// Everything cached about a device is keyed by its ordinal: a process may switch devices
// between calls (cudaSetDevice) and function attributes are per device / context.
struct DeviceSlot {
  std::atomic<int> state{0};      // 0 unknown, 1 usable (sm_100), -1 rejected
  std::atomic<int> sm_count{0};
};
static DeviceSlot g_devices[kMaxDevices];
static std::mutex g_attr_mutex;
struct AttrKey { int device; const void* func; int bytes; };
static std::vector<AttrKey> g_attrs;   // (device, kernel) pairs whose smem limit is raised

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

static int current_device_or(int fallback) {
  int dev = fallback;
  if (cudaGetDevice(&dev) != cudaSuccess) {
    cudaGetLastError();
    return fallback;
  }
  return dev;
}

int require_device() {
  int dev = -1;
  cudaError_t err = cudaGetDevice(&dev);
  if (err == cudaSuccess && dev >= 0 && dev < kMaxDevices &&
      g_devices[dev].state.load(std::memory_order_acquire) == 1) {
    return DERL_OK;
  }
  int ndev = 0;
  err = cudaGetDeviceCount(&ndev);
  if (err != cudaSuccess || ndev == 0) {
    cudaGetLastError();
    set_error("derl_b200 needs a CUDA device (cudaGetDeviceCount: %s); there is no CPU fallback",
              err == cudaSuccess ? "0 devices" : cudaGetErrorString(err));
    return DERL_E_NO_DEVICE;
  }
  cudaDeviceProp prop;
  if ((err = cudaGetDevice(&dev)) != cudaSuccess ||
      (err = cudaGetDeviceProperties(&prop, dev)) != cudaSuccess) {
    cudaGetLastError();
    set_error("cannot query CUDA device: %s", cudaGetErrorString(err));
    return DERL_E_NO_DEVICE;
  }
  // the cubin is sm_100a: arch-specific code does not run on sm_103 (which also has no INT8 MMA)
  if (prop.major != 10 || prop.minor != 0 || dev >= kMaxDevices) {
    set_error("derl_b200 is built for sm_100a only; device %d is %s (sm_%d%d)", dev, prop.name,
              prop.major, prop.minor);
    return DERL_E_NO_DEVICE;
  }
  g_devices[dev].sm_count.store(prop.multiProcessorCount, std::memory_order_release);
  g_devices[dev].state.store(1, std::memory_order_release);
  return DERL_OK;
}

int sm_count() {
  const int dev = current_device_or(0);
  int n = dev < kMaxDevices ? g_devices[dev].sm_count.load(std::memory_order_acquire) : 0;
  if (n <= 0 && require_device() == DERL_OK) {
    n = g_devices[current_device_or(0)].sm_count.load(std::memory_order_acquire);
  }
  return n > 0 ? n : 148;
}

int ensure_dynamic_smem(const void* func, int bytes) {
  const int dev = current_device_or(0);
  std::lock_guard<std::mutex> lock(g_attr_mutex);
  for (const AttrKey& k : g_attrs) {
    if (k.device == dev && k.func == func && k.bytes >= bytes) return DERL_OK;
  }
  // not a stream operation: done once per (device, kernel), outside any graph capture in practice
  cudaError_t err = cudaFuncSetAttribute(func, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
  if (err != cudaSuccess) return cuda_fail(err, "cudaFuncSetAttribute(MaxDynamicSharedMemorySize)");
  g_attrs.push_back(AttrKey{dev, func, bytes});
  return DERL_OK;
}

TmapEncodeFn tmap_encode_fn() {
  static std::atomic<TmapEncodeFn> cached{nullptr};
  TmapEncodeFn fn = cached.load(std::memory_order_acquire);
  if (fn == nullptr) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) ==
            cudaSuccess &&
        q == cudaDriverEntryPointSuccess) {
      fn = reinterpret_cast<TmapEncodeFn>(p);
      cached.store(fn, std::memory_order_release);
    } else {
      cudaGetLastError();
    }
  }
  return fn;
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
