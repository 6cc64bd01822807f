// Shared device/host helpers for the derl_b200 kernels (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "derl_b200.h"

namespace derl {

// ---------------------------------------------------------------- host-side error plumbing
void set_error(const char* fmt, ...);
int cuda_fail(cudaError_t err, const char* what);
int require_device();          // DERL_OK or DERL_E_NO_DEVICE for the CURRENT device (cached per device)
int sm_count();                // multiprocessor count of the current device (cached per device)
// raise a kernel's dynamic shared memory limit once per (device, kernel); DERL_OK or an error code
int ensure_dynamic_smem(const void* func, int bytes);
void count_launch(unsigned n = 1);

#define DERL_CUDA(call)                                        \
  do {                                                         \
    cudaError_t err__ = (call);                                \
    if (err__ != cudaSuccess) return derl::cuda_fail(err__, #call); \
  } while (0)

#define DERL_REQUIRE(cond, ...)          \
  do {                                   \
    if (!(cond)) {                       \
      derl::set_error(__VA_ARGS__);      \
      return DERL_E_INVALID;             \
    }                                    \
  } while (0)

#define DERL_LAUNCH_CHECK(name)                                  \
  do {                                                           \
    cudaError_t err__ = cudaGetLastError();                      \
    if (err__ != cudaSuccess) return derl::cuda_fail(err__, name); \
    derl::count_launch();                                        \
  } while (0)

static inline cudaStream_t as_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }

// Partial-sum workspace layout shared by every two-stage reduction in this library:
//   [0, 16)                 unsigned ticket counter (zeroed by the launcher, self-resetting)
//   [16, 16 + 8*K*blocks)   K float64 partials per block, block-major
constexpr size_t kTicketBytes = 16;
constexpr int kMaxReduceBlocks = 4096;

#ifdef __CUDACC__
// ---------------------------------------------------------------- warp / block reductions
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
  return v;
}

// Sum K doubles across the block; result valid in thread 0.  scratch: K * 32 doubles.
template <int K>
__device__ __forceinline__ void block_sum(double (&v)[K], double* scratch) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int nwarps = (blockDim.x + 31) >> 5;
#pragma unroll
  for (int k = 0; k < K; ++k) v[k] = warp_sum(v[k]);
  if (lane == 0) {
#pragma unroll
    for (int k = 0; k < K; ++k) scratch[k * 32 + warp] = v[k];
  }
  __syncthreads();
  if (warp == 0) {
#pragma unroll
    for (int k = 0; k < K; ++k) {
      double x = lane < nwarps ? scratch[k * 32 + lane] : 0.0;
      v[k] = warp_sum(x);
    }
  }
}

// Publish this block's K partials, take a ticket; returns true in ALL threads of the
// last block to arrive (which then reads every block's partials in block order, so the
// final sum is independent of scheduling).
template <int K>
__device__ __forceinline__ bool publish_partials(const double (&v)[K], void* workspace,
                                                 int* is_last_smem) {
  unsigned* ticket = reinterpret_cast<unsigned*>(workspace);
  double* partials = reinterpret_cast<double*>(reinterpret_cast<char*>(workspace) + kTicketBytes);
  if (threadIdx.x == 0) {
#pragma unroll
    for (int k = 0; k < K; ++k) partials[(size_t)blockIdx.x * K + k] = v[k];
    __threadfence();
    unsigned t = atomicAdd(ticket, 1u);
    int last = (t == gridDim.x - 1);
    if (last) *ticket = 0u;  // self-reset: workspace is reusable without a memset
    *is_last_smem = last;
  }
  __syncthreads();
  bool last = *is_last_smem != 0;
  if (last) __threadfence();
  return last;
}

// Fixed-order sum of the per-block partials (called by the last block only).
template <int K>
__device__ __forceinline__ void final_sum(double (&out)[K], const void* workspace,
                                          double* scratch) {
  const double* partials =
      reinterpret_cast<const double*>(reinterpret_cast<const char*>(workspace) + kTicketBytes);
#pragma unroll
  for (int k = 0; k < K; ++k) out[k] = 0.0;
  for (unsigned b = threadIdx.x; b < gridDim.x; b += blockDim.x) {
#pragma unroll
    for (int k = 0; k < K; ++k) out[k] += __ldcg(&partials[(size_t)b * K + k]);
  }
  __syncthreads();  // scratch may still be in use by block_sum of the caller
  block_sum<K>(out, scratch);
}

// ---------------------------------------------------------------- mbarrier / TMA PTX
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
               "r"(bytes)
               : "memory");
}
// plain arrival (consumer side of a full/empty pair: "this warp is done with the buffer")
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred P1;\n\t"
      "LAB_WAIT:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
      "@P1 bra DONE;\n\t"
      "bra LAB_WAIT;\n\t"
      "DONE:\n\t"
      "}" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
// generic-proxy writes to smem -> visible to the async proxy (TMA store source)
__device__ __forceinline__ void fence_proxy_async_smem() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
// 1-D bulk copy global -> shared, completion on an mbarrier (SASS: UBLKCP)
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gmem_src, uint32_t bytes,
                                         uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::
          "r"(smem_u32(smem_dst)),
      "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
      : "memory");
}
// 1-D bulk copy shared -> global, tracked by the thread's bulk async-group
__device__ __forceinline__ void bulk_s2g(void* gmem_dst, const void* smem_src, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gmem_dst),
               "r"(smem_u32(smem_src)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() {
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
template <int N>
__device__ __forceinline__ void bulk_wait_read() {
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
template <int N>
__device__ __forceinline__ void bulk_wait() {
  asm volatile("cp.async.bulk.wait_group %0;" ::"n"(N) : "memory");
}
// 2-D tiled TMA load / store through a CUtensorMap (SASS: UTMALDG / UTMASTG)
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const void* tmap, int c0, int c1,
                                            uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes "
      "[%0], [%1, {%2, %3}], [%4];" ::"r"(smem_u32(smem_dst)),
      "l"(tmap), "r"(c0), "r"(c1), "r"(smem_u32(bar))
      : "memory");
}
__device__ __forceinline__ void tma_store_2d(const void* tmap, const void* smem_src, int c0,
                                             int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%1, %2}], [%3];" ::
                   "l"(tmap),
               "r"(c0), "r"(c1), "r"(smem_u32(smem_src))
               : "memory");
}
// One m16 x k32 int8 A fragment = four 8-row x 16-byte matrices: lane l supplies the address of row
// l % 8 of matrix l / 8 and receives, per matrix, bytes 4 (l % 4) .. + 3 of row l / 4 — exactly the
// a0..a3 registers of mma.m16n8k32 when the matrices are (rows 0-7 | rows 8-15) x (k 0-15 | k 16-31).
__device__ __forceinline__ void ldmatrix_x4(unsigned (&a)[4], unsigned smem_addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(a[0]), "=r"(a[1]), "=r"(a[2]), "=r"(a[3])
               : "r"(smem_addr));
}

__device__ __forceinline__ void tma_prefetch_desc(const void* tmap) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(tmap) : "memory");
}
#endif  // __CUDACC__

}  // namespace derl
