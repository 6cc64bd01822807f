// K5 — fused ReLU backward + bias gradient for channels-last activations.
//
// In the reference's autograd graph (nn.ReLU after each nn.Conv2d, derl/models.py:102-109)
// the backward of every conv layer runs threshold_backward over the activation gradient and
// then cuDNN's convolution_backward re-reads the result once more just to sum it into the bias
// gradient.  Both are pure HBM passes over [B*H*W, C] tensors (C = 32 / 64 channels innermost
// in channels-last layout); this kernel does them in one:
//     grad_pre[r, c] = out[r, c] > 0 ? grad_out[r, c] : 0          (written once)
//     bias_grad[c]   = sum_r grad_pre[r, c]                        (float32 accumulate)
// 3 x sizeof(T) bytes per element instead of 4 x, and one launch instead of two.  The
// cross-block sum uses per-block partials reduced by the last block in a fixed order
// (deterministic, no float atomics).
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include "common.cuh"

namespace derl {
namespace {

constexpr int kThreads = 256;
constexpr int kMaxBlocks = 1184;  // 148 SMs x 8

template <typename T>
struct Vec4;
template <>
struct Vec4<float> {
  using type = float4;
  static __device__ __forceinline__ void unpack(const float4& v, float (&f)[4]) {
    f[0] = v.x; f[1] = v.y; f[2] = v.z; f[3] = v.w;
  }
  static __device__ __forceinline__ float4 pack(const float (&f)[4]) {
    return make_float4(f[0], f[1], f[2], f[3]);
  }
};
template <>
struct Vec4<__nv_bfloat16> {
  using type = uint2;
  static __device__ __forceinline__ void unpack(const uint2& v, float (&f)[4]) {
    const __nv_bfloat162 a = *reinterpret_cast<const __nv_bfloat162*>(&v.x);
    const __nv_bfloat162 b = *reinterpret_cast<const __nv_bfloat162*>(&v.y);
    f[0] = __low2float(a); f[1] = __high2float(a); f[2] = __low2float(b); f[3] = __high2float(b);
  }
  static __device__ __forceinline__ uint2 pack(const float (&f)[4]) {
    const __nv_bfloat162 a = __floats2bfloat162_rn(f[0], f[1]);
    const __nv_bfloat162 b = __floats2bfloat162_rn(f[2], f[3]);
    uint2 v;
    v.x = *reinterpret_cast<const unsigned*>(&a);
    v.y = *reinterpret_cast<const unsigned*>(&b);
    return v;
  }
};
template <>
struct Vec4<__half> {
  using type = uint2;
  static __device__ __forceinline__ void unpack(const uint2& v, float (&f)[4]) {
    const __half2 a = *reinterpret_cast<const __half2*>(&v.x);
    const __half2 b = *reinterpret_cast<const __half2*>(&v.y);
    f[0] = __low2float(a); f[1] = __high2float(a); f[2] = __low2float(b); f[3] = __high2float(b);
  }
  static __device__ __forceinline__ uint2 pack(const float (&f)[4]) {
    const __half2 a = __floats2half2_rn(f[0], f[1]);
    const __half2 b = __floats2half2_rn(f[2], f[3]);
    uint2 v;
    v.x = *reinterpret_cast<const unsigned*>(&a);
    v.y = *reinterpret_cast<const unsigned*>(&b);
    return v;
  }
};

// quads = C / 4 divides kThreads; thread t owns channel quad (t % quads) of rows t / quads + k * rpb.
// unblock > 1: the inputs are a space-to-depth(unblock) arrangement [B, hy, wx, unblock^2 * c]
// of an activation [B, hy*unblock, wx*unblock, c]; grad_pre is written in the PLAIN layout
// (the depth-to-space pass is folded into the store addresses).
template <typename T>
__global__ void __launch_bounds__(kThreads)
relu_bwd_bias_kernel(const T* __restrict__ grad_out, const T* __restrict__ out,
                     T* __restrict__ grad_pre, float* __restrict__ bias_grad, long long rows,
                     int quads, float* __restrict__ partials, unsigned* __restrict__ ticket,
                     int unblock, int hy, int wx) {
  using V = typename Vec4<T>::type;
  __shared__ float4 red[kThreads];
  __shared__ int is_last;
  const int quad = threadIdx.x % quads;
  const int rpb = kThreads / quads;  // rows per block pass
  const V* g = reinterpret_cast<const V*>(grad_out);
  const V* o = reinterpret_cast<const V*>(out);
  V* p = reinterpret_cast<V*>(grad_pre);
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll 4   // four iterations' loads in flight per thread (the pointers are __restrict__)
  for (long long r = (long long)blockIdx.x * rpb + threadIdx.x / quads; r < rows;
       r += (long long)gridDim.x * rpb) {
    const long long at = r * quads + quad;
    float gv[4], ov[4];
    Vec4<T>::unpack(__ldg(g + at), gv);
    Vec4<T>::unpack(__ldg(o + at), ov);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      gv[k] = ov[k] > 0.f ? gv[k] : 0.f;
      acc[k] += gv[k];
    }
    long long to = at;
    if (unblock > 1) {
      const int cq = quads / (unblock * unblock);       // quads per plain pixel
      const int blk = quad / cq, c4 = quad - blk * cq;  // blk = i * unblock + j
      const int i = blk / unblock, j = blk - i * unblock;
      const long long by = r / wx;                      // b * hy + Y
      const int X = (int)(r - by * wx);
      const long long pixel = ((by * unblock + i) * wx + X) * unblock + j;
      to = pixel * cq + c4;
    }
    p[to] = Vec4<T>::pack(gv);
  }
  red[threadIdx.x] = make_float4(acc[0], acc[1], acc[2], acc[3]);
  __syncthreads();
  if (threadIdx.x < quads) {  // fixed-order sum over the block's row slots
    float4 s = red[threadIdx.x];
    for (int k = 1; k < rpb; ++k) {
      const float4 v = red[threadIdx.x + k * quads];
      s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
    }
    reinterpret_cast<float4*>(partials)[(size_t)blockIdx.x * quads + threadIdx.x] = s;
  }
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned t = atomicAdd(ticket, 1u);
    is_last = (t == gridDim.x - 1);
    if (is_last) *ticket = 0u;
  }
  __syncthreads();
  if (!is_last) return;
  __threadfence();
  // last block: every thread sums a strided subset of the per-block partials (slot k takes blocks
  // k, k + rpb, ...: many independent loads in flight), then a fixed-order sum over the slots
  __shared__ double wide[kThreads][4];
  {
    double s[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll 8
    for (unsigned b = threadIdx.x / quads; b < gridDim.x; b += rpb) {
      const float4 v = __ldcg(reinterpret_cast<const float4*>(partials) + (size_t)b * quads + quad);
      s[0] += v.x; s[1] += v.y; s[2] += v.z; s[3] += v.w;
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) wide[threadIdx.x][k] = s[k];
  }
  __syncthreads();
  if (threadIdx.x < quads) {
    double s[4] = {0.0, 0.0, 0.0, 0.0};
    for (int k = 0; k < rpb; ++k) {
#pragma unroll
      for (int c = 0; c < 4; ++c) s[c] += wide[threadIdx.x + k * quads][c];
    }
    reinterpret_cast<float4*>(bias_grad)[threadIdx.x] =
        make_float4((float)s[0], (float)s[1], (float)s[2], (float)s[3]);
  }
}

template <typename T>
int launch(const void* grad_out, const void* out, void* grad_pre, float* bias_grad,
           long long rows, int channels, void* workspace, int unblock, int hy, int wx,
           cudaStream_t st) {
  const int quads = channels / 4;
  const int rpb = kThreads / quads;
  long long blocks = (rows + rpb - 1) / rpb;
  // exactly one resident wave: the grid-stride loop balances itself, a partial second wave (8 CTAs
  // per SM requested, 5 resident at 46 registers) left a 20-40 % tail
  int resident = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&resident, relu_bwd_bias_kernel<T>, kThreads,
                                                    0) != cudaSuccess || resident < 1) {
    cudaGetLastError();
    resident = 4;
  }
  long long wave = (long long)sm_count() * resident;
  if (wave > kMaxBlocks) wave = kMaxBlocks;
  if (blocks > wave) blocks = wave;
  unsigned* ticket = reinterpret_cast<unsigned*>(workspace);
  float* partials = reinterpret_cast<float*>(reinterpret_cast<char*>(workspace) + kTicketBytes);
  DERL_CUDA(cudaMemsetAsync(workspace, 0, kTicketBytes, st));
  relu_bwd_bias_kernel<T><<<(unsigned)blocks, kThreads, 0, st>>>(
      reinterpret_cast<const T*>(grad_out), reinterpret_cast<const T*>(out),
      reinterpret_cast<T*>(grad_pre), bias_grad, rows, quads, partials, ticket, unblock, hy, wx);
  DERL_LAUNCH_CHECK("relu_bwd_bias_kernel");
  return DERL_OK;
}

}  // namespace
}  // namespace derl

using namespace derl;

extern "C" size_t derl_b200_relu_bwd_bias_workspace_bytes(int64_t channels) {
  if (channels < 4) channels = 4;
  return kTicketBytes + (size_t)kMaxBlocks * (size_t)channels * sizeof(float);
}

extern "C" int derl_b200_relu_bwd_bias(const void* grad_out, const void* out, void* grad_pre,
                                       float* bias_grad, int64_t rows, int64_t channels,
                                       int dtype, int unblock, int64_t blocked_height,
                                       int64_t blocked_width, void* workspace,
                                       size_t workspace_bytes, void* stream) {
  DERL_REQUIRE(unblock >= 1, "relu_bwd_bias: unblock must be >= 1");
  if (unblock > 1) {
    DERL_REQUIRE(blocked_height >= 1 && blocked_width >= 1 &&
                     rows % (blocked_height * blocked_width) == 0 &&
                     channels % (4 * unblock * unblock) == 0,
                 "relu_bwd_bias: rows=%lld / channels=%lld do not match a space-to-depth(%d) "
                 "tensor of %lld x %lld blocks", (long long)rows, (long long)channels, unblock,
                 (long long)blocked_height, (long long)blocked_width);
    DERL_REQUIRE(grad_pre != grad_out && grad_pre != out,
                 "relu_bwd_bias: the unblocking store cannot run in place");
  }
  DERL_REQUIRE(grad_out && out && grad_pre && bias_grad && workspace,
               "relu_bwd_bias: null pointer");
  DERL_REQUIRE(rows >= 1, "relu_bwd_bias: rows must be >= 1");
  DERL_REQUIRE(channels >= 4 && channels % 4 == 0 && kThreads % (channels / 4) == 0,
               "relu_bwd_bias: channels=%lld must be a multiple of 4 with (channels/4) dividing %d",
               (long long)channels, kThreads);
  DERL_REQUIRE(dtype >= 0 && dtype <= 2, "relu_bwd_bias: dtype %d not in {0,1,2}", dtype);
  DERL_REQUIRE((((uintptr_t)grad_out | (uintptr_t)out | (uintptr_t)grad_pre |
                 (uintptr_t)bias_grad) & 15) == 0, "relu_bwd_bias: pointers must be 16-byte aligned");
  if (workspace_bytes < derl_b200_relu_bwd_bias_workspace_bytes(channels)) {
    set_error("relu_bwd_bias: workspace %zu B too small", workspace_bytes);
    return DERL_E_WORKSPACE;
  }
  int rc = require_device();
  if (rc != DERL_OK) return rc;
  cudaStream_t st = as_stream(stream);
  switch (dtype) {
    case DERL_DTYPE_BF16:
      return launch<__nv_bfloat16>(grad_out, out, grad_pre, bias_grad, rows, (int)channels,
                                   workspace, unblock, (int)blocked_height, (int)blocked_width, st);
    case DERL_DTYPE_F16:
      return launch<__half>(grad_out, out, grad_pre, bias_grad, rows, (int)channels, workspace,
                            unblock, (int)blocked_height, (int)blocked_width, st);
    default:
      return launch<float>(grad_out, out, grad_pre, bias_grad, rows, (int)channels, workspace,
                           unblock, (int)blocked_height, (int)blocked_width, st);
  }
}
