// K9 — the linear output heads of the actor-critic network in one pass over the hidden features.
//
// The reference's NatureCNNModel.forward applies one nn.Linear(512, n) per output to the trunk's
// 512 features (`outputs = [layer(base_outputs) for layer in self.output_layers]`,
// derl/models.py:201-202; for PPO: logits [B, A] and the value [B, 1]).  As library calls that is,
// per forward and backward pass, two GEMVs over the same [B, 512] tensor, two input-gradient GEMMs
// plus the add that joins them, two weight-gradient GEMMs with split-K, two bias-gradient
// reductions and — because the trunk's last nn.Linear(3136, 512) (derl/models.py:112-114) has a
// bias — one more full reduction of the [B, 512] gradient: ~15 small launches that each re-read
// the features or their gradient.  With the heads stacked into one [U, 512] matrix (U = sum of
// the output widths, <= 32):
//     forward    out[r, u]  = (h[r, :] + hb) . W[u, :] + b[u]                 one read of h
//     backward   dh[r, :]   = sum_u g[r, u] W[u, :]                           one write of dh
//                dW[u, :]   = sum_r g[r, u] (h[r, :] + hb),  db[u] = sum_r g[r, u]   one read of h
//                dhb[:]     = sum_u db[u] W[u, :]
// where hb is the trunk's bias when the caller defers it (h then is the bias-free product): the
// bias gradient of the 3136 -> 512 layer is the heads' column sums times W, 512 x U flops
// instead of a reduction over B x 512 values.  float32 FMA arithmetic (no TF32), fixed summation
// order: per-CTA partials, then a fixed-order sum over the CTAs.
#include "common.cuh"

namespace derl {
namespace {

constexpr int kF = 512;           // features of the trunk (derl/models.py:114)
constexpr int kF4 = kF / 4;       // float4 per row
constexpr int kMaxUnits = 32;
constexpr int kThreads = 256;
constexpr int kChunk = 8;         // outputs whose weight gradient one CTA accumulates
constexpr int kMaxCtas = 512;     // partials in the workspace (grid.x <= this)
constexpr int kFwdRows = 4;       // rows in flight per warp
constexpr int kBwdRows = 2;

__device__ __forceinline__ float dot4(const float4& a, const float4& b, float acc) {
  acc = fmaf(a.x, b.x, acc);
  acc = fmaf(a.y, b.y, acc);
  acc = fmaf(a.z, b.z, acc);
  return fmaf(a.w, b.w, acc);
}

__device__ __forceinline__ void axpy4(float s, const float4& x, float4& y) {
  y.x = fmaf(s, x.x, y.x);
  y.y = fmaf(s, x.y, y.y);
  y.z = fmaf(s, x.z, y.z);
  y.w = fmaf(s, x.w, y.w);
}

// ------------------------------------------------------------------------------------ forward
// one warp per row, lane l owns columns 4 (l + 32 j) .. + 3, j = 0..3; W staged in shared memory
__global__ void __launch_bounds__(kThreads)
heads_fwd_kernel(const float4* __restrict__ hidden, const float4* __restrict__ hidden_bias,
                 const float4* __restrict__ weight, const float* __restrict__ bias,
                 float* __restrict__ out, long long batch, int units) {
  extern __shared__ float4 w_sm[];              // [units][128]
  __shared__ float shift[kMaxUnits];            // b[u] + W[u, :] . hb
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < units * kF4; i += kThreads) w_sm[i] = __ldg(weight + i);
  __syncthreads();
  for (int u = warp; u < units; u += kThreads / 32) {
    float acc = 0.f;
    if (hidden_bias != nullptr) {
#pragma unroll
      for (int j = 0; j < 4; ++j)
        acc = dot4(__ldg(hidden_bias + j * 32 + lane), w_sm[u * kF4 + j * 32 + lane], acc);
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
    }
    if (lane == 0) shift[u] = acc + __ldg(bias + u);
  }
  __syncthreads();

  const long long stride = (long long)gridDim.x * (kThreads / 32);
  for (long long base = (long long)blockIdx.x * (kThreads / 32) + warp; base < batch;
       base += stride * kFwdRows) {
    float4 h[kFwdRows][4];
#pragma unroll
    for (int k = 0; k < kFwdRows; ++k) {
      const long long r = base + k * stride;
#pragma unroll
      for (int j = 0; j < 4; ++j)
        h[k][j] = r < batch ? __ldcs(hidden + r * kF4 + j * 32 + lane) : make_float4(0, 0, 0, 0);
    }
    float mine[kFwdRows];
#pragma unroll
    for (int k = 0; k < kFwdRows; ++k) mine[k] = 0.f;
    for (int u = 0; u < units; ++u) {
      float acc[kFwdRows];
#pragma unroll
      for (int k = 0; k < kFwdRows; ++k) acc[k] = 0.f;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float4 w = w_sm[u * kF4 + j * 32 + lane];
#pragma unroll
        for (int k = 0; k < kFwdRows; ++k) acc[k] = dot4(h[k][j], w, acc[k]);
      }
#pragma unroll
      for (int k = 0; k < kFwdRows; ++k) {
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) acc[k] += __shfl_xor_sync(0xffffffffu, acc[k], off);
        if (lane == u) mine[k] = acc[k] + shift[u];
      }
    }
#pragma unroll
    for (int k = 0; k < kFwdRows; ++k) {
      const long long r = base + k * stride;
      if (r < batch && lane < units) out[r * units + lane] = mine[k];
    }
  }
}

// ----------------------------------------------------------------------------------- backward
// CTA (x, y): y = chunk of 8 outputs whose dW / db it accumulates; chunk 0 also writes dh.
// 8 warps = 4 row slots x 2 column halves; lane l of half c owns float4 columns c * 64 + l + 32 j.
__global__ void __launch_bounds__(kThreads, 2)
heads_bwd_kernel(const float4* __restrict__ hidden, const float4* __restrict__ weight,
                 const float* __restrict__ grad_out, float4* __restrict__ grad_hidden,
                 float4* __restrict__ partial_w, float* __restrict__ partial_c, long long batch,
                 int units) {
  extern __shared__ float4 w_sm[];              // chunk 0: [units][128]; then the slot reduction
  __shared__ float csum_sm[4][kChunk];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int slot = warp >> 1, half = warp & 1;
  const int u_lo = blockIdx.y * kChunk;
  const bool lead = blockIdx.y == 0;
  if (lead) {
    for (int i = threadIdx.x; i < units * kF4; i += kThreads) w_sm[i] = __ldg(weight + i);
  }
  __syncthreads();

  float4 acc[kChunk][2];
  float csum[kChunk];
#pragma unroll
  for (int u = 0; u < kChunk; ++u) {
    acc[u][0] = acc[u][1] = make_float4(0, 0, 0, 0);
    csum[u] = 0.f;
  }
  const int col = half * 64 + lane;             // + 32 j
  const long long stride = (long long)gridDim.x * 4;
  for (long long base = (long long)blockIdx.x * 4 + slot; base < batch;
       base += stride * kBwdRows) {
    float4 h[kBwdRows][2];
#pragma unroll
    for (int k = 0; k < kBwdRows; ++k) {
      const long long r = base + k * stride;
#pragma unroll
      for (int j = 0; j < 2; ++j)
        h[k][j] = r < batch ? __ldcs(hidden + r * kF4 + col + 32 * j) : make_float4(0, 0, 0, 0);
    }
#pragma unroll
    for (int k = 0; k < kBwdRows; ++k) {
      const long long r = base + k * stride;
      if (r >= batch) break;
      const float* g = grad_out + r * units;
      if (lead) {                                // dh = g W over ALL outputs
        float4 d0 = make_float4(0, 0, 0, 0), d1 = d0;
        for (int u = 0; u < units; ++u) {
          const float gu = __ldg(g + u);
          axpy4(gu, w_sm[u * kF4 + col], d0);
          axpy4(gu, w_sm[u * kF4 + col + 32], d1);
        }
        __stcs(grad_hidden + r * kF4 + col, d0);
        __stcs(grad_hidden + r * kF4 + col + 32, d1);
      }
#pragma unroll
      for (int u = 0; u < kChunk; ++u) {         // dW, db of this CTA's chunk
        const float gu = u_lo + u < units ? __ldg(g + u_lo + u) : 0.f;
        axpy4(gu, h[k][0], acc[u][0]);
        axpy4(gu, h[k][1], acc[u][1]);
        csum[u] += gu;
      }
    }
  }

  // fixed-order sum over the 4 row slots through shared memory, then one partial per CTA
  __syncthreads();                               // every warp is done with w_sm
  float4* red = w_sm;                            // [kChunk][128]
  for (int s = 0; s < 4; ++s) {
    if (slot == s) {
#pragma unroll
      for (int u = 0; u < kChunk; ++u) {
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          float4 v = acc[u][j];
          if (s > 0) {
            const float4 p = red[u * kF4 + col + 32 * j];
            v.x += p.x; v.y += p.y; v.z += p.z; v.w += p.w;
          }
          red[u * kF4 + col + 32 * j] = v;
        }
      }
      if (half == 0 && lane == 0) {
#pragma unroll
        for (int u = 0; u < kChunk; ++u) csum_sm[s][u] = csum[u];
      }
    }
    __syncthreads();
  }
  const size_t cta = (size_t)blockIdx.y * gridDim.x + blockIdx.x;
  for (int i = threadIdx.x; i < kChunk * kF4; i += kThreads) partial_w[cta * kChunk * kF4 + i] = red[i];
  if (threadIdx.x < kChunk) {
    const int u = threadIdx.x;
    partial_c[cta * kChunk + u] = ((csum_sm[0][u] + csum_sm[1][u]) + csum_sm[2][u]) + csum_sm[3][u];
  }
}

// block u: dW[u, :] and db[u] = fixed-order sums of the per-CTA partials (+ db[u] hb for dW when the
// trunk's bias was deferred).  1024 threads = 8 groups x 128 float4 columns.
__global__ void __launch_bounds__(1024)
heads_bwd_reduce_kernel(const float4* __restrict__ partial_w, const float* __restrict__ partial_c,
                        const float4* __restrict__ hidden_bias, float4* __restrict__ grad_weight,
                        float* __restrict__ grad_bias, int ctas, int units) {
  __shared__ float4 part[8][kF4];
  __shared__ float cpart[1024];
  const int u = blockIdx.x, chunk = u / kChunk, slot = u % kChunk;
  const int c4 = threadIdx.x & (kF4 - 1), group = threadIdx.x >> 7;
  const float4* pw = partial_w + ((size_t)chunk * ctas * kChunk + slot) * kF4 + c4;
  float4 s = make_float4(0, 0, 0, 0);
#pragma unroll 4
  for (int b = group; b < ctas; b += 8) {
    const float4 v = __ldcg(pw + (size_t)b * kChunk * kF4);
    s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
  }
  part[group][c4] = s;
  float c = 0.f;
  for (int b = threadIdx.x; b < ctas; b += 1024)
    c += __ldcg(partial_c + ((size_t)chunk * ctas + b) * kChunk + slot);
  cpart[threadIdx.x] = c;
  __syncthreads();
  for (int half = 512; half > 0; half >>= 1) {   // fixed tree
    if (threadIdx.x < half) cpart[threadIdx.x] += cpart[threadIdx.x + half];
    __syncthreads();
  }
  const float colsum = cpart[0];
  if (threadIdx.x == 0) grad_bias[u] = colsum;
  if (group == 0) {
    float4 t = part[0][c4];
#pragma unroll
    for (int g = 1; g < 8; ++g) {
      const float4 v = part[g][c4];
      t.x += v.x; t.y += v.y; t.z += v.z; t.w += v.w;
    }
    if (hidden_bias != nullptr) axpy4(colsum, __ldg(hidden_bias + c4), t);
    grad_weight[(size_t)u * kF4 + c4] = t;
  }
}

// dhb[c] = sum_u db[u] W[u, c]
__global__ void __launch_bounds__(kF)
heads_bwd_bias_kernel(const float* __restrict__ weight, const float* __restrict__ grad_bias,
                      float* __restrict__ grad_hidden_bias, int units) {
  const int c = threadIdx.x;
  float s = 0.f;
  for (int u = 0; u < units; ++u) s = fmaf(__ldg(grad_bias + u), __ldg(weight + (size_t)u * kF + c), s);
  grad_hidden_bias[c] = s;
}

size_t partial_w_bytes(int units) {
  const int chunks = (units + kChunk - 1) / kChunk;
  return (size_t)chunks * kMaxCtas * kChunk * kF * sizeof(float);
}

bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

}  // namespace
}  // namespace derl

using namespace derl;

extern "C" size_t derl_b200_linear_heads_workspace_bytes(int units) {
  if (units < 1) units = 1;
  if (units > kMaxUnits) units = kMaxUnits;
  const int chunks = (units + kChunk - 1) / kChunk;
  return partial_w_bytes(units) + (size_t)chunks * kMaxCtas * kChunk * sizeof(float);
}

extern "C" int derl_b200_linear_heads_forward(const float* hidden, const float* hidden_bias,
                                              const float* weight, const float* bias, float* out,
                                              int64_t batch, int features, int units,
                                              void* stream) {
  DERL_REQUIRE(features == kF, "linear_heads: %d features, this kernel is built for %d", features,
               kF);
  DERL_REQUIRE(units >= 1 && units <= kMaxUnits, "linear_heads: 1..%d output units, got %d",
               kMaxUnits, units);
  DERL_REQUIRE(batch >= 0, "linear_heads: negative batch");
  DERL_REQUIRE(weight && bias && (batch == 0 || (hidden && out)), "linear_heads: null pointer");
  DERL_REQUIRE(aligned16(hidden) && aligned16(weight) && aligned16(hidden_bias),
               "linear_heads: hidden / weight / hidden_bias must be 16-byte aligned");
  if (int rc = require_device()) return rc;
  if (batch == 0) return DERL_OK;
  const int smem = units * kF * (int)sizeof(float);
  if (int rc = ensure_dynamic_smem(reinterpret_cast<const void*>(heads_fwd_kernel), smem)) return rc;
  long long grid = (batch + (kThreads / 32) * kFwdRows - 1) / ((kThreads / 32) * kFwdRows);
  const long long wave = (long long)sm_count() * 2;
  if (grid > wave) grid = wave;
  heads_fwd_kernel<<<(unsigned)grid, kThreads, smem, as_stream(stream)>>>(
      reinterpret_cast<const float4*>(hidden), reinterpret_cast<const float4*>(hidden_bias),
      reinterpret_cast<const float4*>(weight), bias, out, batch, units);
  DERL_LAUNCH_CHECK("heads_fwd_kernel");
  return DERL_OK;
}

extern "C" int derl_b200_linear_heads_backward(const float* hidden, const float* hidden_bias,
                                               const float* weight, const float* grad_out,
                                               float* grad_hidden, float* grad_weight,
                                               float* grad_bias, float* grad_hidden_bias,
                                               int64_t batch, int features, int units,
                                               void* workspace, size_t workspace_bytes,
                                               void* stream) {
  DERL_REQUIRE(features == kF, "linear_heads: %d features, this kernel is built for %d", features,
               kF);
  DERL_REQUIRE(units >= 1 && units <= kMaxUnits, "linear_heads: 1..%d output units, got %d",
               kMaxUnits, units);
  DERL_REQUIRE(batch >= 1, "linear_heads_backward: batch must be >= 1");
  DERL_REQUIRE(hidden && weight && grad_out && grad_hidden && grad_weight && grad_bias && workspace,
               "linear_heads_backward: null pointer");
  DERL_REQUIRE((hidden_bias == nullptr) == (grad_hidden_bias == nullptr),
               "linear_heads_backward: hidden_bias and grad_hidden_bias go together");
  DERL_REQUIRE(aligned16(hidden) && aligned16(weight) && aligned16(hidden_bias) &&
                   aligned16(grad_hidden) && aligned16(grad_weight) && aligned16(workspace),
               "linear_heads_backward: tensors must be 16-byte aligned");
  DERL_REQUIRE(workspace_bytes >= derl_b200_linear_heads_workspace_bytes(units),
               "linear_heads_backward: workspace too small");
  if (int rc = require_device()) return rc;
  const int chunks = (units + kChunk - 1) / kChunk;
  int smem = units * kF * (int)sizeof(float);
  if (smem < kChunk * kF * (int)sizeof(float)) smem = kChunk * kF * (int)sizeof(float);
  if (int rc = ensure_dynamic_smem(reinterpret_cast<const void*>(heads_bwd_kernel), smem)) return rc;
  long long grid = (batch + 4 * kBwdRows - 1) / (4 * kBwdRows);
  long long wave = (long long)sm_count() * 2;
  if (wave > kMaxCtas) wave = kMaxCtas;
  if (grid > wave) grid = wave;
  float4* partial_w = reinterpret_cast<float4*>(workspace);
  float* partial_c = reinterpret_cast<float*>(reinterpret_cast<char*>(workspace) +
                                              partial_w_bytes(units));
  cudaStream_t st = as_stream(stream);
  heads_bwd_kernel<<<dim3((unsigned)grid, (unsigned)chunks), kThreads, smem, st>>>(
      reinterpret_cast<const float4*>(hidden), reinterpret_cast<const float4*>(weight), grad_out,
      reinterpret_cast<float4*>(grad_hidden), partial_w, partial_c, batch, units);
  DERL_LAUNCH_CHECK("heads_bwd_kernel");
  heads_bwd_reduce_kernel<<<units, 1024, 0, st>>>(
      partial_w, partial_c, reinterpret_cast<const float4*>(hidden_bias),
      reinterpret_cast<float4*>(grad_weight), grad_bias, (int)grid, units);
  DERL_LAUNCH_CHECK("heads_bwd_reduce_kernel");
  if (grad_hidden_bias != nullptr) {
    heads_bwd_bias_kernel<<<1, kF, 0, st>>>(weight, grad_bias, grad_hidden_bias, units);
    DERL_LAUNCH_CHECK("heads_bwd_bias_kernel");
  }
  return DERL_OK;
}
