// K4 — frame-stack preparation for the network stem: uint8 NHWC frames -> space-to-depth,
// scaled floating-point, channels-last tensor in ONE pass.
//
// Replaces the reference's NatureCNNBase.forward input pipeline (derl/models.py:117-123:
// permute NHWC->NCHW, `.float() / 255`, `.contiguous()` transpose copy) for the stem evaluated
// as a 2x2/stride-1 convolution over the space-to-depth(s) tensor (derl_b200/models.py):
//     dst[b, Y, X, (i*s + j)*C + c] = float(src[b, s*Y + i, s*X + j, c]) / divisor
// With s*C == 16 (Atari: s = 4, C = 4) a (j, c) run is 16 contiguous bytes on both sides:
// one thread moves one run — fully coalesced writes, 128-B coalesced read segments.  HBM-bound: 1 B read + sizeof(out) B written per element.
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <stdlib.h>

#include "common.cuh"

namespace derl {
namespace {

template <typename OT>
__device__ __forceinline__ OT to_out(float x);
template <>
__device__ __forceinline__ float to_out<float>(float x) { return x; }
template <>
__device__ __forceinline__ __nv_bfloat16 to_out<__nv_bfloat16>(float x) {
  return __float2bfloat16_rn(x);
}
template <>
__device__ __forceinline__ __half to_out<__half>(float x) { return __float2half_rn(x); }

// One thread per 16-byte (j, c) run.  The s source rows of one (b, Y) block row are contiguous
// in memory (s * WX runs) and so is their destination: the whole transform is a [s x WX] ->
// [WX x s] transpose of 16-byte runs inside each such group.  Threads enumerate runs in SOURCE
// order (kSrcOrder: one contiguous, sector-aligned read stream; each run is written as full
// 32-byte sectors) or in destination order (fully coalesced writes, strided 128-B reads).
template <typename OT, typename IT, bool kSrcOrder>
__global__ void __launch_bounds__(256)
frames_to_s2d_kernel(const uint4* __restrict__ src, OT* __restrict__ dst, IT granules, IT wx,
                     int log2s, float divisor) {
  const IT stride = (IT)gridDim.x * blockDim.x;
  const float recip = __frcp_rn(divisor);
  const IT smask = ((IT)1 << log2s) - 1;
  for (IT t0 = (IT)blockIdx.x * blockDim.x + threadIdx.x; t0 < granules; t0 += stride) {
    IT t, from;
    if (kSrcOrder) {
      const IT group = t0 / (wx << log2s), r = t0 - group * (wx << log2s);
      const IT i = r / wx, X = r - i * wx;
      from = t0;
      t = group * (wx << log2s) + (X << log2s) + i;
    } else {
      const IT i = t0 & smask, q = t0 >> log2s;
      const IT r = q / wx;             // b*HY + Y
      const IT X = q - r * wx;
      from = ((r << log2s) + i) * wx + X;
      t = t0;
    }
    const uint4 raw = __ldg(src + from);
    const unsigned words[4] = {raw.x, raw.y, raw.z, raw.w};
    alignas(16) OT vals[16];
#pragma unroll
    for (int w = 0; w < 4; ++w) {
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        // byte -> float without the conversion pipe: 0x4B0000bb is 2^23 + bb exactly
        const float f = __uint_as_float(__byte_perm(words[w], 0x4B000000u, 0x7440 + k)) - 8388608.f;
        float v = f;
        if (divisor != 1.f) {
          // correctly rounded f / divisor for the 256 byte values: reciprocal estimate plus one
          // FMA residual step (verified bit-exact against IEEE division in the tests)
          v = __fmul_rn(f, recip);
          v = __fmaf_rn(__fmaf_rn(-v, divisor, f), recip, v);
        }
        vals[w * 4 + k] = to_out<OT>(v);
      }
    }
    uint4* o = reinterpret_cast<uint4*>(dst + (size_t)t * 16);
    const uint4* v4 = reinterpret_cast<const uint4*>(vals);
#pragma unroll
    for (int c = 0; c < (int)(sizeof(OT) * 16 / 16); ++c) o[c] = v4[c];
  }
}

// Preferred variant when one group (s * WX runs) fits a CTA: the CTA stages the raw bytes of
// `gpb` whole groups through shared memory.  Global reads are one contiguous stream (exactly
// 1 B/element of DRAM traffic), the [s x WX] -> [WX x s] run transpose happens on the 16-byte
// LDS (<= 2-way bank conflicts), global writes are fully coalesced; all index arithmetic is
// loop-invariant per thread.
template <typename OT>
__global__ void __launch_bounds__(256)
frames_to_s2d_staged_kernel(const uint4* __restrict__ src, OT* __restrict__ dst,
                            long long groups, int wx, int log2s, int gpb, float divisor) {
  __shared__ uint4 stage[256];
  const int per_group = wx << log2s;
  const int active = gpb * per_group;
  const int t = threadIdx.x;
  const float recip = __frcp_rn(divisor);
  // destination-order decomposition of this thread's run inside the CTA tile
  const int g = t / per_group, r = t - g * per_group;
  const int X = r >> log2s, i = r & ((1 << log2s) - 1);
  const int from = g * per_group + i * wx + X;
  for (long long base = (long long)blockIdx.x * gpb; base < groups;
       base += (long long)gridDim.x * gpb) {
    const long long left = (groups - base) * per_group;       // runs left from this tile on
    const int live = left < active ? (int)left : active;
    const long long run0 = base * per_group;
    if (t < live) stage[t] = __ldg(src + run0 + t);
    __syncthreads();
    if (t < live) {
      const uint4 raw = stage[from];
      const unsigned words[4] = {raw.x, raw.y, raw.z, raw.w};
      alignas(16) OT vals[16];
#pragma unroll
      for (int w = 0; w < 4; ++w) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const float f =
              __uint_as_float(__byte_perm(words[w], 0x4B000000u, 0x7440 + k)) - 8388608.f;
          float v = f;
          if (divisor != 1.f) {
            v = __fmul_rn(f, recip);
            v = __fmaf_rn(__fmaf_rn(-v, divisor, f), recip, v);
          }
          vals[w * 4 + k] = to_out<OT>(v);
        }
      }
      uint4* o = reinterpret_cast<uint4*>(dst + (size_t)(run0 + t) * 16);
      const uint4* v4 = reinterpret_cast<const uint4*>(vals);
#pragma unroll
      for (int c = 0; c < (int)(sizeof(OT) * 16 / 16); ++c) o[c] = v4[c];
    }
    __syncthreads();
  }
}

// Space-to-depth / depth-to-space of a channels-last activation, 16 bytes per thread:
//   s2d[b, Y, X, (i*s + j)*C + c] = x[b, s*Y + i, s*X + j, c]      (inverse: roles swapped)
// Runs of s*C elements are contiguous on both sides.  Used (with its inverse as backward) to
// feed a strided conv whose kernel is 2 x stride as a 2x2 / stride-1 conv.
template <bool kInverse>
__global__ void __launch_bounds__(256)
space_to_depth_kernel(const uint4* __restrict__ src, uint4* __restrict__ dst, long long total,
                      int hy, int wx, int s, int cq) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  const int run = s * cq;  // 16-byte units per contiguous (j, c) run
  for (long long d = (long long)blockIdx.x * blockDim.x + threadIdx.x; d < total; d += stride) {
    // d indexes the space-to-depth tensor: ((((b*hy + Y)*wx + X)*s + i)*run + r)
    const long long q = d / run;
    const int r = (int)(d - q * run);
    const long long q2 = q / s;
    const int i = (int)(q - q2 * s);
    const long long q3 = q2 / wx;                 // b*hy + Y
    const int X = (int)(q2 - q3 * wx);
    const long long b = q3 / hy;
    const int Y = (int)(q3 - b * hy);
    const long long x = (((b * hy + Y) * s + i) * (long long)wx + X) * run + r;  // plain layout
    if (kInverse) {
      dst[x] = __ldg(src + d);
    } else {
      dst[d] = __ldg(src + x);
    }
  }
}

template <typename OT>
int launch(const void* src, void* dst, long long granules, int wx, int s, float divisor,
           cudaStream_t st) {
  long long blocks = (granules + 255) / 256;
  const long long cap = (long long)sm_count() * 64;
  if (blocks > cap) blocks = cap;
  int log2s = 0;
  while ((1 << log2s) < s) ++log2s;
  const uint4* in = reinterpret_cast<const uint4*>(src);
  OT* out = reinterpret_cast<OT*>(dst);
  static const char* variant = getenv("DERL_FRAMES_VARIANT");  // tuning knob: "dst" | "src"
  const int per_group = wx * s;
  if (variant == nullptr && per_group <= 256) {
    const int gpb = 256 / per_group;
    const long long groups = granules / per_group;
    long long grid = (groups + gpb - 1) / gpb;
    if (grid > cap) grid = cap;
    frames_to_s2d_staged_kernel<OT><<<(unsigned)grid, 256, 0, st>>>(in, out, groups, wx, log2s,
                                                                    gpb, divisor);
    DERL_LAUNCH_CHECK("frames_to_s2d_staged_kernel");
    return DERL_OK;
  }
  const bool dst_order = variant == nullptr || variant[0] == 'd';
  if (granules >= (1ll << 31)) {
    frames_to_s2d_kernel<OT, unsigned long long, true><<<(unsigned)blocks, 256, 0, st>>>(
        in, out, (unsigned long long)granules, (unsigned long long)wx, log2s, divisor);
  } else if (dst_order) {
    frames_to_s2d_kernel<OT, unsigned, false><<<(unsigned)blocks, 256, 0, st>>>(
        in, out, (unsigned)granules, (unsigned)wx, log2s, divisor);
  } else {
    frames_to_s2d_kernel<OT, unsigned, true><<<(unsigned)blocks, 256, 0, st>>>(
        in, out, (unsigned)granules, (unsigned)wx, log2s, divisor);
  }
  DERL_LAUNCH_CHECK("frames_to_s2d_kernel");
  return DERL_OK;
}

}  // namespace
}  // namespace derl

using namespace derl;

extern "C" int derl_b200_space_to_depth(const void* src, int64_t batch, int64_t height,
                                        int64_t width, int64_t channel_bytes, int64_t block,
                                        int inverse, void* dst, void* stream) {
  DERL_REQUIRE(src && dst && batch >= 0, "space_to_depth: bad arguments");
  DERL_REQUIRE(block >= 1 && height % block == 0 && width % block == 0 && height >= block &&
                   width >= block,
               "space_to_depth: height=%lld, width=%lld must be multiples of block=%lld",
               (long long)height, (long long)width, (long long)block);
  DERL_REQUIRE(channel_bytes >= 16 && channel_bytes % 16 == 0,
               "space_to_depth: channel bytes (%lld) must be a multiple of 16",
               (long long)channel_bytes);
  DERL_REQUIRE(((uintptr_t)src & 15) == 0 && ((uintptr_t)dst & 15) == 0,
               "space_to_depth: src and dst must be 16-byte aligned");
  int rc = require_device();
  if (rc != DERL_OK) return rc;
  const long long total = batch * height * width * (channel_bytes / 16);
  if (total == 0) return DERL_OK;
  long long blocks = (total + 255) / 256;
  const long long cap = (long long)sm_count() * 64;
  if (blocks > cap) blocks = cap;
  const int hy = (int)(height / block), wx = (int)(width / block), cq = (int)(channel_bytes / 16);
  cudaStream_t st = as_stream(stream);
  if (inverse) {
    space_to_depth_kernel<true><<<(unsigned)blocks, 256, 0, st>>>(
        reinterpret_cast<const uint4*>(src), reinterpret_cast<uint4*>(dst), total, hy, wx,
        (int)block, cq);
  } else {
    space_to_depth_kernel<false><<<(unsigned)blocks, 256, 0, st>>>(
        reinterpret_cast<const uint4*>(src), reinterpret_cast<uint4*>(dst), total, hy, wx,
        (int)block, cq);
  }
  DERL_LAUNCH_CHECK("space_to_depth_kernel");
  return DERL_OK;
}

extern "C" int derl_b200_frames_to_s2d(const uint8_t* src, int64_t batch, int64_t height,
                                       int64_t width, int64_t channels, int64_t block,
                                       void* dst, int dst_dtype, double divisor, void* stream) {
  DERL_REQUIRE(src && dst && batch >= 0, "frames_to_s2d: bad arguments");
  DERL_REQUIRE(block >= 1 && block * channels == 16,
               "frames_to_s2d: needs block*channels == 16 (got block=%lld channels=%lld)",
               (long long)block, (long long)channels);
  DERL_REQUIRE(height >= block && width >= block && height % block == 0 && width % block == 0,
               "frames_to_s2d: height=%lld and width=%lld must be multiples of block=%lld",
               (long long)height, (long long)width, (long long)block);
  DERL_REQUIRE(((uintptr_t)src & 15) == 0 && ((uintptr_t)dst & 15) == 0,
               "frames_to_s2d: src and dst must be 16-byte aligned");
  DERL_REQUIRE(dst_dtype >= 0 && dst_dtype <= 2, "frames_to_s2d: dst_dtype %d not in {0,1,2}",
               dst_dtype);
  int rc = require_device();
  if (rc != DERL_OK) return rc;
  if (batch == 0) return DERL_OK;
  const int wx = (int)(width / block);
  const long long granules = batch * height * wx;
  cudaStream_t st = as_stream(stream);
  switch (dst_dtype) {
    case DERL_DTYPE_BF16:
      return launch<__nv_bfloat16>(src, dst, granules, wx, (int)block, (float)divisor, st);
    case DERL_DTYPE_F16:
      return launch<__half>(src, dst, granules, wx, (int)block, (float)divisor, st);
    default:
      return launch<float>(src, dst, granules, wx, (int)block, (float)divisor, st);
  }
}
