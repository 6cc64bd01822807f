// K6 — the NatureCNN stem on uint8 frames: conv 8x8 / stride 4 over [84,84,4] + bias + ReLU,
// reading the gathered frame stacks directly.
//
// Replaces, for the reference's first layer (derl/models.py:102-103,117-123: permute,
// `.float()/255`, `.contiguous()`, nn.Conv2d(4, 32, 8, 4), nn.ReLU), the chain
// frames_to_s2d (K4, writes a 4x larger fp32 copy of every frame) -> cuDNN fprop (reads it back):
// here a frame crosses HBM once as 28 224 bytes and only the [400 x 32] activation is written.
//
//   * one frame per warp-group iteration: TMA bulk copy (cp.async.bulk + mbarrier) lands the
//     raw 28 224-byte frame in shared memory; the group converts it once to bf16 (byte values
//     0..255 are exact in bf16) while the next frame's copy is already in flight;
//   * implicit GEMM [400 pixels x 256 taps] x [256 x 32] on the tensor cores
//     (mma.sync m16n8k16 bf16, fp32 accumulate): each of the 5 warps of a group owns 5 m16
//     tiles; A fragments are 64-bit LDS straight from the bf16 frame (the K order inside a
//     16-tap step is permuted so that a thread's two k-pairs are 8 contiguous bytes; the
//     weights are stored with the same permutation), conflict-free;
//   * weights are split W = hi + lo into two bf16 planes (two MMAs per step), i.e. ~16 mantissa
//     bits — more than the TF32 (11 bits) cuDNN path this replaces; the 1/255 of the
//     reference's input scaling is applied to the fp32 accumulator in the epilogue;
//   * epilogue: acc/255 + bias, ReLU, full-sector stores of the channels-last activation.
#include <cuda_bf16.h>

#include "common.cuh"

namespace derl {
namespace {

constexpr int kImgH = 84, kImgW = 84, kImgC = 4;
constexpr int kImgBytes = kImgH * kImgW * kImgC;      // 28224
constexpr int kRowBf16 = kImgW * kImgC * 2;           // 672 bytes per bf16 image row
constexpr int kOutHW = 20, kOutC = 32, kPix = kOutHW * kOutHW;  // 400 output pixels
constexpr int kTaps = 256;                            // 8 x 8 x 4
constexpr int kSteps = kTaps / 16;                    // 16 k16 steps (2 per kernel row)
constexpr int kGroupWarps = 5, kGroupThreads = kGroupWarps * 32;  // 5 warps x 5 m16 tiles = 400
constexpr int kGroups = 2, kThreads = kGroups * kGroupThreads;
constexpr int kTilesPerWarp = 5;

struct StemSmem {
  static constexpr size_t raw_off = 0;                                   // [2][28224] u8
  static constexpr size_t img_off = raw_off + (size_t)kGroups * kImgBytes;   // [2][28224] bf16
  static constexpr size_t w_off = img_off + (size_t)kGroups * kImgBytes * 2; // 8192 words
  static constexpr size_t bias_off = w_off + (size_t)kSteps * 4 * 32 * 4 * 4;
  static constexpr size_t bar_off = bias_off + kOutC * 4;
  static constexpr size_t bytes = bar_off + 8 * kGroups;
};

__device__ __forceinline__ void group_sync(int group) {
  asm volatile("bar.sync %0, %1;" ::"r"(group + 1), "n"(kGroupThreads) : "memory");
}

__device__ __forceinline__ unsigned pack_bf16(float lo, float hi) {
  const __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<const unsigned*>(&v);
}

__device__ __forceinline__ void mma_bf16(float (&d)[4], const unsigned (&a)[4], unsigned b0,
                                         unsigned b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, "
      "{%8,%9}, {%0,%1,%2,%3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

template <typename OT>
__device__ __forceinline__ void store2(OT* p, float x, float y);
template <>
__device__ __forceinline__ void store2<float>(float* p, float x, float y) {
  *reinterpret_cast<float2*>(p) = make_float2(x, y);
}
template <>
__device__ __forceinline__ void store2<__nv_bfloat16>(__nv_bfloat16* p, float x, float y) {
  *reinterpret_cast<unsigned*>(p) = pack_bf16(x, y);
}

// frames [B,84,84,4] u8; weight [32,4,8,8] f32 (PyTorch conv layout); bias [32]; out [B,20,20,32]
template <typename OT>
__global__ void __launch_bounds__(kThreads, 1)
stem_conv_relu_kernel(const uint8_t* __restrict__ frames, const float* __restrict__ weight,
                      const float* __restrict__ bias, OT* __restrict__ out, long long batch) {
  extern __shared__ __align__(128) uint8_t smem[];
  unsigned* wsm = reinterpret_cast<unsigned*>(smem + StemSmem::w_off);
  float* bsm = reinterpret_cast<float*>(smem + StemSmem::bias_off);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + StemSmem::bar_off);

  const int tid = threadIdx.x;
  const int group = tid / kGroupThreads, gtid = tid - group * kGroupThreads;
  const int warp = gtid >> 5, lane = gtid & 31;
  const int g = lane >> 2, t = lane & 3;
  uint8_t* raw = smem + StemSmem::raw_off + (size_t)group * kImgBytes;
  uint8_t* img = smem + StemSmem::img_off + (size_t)group * kImgBytes * 2;

  // frames of this group: first, first + stride, ...
  const long long first = (long long)blockIdx.x * kGroups + group;
  const long long stride = (long long)gridDim.x * kGroups;
  if (gtid == 0) {
    mbar_init(&full[group], 1);
    mbar_fence_init();
    if (first < batch) {
      mbar_expect_tx(&full[group], kImgBytes);
      bulk_g2s(raw, frames + first * kImgBytes, kImgBytes, &full[group]);
    }
  }

  // ---- weights: [32,4,8,8] f32 -> per-step, per-lane B fragments, split into bf16 hi + lo
  for (int e = tid; e < kSteps * 32 * 16; e += kThreads) {
    const int w = e & 15, ln = (e >> 4) & 31, s = e >> 9;
    const int nt = w >> 2, hl = (w >> 1) & 1, r = w & 1;
    const int tt = ln & 3, gg = ln >> 2;
    const int n = nt * 8 + gg, kh = s >> 1;
    const int kb = (s & 1) * 16 + 4 * tt + 2 * r;   // byte pair (kb, kb+1) of the 32-byte tap row
    const int kw = kb >> 2, c = kb & 3;
    const float v0 = __ldg(weight + ((n * kImgC + c) * 8 + kh) * 8 + kw);
    const float v1 = __ldg(weight + ((n * kImgC + c + 1) * 8 + kh) * 8 + kw);
    const float h0 = __bfloat162float(__float2bfloat16_rn(v0));
    const float h1 = __bfloat162float(__float2bfloat16_rn(v1));
    wsm[((s * 4 + (w >> 2)) * 32 + ln) * 4 + (w & 3)] =
        hl ? pack_bf16(v0 - h0, v1 - h1) : pack_bf16(h0, h1);
  }
  if (tid < kOutC) bsm[tid] = __ldg(bias + tid);
  __syncthreads();

  // byte offsets (into the bf16 frame, kernel row 0) of this lane's two rows of each m16 tile
  int base0[kTilesPerWarp], base1[kTilesPerWarp];
#pragma unroll
  for (int m = 0; m < kTilesPerWarp; ++m) {
    const int p0 = (warp * kTilesPerWarp + m) * 16 + g, p1 = p0 + 8;
    base0[m] = (p0 / kOutHW) * 4 * kRowBf16 + (p0 % kOutHW) * 32 + 8 * t;
    base1[m] = (p1 / kOutHW) * 4 * kRowBf16 + (p1 % kOutHW) * 32 + 8 * t;
  }

  unsigned parity = 0;
  for (long long f = first; f < batch; f += stride, parity ^= 1) {
    mbar_wait(&full[group], parity);
    // ---- raw bytes -> bf16 frame (exact), one 32-bit word (4 bytes) per thread step
    for (int wi = gtid; wi < kImgBytes / 4; wi += kGroupThreads) {
      const unsigned word = reinterpret_cast<const unsigned*>(raw)[wi];
      float v[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        v[k] = __uint_as_float(__byte_perm(word, 0x4B000000u, 0x7440 + k)) - 8388608.f;
      }
      reinterpret_cast<uint2*>(img)[wi] = make_uint2(pack_bf16(v[0], v[1]), pack_bf16(v[2], v[3]));
    }
    group_sync(group);  // bf16 frame complete; raw buffer free again
    if (gtid == 0 && f + stride < batch) {
      mbar_expect_tx(&full[group], kImgBytes);
      bulk_g2s(raw, frames + (f + stride) * kImgBytes, kImgBytes, &full[group]);
    }

    float acc[kTilesPerWarp][4][4];
#pragma unroll
    for (int m = 0; m < kTilesPerWarp; ++m)
#pragma unroll
      for (int n = 0; n < 4; ++n)
#pragma unroll
        for (int k = 0; k < 4; ++k) acc[m][n][k] = 0.f;

#pragma unroll 1
    for (int s = 0; s < kSteps; ++s) {
      uint4 bq[4];  // bq[nt] = {hi b0, hi b1, lo b0, lo b1}
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        bq[q] = reinterpret_cast<const uint4*>(wsm)[(s * 4 + q) * 32 + lane];
      }
      const int koff = (s >> 1) * kRowBf16 + (s & 1) * 32;
#pragma unroll
      for (int m = 0; m < kTilesPerWarp; ++m) {
        const uint2 r0 = *reinterpret_cast<const uint2*>(img + base0[m] + koff);
        const uint2 r1 = *reinterpret_cast<const uint2*>(img + base1[m] + koff);
        const unsigned a[4] = {r0.x, r1.x, r0.y, r1.y};
#pragma unroll
        for (int n = 0; n < 4; ++n) mma_bf16(acc[m][n], a, bq[n].x, bq[n].y);
#pragma unroll
        for (int n = 0; n < 4; ++n) mma_bf16(acc[m][n], a, bq[n].z, bq[n].w);
      }
    }

    // ---- epilogue: /255, + bias, ReLU, channels-last store
    OT* dst = out + f * (long long)(kPix * kOutC);
    const float inv255 = 1.0f / 255.0f;
#pragma unroll
    for (int m = 0; m < kTilesPerWarp; ++m) {
      const int p0 = (warp * kTilesPerWarp + m) * 16 + g;
#pragma unroll
      for (int n = 0; n < 4; ++n) {
        const int ch = n * 8 + 2 * t;
        const float b0 = bsm[ch], b1 = bsm[ch + 1];
        store2<OT>(dst + p0 * kOutC + ch, fmaxf(fmaf(acc[m][n][0], inv255, b0), 0.f),
                   fmaxf(fmaf(acc[m][n][1], inv255, b1), 0.f));
        store2<OT>(dst + (p0 + 8) * kOutC + ch, fmaxf(fmaf(acc[m][n][2], inv255, b0), 0.f),
                   fmaxf(fmaf(acc[m][n][3], inv255, b1), 0.f));
      }
    }
    group_sync(group);  // every warp is done reading the bf16 frame before it is overwritten
  }
}

template <typename OT>
int launch(const uint8_t* frames, const float* weight, const float* bias, void* out,
           long long batch, cudaStream_t st) {
  auto kern = stem_conv_relu_kernel<OT>;
  static bool attr_set = false;
  if (!attr_set) {
    DERL_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   (int)StemSmem::bytes));
    attr_set = true;
  }
  long long grid = (batch + kGroups - 1) / kGroups;
  if (grid > sm_count()) grid = sm_count();
  kern<<<(unsigned)grid, kThreads, StemSmem::bytes, st>>>(frames, weight, bias,
                                                           reinterpret_cast<OT*>(out), batch);
  DERL_LAUNCH_CHECK("stem_conv_relu_kernel");
  return DERL_OK;
}

}  // namespace
}  // namespace derl

using namespace derl;

extern "C" int derl_b200_stem_conv_relu(const uint8_t* frames, int64_t batch, const float* weight,
                                        const float* bias, void* out, int out_dtype,
                                        void* stream) {
  DERL_REQUIRE(frames && weight && bias && out && batch >= 0, "stem_conv_relu: bad arguments");
  DERL_REQUIRE(out_dtype == DERL_DTYPE_F32 || out_dtype == DERL_DTYPE_BF16,
               "stem_conv_relu: out_dtype must be DERL_DTYPE_F32 or DERL_DTYPE_BF16");
  DERL_REQUIRE(((uintptr_t)frames & 15) == 0 && ((uintptr_t)out & 15) == 0,
               "stem_conv_relu: frames and out must be 16-byte aligned");
  int rc = require_device();
  if (rc != DERL_OK) return rc;
  if (batch == 0) return DERL_OK;
  cudaStream_t st = as_stream(stream);
  return out_dtype == DERL_DTYPE_BF16
             ? launch<__nv_bfloat16>(frames, weight, bias, out, batch, st)
             : launch<float>(frames, weight, bias, out, batch, st);
}
