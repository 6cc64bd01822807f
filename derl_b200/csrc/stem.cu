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
#include <stdlib.h>

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

// ------------------------------------------------------------------ integer tensor-core variant
// sm_100a still has INT8 tensor cores (removed on sm_103): the frames ARE uint8, so the MMA can
// consume them raw — no float conversion of the frame at all — if the weights are expressed as
// signed 8-bit digits.  Per output channel n: scale s = max|W[:, n]| / 127,
//   W ~= s * (q1 + q2 / 254),  q1 = round(W / s),  q2 = round((W - s*q1) * 254 / s),
// residual <= s / 508 (1.6e-5 of the channel's largest weight: below TF32's 2^-11 per product).
// mma.sync m16n8k32 u8 x s8 -> s32 accumulates both digit planes EXACTLY (|acc| < 2^24); the
// epilogue recombines them in fp32: y = relu((acc1 * s + acc2 * s / 254) / 255 + bias).
// One k32 step is one 32-byte tap row in natural order; a CTA of 9 warps takes one frame per
// iteration (3 m16 tiles per warp), raw frames double-buffered by TMA bulk copies.
constexpr int kI8Warps = 9, kI8Threads = kI8Warps * 32, kI8Tiles = 3, kPlanes = 2;

struct StemI8Smem {
  static constexpr size_t raw_off = 0;                                  // [2][28224] u8
  static constexpr size_t w_off = raw_off + 2 * (size_t)kImgBytes;      // [8][4][32] uint4
  static constexpr size_t scale_off = w_off + 8 * 4 * 32 * 16;          // [32] float s
  static constexpr size_t bias_off = scale_off + kOutC * 4;
  static constexpr size_t bar_off = bias_off + kOutC * 4;
  static constexpr size_t bytes = bar_off + 16;
};

__device__ __forceinline__ void mma_u8s8(int (&d)[4], const unsigned (&a)[4], unsigned b0,
                                         unsigned b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k32.row.col.s32.u8.s8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, "
      "{%0,%1,%2,%3};"
      : "+r"(d[0]), "+r"(d[1]), "+r"(d[2]), "+r"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

template <typename OT>
__global__ void __launch_bounds__(kI8Threads)
stem_conv_relu_i8_kernel(const uint8_t* __restrict__ frames, const float* __restrict__ weight,
                         const float* __restrict__ bias, OT* __restrict__ out, long long batch) {
  extern __shared__ __align__(128) uint8_t smem[];
  uint4* wsm = reinterpret_cast<uint4*>(smem + StemI8Smem::w_off);
  float* ssm = reinterpret_cast<float*>(smem + StemI8Smem::scale_off);
  float* bsm = reinterpret_cast<float*>(smem + StemI8Smem::bias_off);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + StemI8Smem::bar_off);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int g = lane >> 2, t = lane & 3;

  const long long first = blockIdx.x, stride = gridDim.x;
  if (tid == 0) {
    mbar_init(&full[0], 1);
    mbar_init(&full[1], 1);
    mbar_fence_init();
    for (int b = 0; b < 2; ++b) {
      if (first + b * stride < batch) {
        mbar_expect_tx(&full[b], kImgBytes);
        bulk_g2s(smem + StemI8Smem::raw_off + (size_t)b * kImgBytes,
                 frames + (first + b * stride) * kImgBytes, kImgBytes, &full[b]);
      }
    }
  }
  // ---- per-channel scales: warp w reduces channels w, w+9, ...
  for (int n = warp; n < kOutC; n += kI8Warps) {
    float m = 0.f;
    for (int k = lane; k < kTaps; k += 32) m = fmaxf(m, fabsf(__ldg(weight + n * kTaps + k)));
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, off));
    if (lane == 0) ssm[n] = m > 0.f ? m / 127.f : 1.f;
  }
  if (tid < kOutC) bsm[tid] = __ldg(bias + tid);
  __syncthreads();
  // ---- digit planes as B fragments: wsm[kh][nt][lane] = {p1 b0, p1 b1, p2 b0, p2 b1}
  for (int e = tid; e < 8 * 4 * 32; e += kI8Threads) {
    const int ln = e & 31, nt = (e >> 5) & 3, kh = e >> 7;
    const int tt = ln & 3, n = nt * 8 + (ln >> 2);
    const float s = ssm[n], inv = 1.f / s;
    unsigned words[4] = {0u, 0u, 0u, 0u};
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      const int kw = half * 4 + tt;                 // tap-row bytes 16*half + 4*tt + c
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const float w = __ldg(weight + ((n * kImgC + c) * 8 + kh) * 8 + kw);
        const float q1 = rintf(w * inv);
        const float q2 = fminf(fmaxf(rintf((w - q1 * s) * 254.f * inv), -127.f), 127.f);
        words[half] |= ((unsigned)(int)q1 & 0xffu) << (8 * c);
        words[2 + half] |= ((unsigned)(int)q2 & 0xffu) << (8 * c);
      }
    }
    wsm[e] = make_uint4(words[0], words[1], words[2], words[3]);
  }
  __syncthreads();

  // byte offsets (raw frame, tap row 0) of this lane's rows; tiles beyond 24 are masked
  int base0[kI8Tiles], base1[kI8Tiles];
  bool live[kI8Tiles];
#pragma unroll
  for (int m = 0; m < kI8Tiles; ++m) {
    const int tile = warp * kI8Tiles + m;
    live[m] = tile < kPix / 16;
    const int p0 = (live[m] ? tile : 0) * 16 + g, p1 = p0 + 8;
    base0[m] = (p0 / kOutHW) * 4 * (kImgW * kImgC) + (p0 % kOutHW) * 16 + 4 * t;
    base1[m] = (p1 / kOutHW) * 4 * (kImgW * kImgC) + (p1 % kOutHW) * 16 + 4 * t;
  }

  int it = 0;
  for (long long f = first; f < batch; f += stride, ++it) {
    const int buf = it & 1;
    const uint8_t* raw = smem + StemI8Smem::raw_off + (size_t)buf * kImgBytes;
    mbar_wait(&full[buf], (unsigned)((it >> 1) & 1));

    int acc[kI8Tiles][4][kPlanes][4];
#pragma unroll
    for (int m = 0; m < kI8Tiles; ++m)
#pragma unroll
      for (int n = 0; n < 4; ++n)
#pragma unroll
        for (int p = 0; p < kPlanes; ++p)
#pragma unroll
          for (int k = 0; k < 4; ++k) acc[m][n][p][k] = 0;

#pragma unroll 2
    for (int kh = 0; kh < 8; ++kh) {
      uint4 bq[4];
#pragma unroll
      for (int n = 0; n < 4; ++n) bq[n] = wsm[(kh * 4 + n) * 32 + lane];
      const int koff = kh * (kImgW * kImgC);
#pragma unroll
      for (int m = 0; m < kI8Tiles; ++m) {
        if (!live[m]) continue;  // warp-uniform
        const unsigned a[4] = {*reinterpret_cast<const unsigned*>(raw + base0[m] + koff),
                               *reinterpret_cast<const unsigned*>(raw + base1[m] + koff),
                               *reinterpret_cast<const unsigned*>(raw + base0[m] + koff + 16),
                               *reinterpret_cast<const unsigned*>(raw + base1[m] + koff + 16)};
#pragma unroll
        for (int n = 0; n < 4; ++n) {
          mma_u8s8(acc[m][n][0], a, bq[n].x, bq[n].y);
          mma_u8s8(acc[m][n][1], a, bq[n].z, bq[n].w);
        }
      }
    }
    __syncthreads();  // all warps are done with raw[buf]: refill it with the frame after next
    if (tid == 0 && f + 2 * stride < batch) {
      mbar_expect_tx(&full[buf], kImgBytes);
      bulk_g2s(smem + StemI8Smem::raw_off + (size_t)buf * kImgBytes,
               frames + (f + 2 * stride) * kImgBytes, kImgBytes, &full[buf]);
    }

    OT* dst = out + f * (long long)(kPix * kOutC);
    const float inv255 = 1.0f / 255.0f, inv254 = 1.0f / 254.0f;
#pragma unroll
    for (int m = 0; m < kI8Tiles; ++m) {
      if (!live[m]) continue;
      const int p0 = (warp * kI8Tiles + m) * 16 + g;
#pragma unroll
      for (int n = 0; n < 4; ++n) {
        const int ch = n * 8 + 2 * t;
        const float s0 = ssm[ch] * inv255, s1 = ssm[ch + 1] * inv255;
        const float b0 = bsm[ch], b1 = bsm[ch + 1];
        float y[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const float sc = (k & 1) ? s1 : s0;
          y[k] = ((float)acc[m][n][0][k] + (float)acc[m][n][1][k] * inv254) * sc +
                 ((k & 1) ? b1 : b0);
          y[k] = fmaxf(y[k], 0.f);
        }
        store2<OT>(dst + p0 * kOutC + ch, y[0], y[1]);
        store2<OT>(dst + (p0 + 8) * kOutC + ch, y[2], y[3]);
      }
    }
  }
}

template <typename OT>
int launch_i8(const uint8_t* frames, const float* weight, const float* bias, void* out,
              long long batch, cudaStream_t st) {
  auto kern = stem_conv_relu_i8_kernel<OT>;
  static bool attr_set = false;
  if (!attr_set) {
    DERL_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   (int)StemI8Smem::bytes));
    attr_set = true;
  }
  long long grid = batch;
  const long long cap = (long long)sm_count() * 3;   // 3 CTAs of ~72 KB smem per SM
  if (grid > cap) grid = cap;
  kern<<<(unsigned)grid, kI8Threads, StemI8Smem::bytes, st>>>(frames, weight, bias,
                                                              reinterpret_cast<OT*>(out), batch);
  DERL_LAUNCH_CHECK("stem_conv_relu_i8_kernel");
  return DERL_OK;
}

template <typename OT>
int launch(const uint8_t* frames, const float* weight, const float* bias, void* out,
           long long batch, cudaStream_t st) {
  static const char* variant = getenv("DERL_STEM_VARIANT");  // tuning knob: "bf16" | "i8"
  if (variant == nullptr || variant[0] == 'i') return launch_i8<OT>(frames, weight, bias, out, batch, st);
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
