// K6 — the NatureCNN stem on uint8 frames: conv 8x8 / stride 4 over [84,84,4] + bias + ReLU,
// reading the gathered frame stacks directly.
//
// Replaces, for the reference's first layer (derl/models.py:102-103,117-123: permute,
// `.float()/255`, `.contiguous()`, nn.Conv2d(4, 32, 8, 4), nn.ReLU), the chain
// frames_to_s2d (K4, writes a 4x larger fp32 copy of every frame) -> cuDNN fprop (reads it back):
// here a frame crosses HBM once as 28 224 bytes and only the [400 x 32] activation is written.
//
// sm_100a still has INT8 tensor cores (removed on sm_103): the frames ARE uint8, so the MMA can
// consume them raw — no float conversion of the frame at all — if the weights are expressed as
// signed 8-bit digits.  Per output channel n: scale s = max|W[:, n]| / 127,
//   W ~= s * (q1 + q2 / 254),  q1 = round(W / s),  q2 = round((W - s*q1) * 254 / s),
// residual <= s / 508 (1.6e-5 of the channel's largest weight: below TF32's 2^-11 per product).
// mma.sync m16n8k32 u8 x s8 -> s32 accumulates both digit planes EXACTLY (|acc| < 2^24); the
// epilogue recombines them in fp32: y = relu((acc1 * s + acc2 * s / 254) / 255 + bias).
//   * implicit GEMM [400 pixels x 256 taps] x [256 x 32]: one k32 step is one 32-byte tap row
//     of the raw frame in natural order, A fragments are 32-bit LDS (conflict-free);
//   * a CTA of 9 warps takes one frame per iteration (3 m16 tiles per warp); raw frames are
//     double-buffered by TMA bulk copies (cp.async.bulk + mbarrier), weights are quantised into
//     B-fragment order once per CTA;
//   * the activation is stored channels-last, optionally already in the space-to-depth(2)
//     arrangement the next layer (4x4 / stride 2 as a 2x2 / stride-1 conv) consumes.
// History (profiles/): a bf16 mma.sync version (frame converted to bf16 in smem, weights split
// hi + lo) measured 1.48 ms per 32768 frames — exactly the legacy-MMA issue limit (one
// m16n8k16 per 16 cycles per SM sub-partition); the int8 form halves the MMA count: 0.83 ms.
#include <cuda_bf16.h>
#include <stdlib.h>

#include "common.cuh"

namespace derl {
namespace {

constexpr int kImgH = 84, kImgW = 84, kImgC = 4;
constexpr int kImgBytes = kImgH * kImgW * kImgC;      // 28224
constexpr int kOutHW = 20, kOutC = 32, kPix = kOutHW * kOutHW;  // 400 output pixels
constexpr int kTaps = 256;                            // 8 x 8 x 4
__device__ __forceinline__ unsigned pack_bf16(float lo, float hi) {
  const __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<const unsigned*>(&v);
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

// frames [B,84,84,4] u8; weight [32,4,8,8] f32 (PyTorch conv layout); bias [32];
// out [B,20,20,32] channels-last, or [B,10,10,128] when out_block == 2.
// Tuning knobs (tools/tune_stem.py builds and times variants): warps x m16-tiles per warp must
// cover the 25 tiles of a frame.
#ifndef DERL_STEM_WARPS
#define DERL_STEM_WARPS 9
#endif
#ifndef DERL_STEM_TILES
#define DERL_STEM_TILES 3
#endif
#ifndef DERL_STEM_UNROLL
#define DERL_STEM_UNROLL 8
#endif
// 1: warps hand the frame buffer back through an `empty` mbarrier and run their epilogues
// independently (tensor work of some warps overlaps the stores of others); 0: one
// __syncthreads per frame keeps every warp in the same phase.
#ifndef DERL_STEM_DECOUPLED
#define DERL_STEM_DECOUPLED 1
#endif
#define DERL_STEM_PRAGMA_(x) _Pragma(#x)
#define DERL_STEM_PRAGMA(x) DERL_STEM_PRAGMA_(x)
constexpr int kI8Warps = DERL_STEM_WARPS, kI8Threads = kI8Warps * 32, kI8Tiles = DERL_STEM_TILES;
constexpr int kPlanes = 2;
static_assert(kI8Warps * kI8Tiles * 16 >= kOutHW * kOutHW, "tiles must cover the frame");

struct StemI8Smem {
  static constexpr size_t raw_off = 0;                                  // [2][28224] u8
  static constexpr size_t w_off = raw_off + 2 * (size_t)kImgBytes;      // [8][4][32] uint4
  static constexpr size_t scale_off = w_off + 8 * 4 * 32 * 16;          // [32] float s
  static constexpr size_t bias_off = scale_off + kOutC * 4;
  static constexpr size_t bar_off = bias_off + kOutC * 4;
  static constexpr size_t bytes = bar_off + 32;                         // full[2], empty[2]
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
stem_conv_relu_i8_kernel(const uint8_t* __restrict__ frames, const long long* __restrict__ rows,
                         const float* __restrict__ weight, const float* __restrict__ bias,
                         OT* __restrict__ out, long long batch, int out_block) {
  extern __shared__ __align__(128) uint8_t smem[];
  uint4* wsm = reinterpret_cast<uint4*>(smem + StemI8Smem::w_off);
  float* ssm = reinterpret_cast<float*>(smem + StemI8Smem::scale_off);
  float* bsm = reinterpret_cast<float*>(smem + StemI8Smem::bias_off);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + StemI8Smem::bar_off);
  uint64_t* empty = full + 2;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int g = lane >> 2, t = lane & 3;

  const long long first = blockIdx.x, stride = gridDim.x;
  if (tid == 0) {
    mbar_init(&full[0], 1);
    mbar_init(&full[1], 1);
    mbar_init(&empty[0], kI8Warps);
    mbar_init(&empty[1], kI8Warps);
    mbar_fence_init();
    for (int b = 0; b < 2; ++b) {
      if (first + b * stride < batch) {
        mbar_expect_tx(&full[b], kImgBytes);
        const long long i = first + b * stride;   // rows: fused minibatch gather (frame i = row rows[i])
        bulk_g2s(smem + StemI8Smem::raw_off + (size_t)b * kImgBytes,
                 frames + (rows ? __ldg(rows + i) : i) * kImgBytes, kImgBytes, &full[b]);
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

    DERL_STEM_PRAGMA(unroll DERL_STEM_UNROLL)
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
    auto refill = [&]() {   // raw[buf] <- the frame after next
      mbar_expect_tx(&full[buf], kImgBytes);
      const long long i = f + 2 * stride;
      bulk_g2s(smem + StemI8Smem::raw_off + (size_t)buf * kImgBytes,
               frames + (rows ? __ldg(rows + i) : i) * kImgBytes, kImgBytes, &full[buf]);
    };
#if DERL_STEM_DECOUPLED
    __syncwarp();
    if (lane == 0) mbar_arrive(&empty[buf]);   // this warp is done reading raw[buf]
#else
    __syncthreads();  // all warps are done with raw[buf]
    if (tid == 0 && f + 2 * stride < batch) refill();
#endif

    OT* dst = out + f * (long long)(kPix * kOutC);
#pragma unroll
    for (int m = 0; m < kI8Tiles; ++m) {
      if (!live[m]) continue;
      const int p0 = (warp * kI8Tiles + m) * 16 + g, p1 = p0 + 8;
      int o0 = p0 * kOutC, o1 = p1 * kOutC;
      if (out_block == 2) {  // pixel (oy, ox) -> ((oy/2 * 10 + ox/2) * 4 + (oy%2)*2 + ox%2) * 32
        const int y0 = p0 / kOutHW, x0 = p0 % kOutHW, y1 = p1 / kOutHW, x1 = p1 % kOutHW;
        o0 = (((y0 >> 1) * (kOutHW / 2) + (x0 >> 1)) * 4 + (y0 & 1) * 2 + (x0 & 1)) * kOutC;
        o1 = (((y1 >> 1) * (kOutHW / 2) + (x1 >> 1)) * 4 + (y1 & 1) * 2 + (x1 & 1)) * kOutC;
      }
#pragma unroll
      for (int n = 0; n < 4; ++n) {
        const int ch = n * 8 + 2 * t;
        // (q1 * 254 + q2) * s / (254 * 255) = (q1 + q2 / 254) * s / 255: the integer sum is exact
        // (< 2^31), one conversion, one rounding (the same arithmetic as K6t, stem_tc.cu)
        const float s0 = ssm[ch] / 64770.f, s1 = ssm[ch + 1] / 64770.f;
        const float b0 = bsm[ch], b1 = bsm[ch + 1];
        float y[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const float sc = (k & 1) ? s1 : s0;
          y[k] = (float)(acc[m][n][0][k] * 254 + acc[m][n][1][k]) * sc + ((k & 1) ? b1 : b0);
          y[k] = fmaxf(y[k], 0.f);
        }
        store2<OT>(dst + o0 + ch, y[0], y[1]);
        store2<OT>(dst + o1 + ch, y[2], y[3]);
      }
    }
#if DERL_STEM_DECOUPLED
    // producer: by the time warp 0 has stored its tiles the other warps have normally left the
    // MMA loop of this frame; the copy then has the whole next frame's time to land
    if (tid == 0 && f + 2 * stride < batch) {
      mbar_wait(&empty[buf], (unsigned)((it >> 1) & 1));
      refill();
    }
#endif
  }
}

template <typename OT>
int launch(const uint8_t* frames, const long long* rows, const float* weight, const float* bias,
           void* out, long long batch, int out_block, cudaStream_t st) {
  auto kern = stem_conv_relu_i8_kernel<OT>;
  if (int rc_attr = ensure_dynamic_smem(reinterpret_cast<const void*>(kern), (int)StemI8Smem::bytes)) return rc_attr;
  long long grid = batch;
  const long long cap = (long long)sm_count() * 2;
  if (grid > cap) grid = cap;
  kern<<<(unsigned)grid, kI8Threads, StemI8Smem::bytes, st>>>(
      frames, rows, weight, bias, reinterpret_cast<OT*>(out), batch, out_block);
  DERL_LAUNCH_CHECK("stem_conv_relu_i8_kernel");
  return DERL_OK;
}

}  // namespace
}  // namespace derl

namespace derl {
bool stem_tc_available();
int launch_stem_tc(const uint8_t* frames, const long long* rows, long long batch,
                   long long frames_total, const float* weight, const float* bias, float* out,
                   unsigned* mask_out, int out_block, cudaStream_t st);
}  // namespace derl

using namespace derl;

extern "C" int derl_b200_stem_conv_relu(const uint8_t* frames, const int64_t* rows, int64_t batch,
                                        const float* weight, const float* bias, void* out,
                                        int out_dtype, int out_block, void* stream) {
  DERL_REQUIRE(out_block == 1 || out_block == 2, "stem_conv_relu: out_block must be 1 or 2");
  DERL_REQUIRE(frames && weight && bias && out && batch >= 0, "stem_conv_relu: bad arguments");
  DERL_REQUIRE(out_dtype == DERL_DTYPE_F32 || out_dtype == DERL_DTYPE_BF16,
               "stem_conv_relu: out_dtype must be DERL_DTYPE_F32 or DERL_DTYPE_BF16");
  DERL_REQUIRE(((uintptr_t)frames & 15) == 0 && ((uintptr_t)out & 15) == 0,
               "stem_conv_relu: frames and out must be 16-byte aligned");
  int rc = require_device();
  if (rc != DERL_OK) return rc;
  if (batch == 0) return DERL_OK;
  cudaStream_t st = as_stream(stream);
  // float32 activations: the tcgen05 / tensor-memory kernel (stem_tc.cu; bit-identical results).
  // DERL_STEM_MMA_SYNC=1 keeps the legacy mma.sync kernel (A/B measurements, tests).
  const char* legacy = getenv("DERL_STEM_MMA_SYNC");
  if (out_dtype == DERL_DTYPE_F32 && !(legacy && legacy[0] == '1') && stem_tc_available() &&
      batch * 400 < (1ll << 31)) {
    return launch_stem_tc(frames, reinterpret_cast<const long long*>(rows), batch, 0, weight, bias,
                          reinterpret_cast<float*>(out), nullptr, out_block, st);
  }
  return out_dtype == DERL_DTYPE_BF16
             ? launch<__nv_bfloat16>(frames, reinterpret_cast<const long long*>(rows), weight, bias,
                                     out, batch, out_block, st)
             : launch<float>(frames, reinterpret_cast<const long long*>(rows), weight, bias, out,
                             batch, out_block, st);
}

extern "C" int derl_b200_stem_conv_relu_mask(const uint8_t* frames, const int64_t* rows,
                                             int64_t batch, const float* weight, const float* bias,
                                             float* out, uint32_t* relu_mask, int out_block,
                                             void* stream) {
  DERL_REQUIRE(out_block == 1 || out_block == 2, "stem_conv_relu_mask: out_block must be 1 or 2");
  DERL_REQUIRE(frames && weight && bias && out && relu_mask && batch >= 0,
               "stem_conv_relu_mask: bad arguments");
  DERL_REQUIRE(((uintptr_t)frames & 15) == 0 && ((uintptr_t)out & 127) == 0 &&
                   ((uintptr_t)relu_mask & 15) == 0,
               "stem_conv_relu_mask: frames / mask must be 16-byte, out 128-byte aligned");
  DERL_REQUIRE(batch * 400 < (1ll << 31), "stem_conv_relu_mask: batch too large");
  int rc = require_device();
  if (rc != DERL_OK) return rc;
  if (batch == 0) return DERL_OK;
  DERL_REQUIRE(stem_tc_available(), "stem_conv_relu_mask: cuTensorMapEncodeTiled is unavailable");
  return launch_stem_tc(frames, reinterpret_cast<const long long*>(rows), batch, 0, weight, bias,
                        out, relu_mask, out_block, as_stream(stream));
}
