// K7 — backward of the NatureCNN stem, straight from the uint8 frames on the INT8 tensor cores.
//
// Replaces, for the reference's first layer (derl/models.py:102-103: nn.Conv2d(4, 32, 8, 4) +
// nn.ReLU), the three-pass backward K5 (ReLU mask + bias gradient) -> K4 (re-create the float32
// space-to-depth frames) -> cuDNN wgrad with ONE kernel that reads, per frame, the raw 28 224
// bytes, the incoming gradient and the saved activation (for the ReLU mask) exactly once:
//     g[p, n]     = out[p, n] > 0 ? grad_out[p, n] : 0                       (400 pixels x 32)
//     dbias[n]   += sum_p g[p, n]
//     dW[n, c, 4a+i, 4b+j] += (1/255) * sum_{oy,ox} x[4(oy+a)+i, 4(ox+b)+j, c] * g[(oy,ox), n]
// The frame operand is uint8 and exact; the gradient is quantised per (frame, channel) to two
// signed 8-bit digit planes (block floating point): s = max_p |g[p, n]| / 127,
// g ~= s * (q1 + q2 / 254), residual <= s / 508.  mma.sync m16n8k32 s8 x u8 -> s32 accumulates
// each frame exactly; the int32 tile is then scaled and added to fp32 running sums.
//   GEMM view per frame and per kernel quadrant (a, b):  D[32 ch x 64 taps] += Gq^T [32 x 400]
//   * Z_ab [400 x 64], reduction over the output pixels.  MMA packs 4 consecutive reduction
//   indices per register, so both operands are laid out pixel-fastest in shared memory:
//   Gq[plane][group of 16 pixels][channel][16 digits] (written by the quantiser) and
//   Zt[c*16 + i*4 + j][Y][X], a byte transpose of the space-to-depth(4) frame (4x4 byte blocks,
//   PRMT), in which 4 consecutive output pixels of one row are 4 consecutive bytes.  Both layouts
//   are bank-conflict free for their writer and for the MMA fragment loads.
//   Two CTAs (8 warps each) per SM loop over frames; each CTA runs its phases (transpose,
//   mask + quantise, MMA) back to back and the SM overlaps one CTA's ALU phases with the other's
//   tensor phase.  The frame arrives by a TMA bulk copy; the gradient and activation tiles are
//   read with coalesced streaming loads straight into registers (one channel per lane), issued
//   before the transpose so that they land while it runs.  Warp w owns kernel rows 4 (w / 4) .. + 3
//   and frame channel w % 4 for both column halves (which share their Zt words, one byte apart).
// Per-CTA partial sums are reduced (fixed order) and re-indexed to [32, 4, 8, 8] by a second
// small kernel.
#include "common.cuh"

namespace derl {
namespace {

constexpr int kFrameBytes = 84 * 84 * 4;      // 28224
constexpr int kPix = 400, kCh = 32;           // output pixels, output channels
constexpr int kQuads = 100;                   // 4 consecutive ox of one oy (order: see the tables)
constexpr int kSteps = 13;                    // ceil(100 quads / 8 quads per k32 step)
constexpr int kGroups = 26;                   // groups of 4 quads (16 pixels); the 26th is zero padding
constexpr int kGqPlane = kGroups * kCh * 16;  // Gq[plane][group][channel][4 quads x 4 digits]
// Zt[plane][Y (21)][X (24, 21 used)], plane = c * 16 + (i * 4 + j).  A pitch of 130 words puts plane k
// at bank 2k: the transpose's stores (16 (i, j) x 2 X-blocks per warp) and the MMA's B loads (8 planes
// of equal c and equal (i, j) parity x 4 words) both touch every bank once.
constexpr int kZtRow = 24, kZtTap = 520;
constexpr int kWarps = 8, kThreads = kWarps * 32, kCtasPerSm = 2;
constexpr int kPartial = 4 * 64 * kCh;        // floats of one CTA's weight partial

struct BwdSmem {
  static constexpr size_t frame_off = 0;                              // [28224] u8 (+64: the
  static constexpr size_t zt_off = frame_off + kFrameBytes + 64;      //  transpose over-reads)
  static constexpr size_t gq_off = zt_off + 64 * (size_t)kZtTap;      // s8 [2][26][32][16]
  static constexpr size_t red_off = gq_off + 2 * (size_t)kGqPlane;    // float [kWarps][32]
  static constexpr size_t scale_off = red_off + kWarps * kCh * 4;     // float [32]
  static constexpr size_t qtile_off = scale_off + kCh * 4;            // int [104]: quad -> tile offset
  static constexpr size_t qzt_off = qtile_off + 104 * 4;              // int [104]: quad -> Zt offset
  static constexpr size_t bar_off = qzt_off + 104 * 4;                // 1 mbarrier
  static constexpr size_t bytes = bar_off + 16;
};
static_assert(BwdSmem::zt_off % 16 == 0 && BwdSmem::gq_off % 16 == 0, "smem alignment");
static_assert(kCtasPerSm * (BwdSmem::bytes + 1024) <= 228 * 1024, "two CTAs must fit one SM");

__device__ __forceinline__ void mma_s8u8(int (&d)[4], const unsigned (&a)[4], unsigned b0,
                                         unsigned b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k32.row.col.s32.s8.u8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, "
      "{%0,%1,%2,%3};"
      : "+r"(d[0]), "+r"(d[1]), "+r"(d[2]), "+r"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// index of output pixel (oy, ox) inside a frame's [400 x 32] tile, in units of pixels
__device__ __forceinline__ int tile_pixel(int oy, int ox, int blocked) {
  if (blocked) return (((oy >> 1) * 10 + (ox >> 1)) << 2) + ((oy & 1) << 1) + (ox & 1);
  return oy * 20 + ox;
}

// Two-digit quantisation of x in [-127, 127]: q1 = rint(x), q2 = rint((x - q1) * 254) (|q2| <= 127
// because |x - q1| <= 1/2 exactly).  Rounding by the 1.5 * 2^23 trick keeps the conversion pipe
// idle; the digits are the low bytes of the biased sums.
__device__ __forceinline__ void quantise(float x, unsigned* b1, unsigned* b2) {
  const float m1 = x + 12582912.f;
  *b1 = __float_as_uint(m1);
  *b2 = __float_as_uint(__fmaf_rn(x - (m1 - 12582912.f), 254.f, 12582912.f));
}

// low bytes of four words -> one word
__device__ __forceinline__ unsigned pack4(unsigned a, unsigned b, unsigned c, unsigned d) {
  return __byte_perm(__byte_perm(a, b, 0x0040), __byte_perm(c, d, 0x0040), 0x5410);
}

// grad_out / out: [B, 400, 32] float32 tiles (plain pixel order, or space-to-depth(2) order when
// blocked != 0); partial_w: [grid][4 quadrants][64 taps][32 ch]; partial_b: [grid][32].
__global__ void __launch_bounds__(kThreads, kCtasPerSm)
stem_bwd_kernel(const uint8_t* __restrict__ frames, const long long* __restrict__ rows,
                const float* __restrict__ grad_out, const float* __restrict__ out,
                float* __restrict__ partial_w, float* __restrict__ partial_b, long long batch,
                int blocked) {
  extern __shared__ __align__(128) uint8_t smem[];
  uint8_t* raw = smem + BwdSmem::frame_off;
  uint8_t* zt = smem + BwdSmem::zt_off;
  uint8_t* gq = smem + BwdSmem::gq_off;
  float* red = reinterpret_cast<float*>(smem + BwdSmem::red_off);
  float* scale = reinterpret_cast<float*>(smem + BwdSmem::scale_off);
  int* qtile = reinterpret_cast<int*>(smem + BwdSmem::qtile_off);
  int* qzt = reinterpret_cast<int*>(smem + BwdSmem::qzt_off);
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + BwdSmem::bar_off);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int g = lane >> 2, t = lane & 3;
  const long long first = blockIdx.x, stride = gridDim.x;

  auto load_frame = [&](long long f) {   // rows: fused minibatch gather (frame f = row rows[f])
    mbar_expect_tx(bar, kFrameBytes);
    bulk_g2s(raw, frames + (rows ? __ldg(rows + f) : f) * kFrameBytes, kFrameBytes, bar);
  };
  if (tid == 0) {
    mbar_init(bar, 1);
    mbar_fence_init();
    if (first < batch) load_frame(first);
  }
  // zero the quantised-gradient buffer once: its padding (quads 100..103) must stay zero
  for (int i = tid; i < 2 * kGqPlane / 4; i += kThreads) {
    reinterpret_cast<unsigned*>(gq)[i] = 0u;
  }
  // quad q = 4 consecutive ox of one row.  The reduction order is ours to choose: quads 0..79 are
  // (oy = q / 4, ox0 = 4 (q % 4)), quads 80..99 the fifth quad (ox0 = 16) of row q - 80, so that the
  // 4 quads one B-fragment load touches (4 consecutive q) are 4 consecutive words of one Zt row.
  // Tables: float offset of the quad's first pixel inside a tile, byte offset inside a Zt plane
  // (quads >= 100 are padding: A is zero there)
  if (tid < 104) {
    const int q = tid < kQuads ? tid : kQuads - 1;
    const int oy = q < 80 ? q >> 2 : q - 80, ox0 = q < 80 ? (q & 3) * 4 : 16;
    qtile[tid] = tile_pixel(oy, ox0, blocked) * kCh;
    qzt[tid] = oy * kZtRow + ox0;
  }
  // float offsets of a quad's 4 pixels relative to its first one
  const int step[4] = {0, kCh, (blocked ? 4 : 2) * kCh, (blocked ? 5 : 3) * kCh};
  __syncthreads();

  // ---- loop-invariant roles
  // MMA role: warp = (qa, c): kernel rows 4 qa .. 4 qa + 3, frame channel c, BOTH column halves
  // qb = 0, 1 (they read the same Zt words, one byte apart) and both (i, j) parities:
  // acc[m][n][qb] is n-tile nt = 2 c + n (channel nt >> 1, (i, j) parity nt & 1) of quadrant (qa, qb)
  const int qa = warp >> 2;
  const int ntile0 = (warp & 3) * 2;
  // A fragments by ldmatrix: lane -> (matrix = lane / 8, row = lane % 8); matrices 0/1 are channels
  // +0..7 / +8..15 of group 2s, matrices 2/3 the same channels of group 2s + 1 (16-byte rows of Gq)
  const unsigned gq_lane = (unsigned)__cvta_generic_to_shared(gq) +
                           (((lane >> 4) * kCh + ((lane >> 3) & 1) * 8 + (lane & 7)) * 16);
  // mask / quantiser role: lane = channel; warp `sub` owns groups sub, sub + 8, sub + 16 (12 quads)
  // and, for sub < 4, quad 96 + sub of the last group
  const int ch = lane, sub = warp;
  const bool extra = sub < 4;

  float wsum[2][2][2][4];                         // [m-tile][n-tile][qb][c-frag] fp32 running sums
#pragma unroll
  for (int m = 0; m < 2; ++m)
#pragma unroll
    for (int n = 0; n < 2; ++n)
#pragma unroll
      for (int q = 0; q < 2; ++q)
#pragma unroll
        for (int k = 0; k < 4; ++k) wsum[m][n][q][k] = 0.f;
  float bsum = 0.f;

  int it = 0;
  for (long long f = first; f < batch; f += stride, ++it) {
    const float* gt = grad_out + f * (kPix * kCh) + ch;
    const float* ot = out + f * (kPix * kCh) + ch;

    // ---- (0a) issue this thread's activation loads (each warp-load is one pixel: 128 contiguous B)
    float v[3][16], vx[4];
#pragma unroll
    for (int k = 0; k < 3; ++k)
#pragma unroll
      for (int qq = 0; qq < 4; ++qq) {
        const float* px = ot + qtile[4 * (sub + 8 * k) + qq];
#pragma unroll
        for (int e = 0; e < 4; ++e) v[k][4 * qq + e] = __ldcs(px + step[e]);
      }
    {
      const float* px = ot + qtile[96 + (sub & 3)];
#pragma unroll
      for (int e = 0; e < 4; ++e) vx[e] = extra ? __ldcs(px + step[e]) : 0.f;
    }

    // ---- (1) byte transpose of the frame: Zt[c*16+i*4+j][Y][X] = raw[4Y+i][4X+j][c]
    // one 4x4 byte block per lane and step: 4 source words (X..X+3, channels c=0..3 each) ->
    // 4 destination words (planes c=0..3, bytes X..X+3).
    // lane = j + 4*(xg & 1) + 8*i makes the 32 source words of a warp hit 32 distinct banks.
    // Done in two halves, each covering the latency of one set of global loads.
    auto transpose_blocks = [&](int blk_begin, int blk_end) {
      for (int blk = blk_begin + warp; blk < blk_end; blk += kWarps) {
        const int Y = blk / 3, xg = (blk - Y * 3) * 2 + ((lane >> 2) & 1);
        const int i = lane >> 3, j = lane & 3, ij = i * 4 + j;
        const uint8_t* src = raw + (4 * Y + i) * 336 + (16 * xg + j) * 4;
        unsigned w0 = *reinterpret_cast<const unsigned*>(src);
        unsigned w1 = *reinterpret_cast<const unsigned*>(src + 16);
        unsigned w2 = *reinterpret_cast<const unsigned*>(src + 32);
        unsigned w3 = *reinterpret_cast<const unsigned*>(src + 48);
        // 4x4 byte transpose (rows w0..w3, columns = channel bytes)
        const unsigned t0 = __byte_perm(w0, w1, 0x5140), t1 = __byte_perm(w0, w1, 0x7362);
        const unsigned t2 = __byte_perm(w2, w3, 0x5140), t3 = __byte_perm(w2, w3, 0x7362);
        const unsigned c0 = __byte_perm(t0, t2, 0x5410), c1 = __byte_perm(t0, t2, 0x7632);
        const unsigned c2 = __byte_perm(t1, t3, 0x5410), c3 = __byte_perm(t1, t3, 0x7632);
        uint8_t* dst = zt + ij * kZtTap + Y * kZtRow + 4 * xg;
        *reinterpret_cast<unsigned*>(dst) = c0;
        *reinterpret_cast<unsigned*>(dst + 16 * kZtTap) = c1;
        *reinterpret_cast<unsigned*>(dst + 32 * kZtTap) = c2;
        *reinterpret_cast<unsigned*>(dst + 48 * kZtTap) = c3;
      }
    };
    mbar_wait(bar, (unsigned)(it & 1));
    transpose_blocks(0, 32);

    // ---- (0b) activations -> ReLU mask bits; issue the gradient loads into the same registers
    unsigned keep[3] = {0u, 0u, 0u}, keepx = 0u;
#pragma unroll
    for (int k = 0; k < 3; ++k)
#pragma unroll
      for (int e = 0; e < 16; ++e) keep[k] |= (v[k][e] > 0.f ? 1u : 0u) << e;
#pragma unroll
    for (int e = 0; e < 4; ++e) keepx |= (vx[e] > 0.f ? 1u : 0u) << e;
#pragma unroll
    for (int k = 0; k < 3; ++k)
#pragma unroll
      for (int qq = 0; qq < 4; ++qq) {
        const float* px = gt + qtile[4 * (sub + 8 * k) + qq];
#pragma unroll
        for (int e = 0; e < 4; ++e) v[k][4 * qq + e] = __ldcs(px + step[e]);
      }
    {
      const float* px = gt + qtile[96 + (sub & 3)];
#pragma unroll
      for (int e = 0; e < 4; ++e) vx[e] = extra ? __ldcs(px + step[e]) : 0.f;
    }
    transpose_blocks(32, 21 * 3);

    // ---- (2) apply the mask; per-channel max and bias sum; values stay in registers
    float vmax = 0.f, vsum = 0.f;
#pragma unroll
    for (int k = 0; k < 3; ++k)
#pragma unroll
      for (int e = 0; e < 16; ++e) {
        const float gv = (keep[k] >> e) & 1u ? v[k][e] : 0.f;
        v[k][e] = gv;
        vmax = fmaxf(vmax, fabsf(gv));
        vsum += gv;
      }
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float gv = (keepx >> e) & 1u ? vx[e] : 0.f;
      vx[e] = gv;
      vmax = fmaxf(vmax, fabsf(gv));
      vsum += gv;
    }
    bsum += vsum;
    red[sub * kCh + ch] = vmax;
    __syncthreads();   // maxima published; every warp is done with the raw frame
    if (tid == 0 && f + stride < batch) load_frame(f + stride);

    // ---- (3) quantise from registers: one 16-byte store per group and digit plane
    {
      float m = red[ch];
#pragma unroll
      for (int k = 1; k < kWarps; ++k) m = fmaxf(m, red[k * kCh + ch]);
      const float s = m > 0.f ? m / 127.f : 1.f, inv = 1.f / s;
      if (sub == 0) scale[ch] = s;
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        unsigned w1[4], w2[4];
#pragma unroll
        for (int qq = 0; qq < 4; ++qq) {
          unsigned b1[4], b2[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) quantise(v[k][4 * qq + e] * inv, &b1[e], &b2[e]);
          w1[qq] = pack4(b1[0], b1[1], b1[2], b1[3]);
          w2[qq] = pack4(b2[0], b2[1], b2[2], b2[3]);
        }
        uint8_t* at = gq + ((sub + 8 * k) * kCh + ch) * 16;
        *reinterpret_cast<uint4*>(at) = make_uint4(w1[0], w1[1], w1[2], w1[3]);
        *reinterpret_cast<uint4*>(at + kGqPlane) = make_uint4(w2[0], w2[1], w2[2], w2[3]);
      }
      if (extra) {
        unsigned b1[4], b2[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) quantise(vx[e] * inv, &b1[e], &b2[e]);
        uint8_t* at = gq + (24 * kCh + ch) * 16 + 4 * sub;
        *reinterpret_cast<unsigned*>(at) = pack4(b1[0], b1[1], b1[2], b1[3]);
        *reinterpret_cast<unsigned*>(at + kGqPlane) = pack4(b2[0], b2[1], b2[2], b2[3]);
      }
    }
    __syncthreads();   // Zt, Gq and scale complete

    // ---- (4) MMAs: acc[m][n][qb][plane] over 13 k32 steps (8 quads each)
    int acc[2][2][2][2][4];
#pragma unroll
    for (int m = 0; m < 2; ++m)
#pragma unroll
      for (int n = 0; n < 2; ++n)
#pragma unroll
        for (int q = 0; q < 2; ++q)
#pragma unroll
          for (int p = 0; p < 2; ++p)
#pragma unroll
            for (int k = 0; k < 4; ++k) acc[m][n][q][p][k] = 0;

#pragma unroll 1
    for (int s = 0; s < kSteps; ++s) {
      const int quad0 = 8 * s + t, quad1 = quad0 + 4;       // this lane's two quads of the step
      // byte offsets of the two quads inside a Zt plane, shifted by this warp's kernel-row half
      const int off0 = qzt[quad0] + qa * kZtRow, off1 = qzt[quad1] + qa * kZtRow;
      unsigned b[2][2][2];                                   // [n][qb][b0, b1]
#pragma unroll
      for (int n = 0; n < 2; ++n) {
        const int nt = ntile0 + n;
        const uint8_t* plane = zt + ((nt >> 1) * 16 + 2 * g + (nt & 1)) * kZtTap;
        const unsigned* p0 = reinterpret_cast<const unsigned*>(plane + off0);
        const unsigned* p1 = reinterpret_cast<const unsigned*>(plane + off1);
        const unsigned w00 = p0[0], w01 = p0[1], w10 = p1[0], w11 = p1[1];
        b[n][0][0] = w00;                                    // qb = 0: the aligned word
        b[n][0][1] = w10;
        b[n][1][0] = __byte_perm(w00, w01, 0x4321);          // qb = 1: bytes 1..4
        b[n][1][1] = __byte_perm(w10, w11, 0x4321);
      }
#pragma unroll
      for (int m = 0; m < 2; ++m)
#pragma unroll
        for (int p = 0; p < 2; ++p) {
          // quads 8s + t and 8s + 4 + t live in groups 2s and 2s + 1 at word t
          unsigned a[4];
          ldmatrix_x4(a, gq_lane + p * kGqPlane + ((2 * s) * kCh + 16 * m) * 16);
#pragma unroll
          for (int n = 0; n < 2; ++n)
#pragma unroll
            for (int q = 0; q < 2; ++q) mma_s8u8(acc[m][n][q][p], a, b[n][q][0], b[n][q][1]);
        }
    }

    // ---- (5) int32 -> fp32, apply the frame's per-channel scales
#pragma unroll
    for (int m = 0; m < 2; ++m) {
      const float s_lo = scale[16 * m + g], s_hi = scale[16 * m + g + 8];
#pragma unroll
      for (int n = 0; n < 2; ++n)
#pragma unroll
        for (int q = 0; q < 2; ++q)
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const float x = (float)acc[m][n][q][0][k] + (float)acc[m][n][q][1][k] * (1.f / 254.f);
            wsum[m][n][q][k] += x * (k < 2 ? s_lo : s_hi);
          }
    }
    __syncthreads();   // everyone is done with Zt / Gq / scale / red before the next frame
  }

  // ---- per-CTA partials: partial_w[cta][quadrant][tap][channel], partial_b[cta][channel]
#pragma unroll
  for (int q = 0; q < 2; ++q) {
    float* pw = partial_w + (size_t)blockIdx.x * kPartial + (size_t)(qa * 2 + q) * (64 * kCh);
#pragma unroll
    for (int m = 0; m < 2; ++m)
#pragma unroll
      for (int n = 0; n < 2; ++n) {
        // D column col (= 2t, 2t + 1) of n-tile nt is (i, j) = 2 col + (nt & 1), channel nt >> 1
        const int nt = ntile0 + n, c_lo = 16 * m + g;
        const int tap = ((4 * t + (nt & 1)) << 2) + (nt >> 1);   // (i*4+j)*4 + c of column 2t
        pw[tap * kCh + c_lo] = wsum[m][n][q][0];
        pw[(tap + 8) * kCh + c_lo] = wsum[m][n][q][1];
        pw[tap * kCh + c_lo + 8] = wsum[m][n][q][2];
        pw[(tap + 8) * kCh + c_lo + 8] = wsum[m][n][q][3];
      }
  }
  __syncthreads();
  red[sub * kCh + ch] = bsum;
  __syncthreads();
  if (tid < kCh) {
    float b = 0.f;
#pragma unroll
    for (int k = 0; k < kWarps; ++k) b += red[k * kCh + tid];
    partial_b[(size_t)blockIdx.x * kCh + tid] = b;
  }
}

// grad_w[n][c][4a+i][4b+j] = (1/255) * sum_cta partial_w[cta][a*2+b][(i*4+j)*4+c][n]
// Blocks 0 .. kPartial/32 - 1 take 32 consecutive elements of the partial layout (one 128-byte
// line per CTA partial); the block's 8 warps sum interleaved subsets of the CTAs and the subsets
// are then added in fixed order.  The last block sums the bias partials the same way.
__global__ void __launch_bounds__(256)
stem_bwd_reduce_kernel(const float* __restrict__ partial_w, const float* __restrict__ partial_b,
                       int ctas, float* __restrict__ grad_w, float* __restrict__ grad_b) {
  constexpr int kSlots = 8;
  __shared__ double slot[kSlots][32];
  const int lane = threadIdx.x & 31, k = threadIdx.x >> 5;
  const bool bias = blockIdx.x == kPartial / 32;
  const float* src = bias ? partial_b + lane : partial_w + (size_t)blockIdx.x * 32 + lane;
  const size_t pitch = bias ? kCh : kPartial;
  double s = 0.0;
#pragma unroll 4
  for (int b = k; b < ctas; b += kSlots) s += (double)__ldg(src + (size_t)b * pitch);
  slot[k][lane] = s;
  __syncthreads();
  if (threadIdx.x < 32) {
    double total = 0.0;
#pragma unroll
    for (int q = 0; q < kSlots; ++q) total += slot[q][lane];
    if (bias) {
      grad_b[lane] = (float)total;
    } else {
      const int e = blockIdx.x * 32 + lane;   // over [4 quadrants][64 taps][32 ch]
      const int n = e & 31, tap = (e >> 5) & 63, quadrant = e >> 11;
      const int c = tap & 3, j = (tap >> 2) & 3, i = tap >> 4;
      const int kh = 4 * (quadrant >> 1) + i, kw = 4 * (quadrant & 1) + j;
      grad_w[((n * 4 + c) * 8 + kh) * 8 + kw] = (float)(total * (1.0 / 255.0));
    }
  }
}

}  // namespace

int launch_stem_bwd_reduce(const float* partial_w, const float* partial_b, int ctas, float* grad_w,
                           float* grad_b, cudaStream_t st) {
  stem_bwd_reduce_kernel<<<kPartial / 32 + 1, 256, 0, st>>>(partial_w, partial_b, ctas, grad_w,
                                                           grad_b);
  DERL_LAUNCH_CHECK("stem_bwd_reduce_kernel");
  return DERL_OK;
}

bool stem_tc_available();
int launch_stem_bwd_tc(const uint8_t* frames, const long long* rows, long long batch,
                       const float* grad_out, const unsigned* mask, int blocked, float* grad_w,
                       float* grad_b, void* workspace, cudaStream_t st);
}  // namespace derl

using namespace derl;

extern "C" size_t derl_b200_stem_backward_workspace_bytes(void) {
  return (size_t)kCtasPerSm * sm_count() * (kPartial + kCh) * sizeof(float);
}

extern "C" int derl_b200_stem_backward(const uint8_t* frames, const int64_t* rows, int64_t batch,
                                       const float* grad_out, const float* out, int blocked,
                                       float* grad_weight,
                                       float* grad_bias, void* workspace, size_t workspace_bytes,
                                       void* stream) {
  DERL_REQUIRE(frames && grad_out && out && grad_weight && grad_bias && workspace && batch >= 1,
               "stem_backward: bad arguments");
  DERL_REQUIRE(blocked == 0 || blocked == 1, "stem_backward: blocked must be 0 or 1");
  DERL_REQUIRE((((uintptr_t)frames | (uintptr_t)grad_out | (uintptr_t)out) & 15) == 0,
               "stem_backward: inputs must be 16-byte aligned");
  int rc = require_device();
  if (rc != DERL_OK) return rc;
  if (workspace_bytes < derl_b200_stem_backward_workspace_bytes()) {
    set_error("stem_backward: workspace %zu B too small", workspace_bytes);
    return DERL_E_WORKSPACE;
  }
  cudaStream_t st = as_stream(stream);
  if (int rc_attr = ensure_dynamic_smem(reinterpret_cast<const void*>(stem_bwd_kernel), (int)BwdSmem::bytes)) return rc_attr;
  const long long max_grid = (long long)kCtasPerSm * sm_count();
  long long grid = batch < max_grid ? batch : max_grid;
  float* partial_w = reinterpret_cast<float*>(workspace);
  float* partial_b = partial_w + (size_t)max_grid * kPartial;
  stem_bwd_kernel<<<(unsigned)grid, kThreads, BwdSmem::bytes, st>>>(
      frames, reinterpret_cast<const long long*>(rows), grad_out, out, partial_w, partial_b, batch,
      blocked);
  DERL_LAUNCH_CHECK("stem_bwd_kernel");
  return launch_stem_bwd_reduce(partial_w, partial_b, (int)grid, grad_weight, grad_bias, st);
}

extern "C" int derl_b200_stem_backward_masked(const uint8_t* frames, const int64_t* rows,
                                              int64_t batch, const float* grad_out,
                                              const uint32_t* relu_mask, int blocked,
                                              float* grad_weight, float* grad_bias,
                                              void* workspace, size_t workspace_bytes,
                                              void* stream) {
  DERL_REQUIRE(frames && grad_out && relu_mask && grad_weight && grad_bias && workspace &&
                   batch >= 1, "stem_backward_masked: bad arguments");
  DERL_REQUIRE(blocked == 0 || blocked == 1, "stem_backward_masked: blocked must be 0 or 1");
  DERL_REQUIRE((((uintptr_t)frames | (uintptr_t)grad_out | (uintptr_t)relu_mask) & 15) == 0,
               "stem_backward_masked: inputs must be 16-byte aligned");
  int rc = require_device();
  if (rc != DERL_OK) return rc;
  if (workspace_bytes < derl_b200_stem_backward_workspace_bytes()) {
    set_error("stem_backward_masked: workspace %zu B too small", workspace_bytes);
    return DERL_E_WORKSPACE;
  }
  DERL_REQUIRE(stem_tc_available(), "stem_backward_masked: cuTensorMapEncodeTiled is unavailable");
  return launch_stem_bwd_tc(frames, reinterpret_cast<const long long*>(rows), batch, grad_out,
                            relu_mask, blocked, grad_weight, grad_bias, workspace,
                            as_stream(stream));
}
