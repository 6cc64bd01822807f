// tcgen05 / TMEM / tensor-map PTX wrappers for the sm_100a stem kernels (K6t, K7t).
// Bit layouts follow the PTX ISA "tcgen05" matrix / instruction descriptors (the same fields
// CUTLASS's cute/arch/mma_sm100_desc.hpp spells out); nothing here is library code.
#pragma once

#include <cuda.h>

#include "common.cuh"

namespace derl {

// ---------------------------------------------------------------- host: tensor-map encoder
typedef CUresult (*TmapEncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                 const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                 const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                 CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
TmapEncodeFn tmap_encode_fn();   // cuTensorMapEncodeTiled through the runtime's driver entry point

#ifdef __CUDACC__
// ---------------------------------------------------------------- shared-memory matrix descriptor
// No-swizzle ("interleaved") canonical layout: a core matrix is 8 rows x 16 bytes stored as 128
// contiguous bytes (row r at +16 r).  For a K-major operand `lbo` is the byte distance between
// the two core matrices that are adjacent along K inside one MMA (K = 32 bytes), `sbo` the
// distance between 8-row groups along M/N; for an MN-major operand the core matrix is 8 k-rows
// of 16 contiguous M/N bytes, `lbo` steps along K (next 8 k) and `sbo` along M/N (next 16).
// Bits: [0,14) start >> 4, [16,30) lbo >> 4, [32,46) sbo >> 4, [46,48) version = 1,
// [49,52) base offset = 0, [61,64) layout type = 0 (no swizzle).
__device__ __forceinline__ uint64_t umma_desc(uint32_t smem_addr, uint32_t lbo, uint32_t sbo) {
  return (uint64_t)((smem_addr & 0x3ffffu) >> 4) | ((uint64_t)((lbo >> 4) & 0x3fffu) << 16) |
         ((uint64_t)((sbo >> 4) & 0x3fffu) << 32) | (1ull << 46);
}

// Instruction descriptor of tcgen05.mma.kind::i8, dense, int32 accumulate:
// [4,6) D format = 2 (s32), [7,10) A format (0 = u8, 1 = s8), [10,13) B format,
// [15] A major (0 = K, 1 = MN), [16] B major, [17,23) N >> 3, [24,29) M >> 4.
__host__ __device__ constexpr uint32_t umma_idesc_i8(int m, int n, int a_signed, int b_signed,
                                                     int a_mn_major, int b_mn_major) {
  return (2u << 4) | ((uint32_t)a_signed << 7) | ((uint32_t)b_signed << 10) |
         ((uint32_t)a_mn_major << 15) | ((uint32_t)b_mn_major << 16) |
         ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}

// D[tmem] (+)= A[smem] * B[smem]; issued by ONE thread on behalf of the CTA.
__device__ __forceinline__ void umma_i8(uint32_t tmem_d, uint64_t a_desc, uint64_t b_desc,
                                        uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n\t"
      "}\n" ::"r"(tmem_d),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate), "r"(0u)
      : "memory");
}

// (with a compile-time `accumulate` the setp folds away when the call site is unrolled)

// mbarrier arrival when every tcgen05.mma issued so far by this thread has completed
// (implies tcgen05.fence::before_thread_sync).
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                   smem_u32(bar))
               : "memory");
}

// One lane of a fully converged warp (the pattern the uniform datapath wants: control flow and
// descriptor arithmetic stay warp-uniform, only the issue itself is predicated).
__device__ __forceinline__ bool elect_one_sync() {
  uint32_t pred;
  asm volatile(
      "{\n\t"
      ".reg .pred P;\n\t"
      "elect.sync _|P, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, P;\n\t"
      "}\n"
      : "=r"(pred));
  return pred != 0;
}

__device__ __forceinline__ void tcgen05_fence_before() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tcgen05_fence_after() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}

// ---------------------------------------------------------------- tensor memory
// Whole-warp operations.  `slot` (shared memory) receives the base address: lane in bits
// [16,32), column in bits [0,16).
__device__ __forceinline__ void tmem_alloc(uint32_t* slot, uint32_t columns) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                   smem_u32(slot)),
               "r"(columns)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t base, uint32_t columns) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(base), "r"(columns)
               : "memory");
}

// 32 lanes x 16 consecutive columns of 32-bit words: thread t of the warp receives lane
// (taddr.lane + t), columns taddr.column .. + 15.  A warp may only touch the 32-lane quarter
// (warp id % 4) of tensor memory.
__device__ __forceinline__ void tmem_ld_32x16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]),
        "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]),
        "=r"(v[14]), "=r"(v[15])
      : "r"(taddr)
      : "memory");
}
// 16 lanes x 64 columns (x8 repeats of 256 bits) spread over all 32 threads: thread t holds, for
// rep = 0..7, v[4 rep + {0,1}] = (lane t / 4, columns 8 rep + 2 (t % 4) + {0,1}) and
// v[4 rep + {2,3}] = the same columns of lane t / 4 + 8  (PTX "16x256b" fragment).
__device__ __forceinline__ void tmem_ld_16x256x8(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.16x256b.x8.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]),
        "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]),
        "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]),
        "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
        "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() {
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// ---------------------------------------------------------------- TMA tensor copies (rank 4 load)
__device__ __forceinline__ void tma_load_4d(void* smem_dst, const void* tmap, int c0, int c1, int c2,
                                            int c3, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes "
      "[%0], [%1, {%2, %3, %4, %5}], [%6];" ::"r"(smem_u32(smem_dst)),
      "l"(tmap), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(smem_u32(bar))
      : "memory");
}

__device__ __forceinline__ void named_bar_sync(int id, int threads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory");
}
#endif  // __CUDACC__

}  // namespace derl
