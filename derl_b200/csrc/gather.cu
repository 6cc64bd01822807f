// K2 — permutation-indexed minibatch gather:  dst[j, :] = src[perm[start + j], :].
//
// Replaces the reference's two NumPy fancy-index copies per minibatch
// (derl/runners/onpolicy.py:44-49 and :57-62).  The rollout stays resident and unshuffled
// in HBM; only the composed int64 permutation changes between epochs.
//
//   gather_rows_tma_kernel  wide rows (84*84*4 = 28224 B frame stacks): persistent CTAs, one
//                           elected thread per CTA drives an S-stage ring of TMA bulk copies
//                           global -> shared -> global (cp.async.bulk + mbarrier); no data
//                           ever passes through registers.
//   gather_rows_vec_kernel  any row width: widest aligned LDG/STG that divides the row.
//   gather_columns_kernel   up to 16 narrow columns in one launch, with the float64
//                           {sum, sumsq, count} of one float32 column (advantages) fused in.
#include "common.cuh"

namespace derl {
namespace {

// ------------------------------------------------------------------ TMA bulk row gather
constexpr int kStageBytes = 28672;  // >= one 28224-B frame stack; multiple of 128
constexpr int kStages = 8;          // 8 * 28 KiB = 224 KiB of the 227 KiB a CTA may own
constexpr size_t kGatherSmem = (size_t)kStages * kStageBytes + 16 * kStages;

// `mirror` (optional): a second destination indexed like the SOURCE — every gathered row is
// also written to mirror[src_row].  With `src` in pinned host memory this is the first-epoch
// upload path: the minibatch and the resident device copy of the rollout are both filled by
// the one pass over PCIe (a permutation's minibatches partition the rollout).
__global__ void __launch_bounds__(32, 1)
gather_rows_tma_kernel(const uint8_t* __restrict__ src, long long row_bytes,
                       const long long* __restrict__ perm, long long start, long long count,
                       uint8_t* __restrict__ dst, uint8_t* __restrict__ mirror, int chunk_bytes,
                       int chunks_per_row) {
  extern __shared__ __align__(128) uint8_t smem[];
  if (threadIdx.x != 0) return;
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + (size_t)kStages * kStageBytes);
  long long* stage_row = reinterpret_cast<long long*>(full + kStages);
#pragma unroll
  for (int s = 0; s < kStages; ++s) mbar_init(&full[s], 1);
  mbar_fence_init();

  // Work unit u = (row j, chunk q); units are dealt round-robin so that concurrently
  // running CTAs write neighbouring destination rows.
  const long long total = count * chunks_per_row;
  const long long first = blockIdx.x, step = gridDim.x;
  const long long mine = first < total ? (total - first + step - 1) / step : 0;

  auto unit_bytes = [&](long long q) -> uint32_t {
    const long long left = row_bytes - q * chunk_bytes;
    return (uint32_t)(left < chunk_bytes ? left : chunk_bytes);
  };
  auto load = [&](long long k, long long src_row) {
    const long long u = first + k * step;
    const long long q = u % chunks_per_row;
    const int s = (int)(k % kStages);
    const uint32_t bytes = unit_bytes(q);
    stage_row[s] = src_row;
    mbar_expect_tx(&full[s], bytes);
    bulk_g2s(smem + (size_t)s * kStageBytes, src + src_row * row_bytes + q * chunk_bytes, bytes,
             &full[s]);
  };
  auto src_row_of = [&](long long k) -> long long {
    const long long u = first + k * step;
    return __ldg(perm + start + u / chunks_per_row);
  };

  long long issued = 0;
  for (; issued < kStages && issued < mine; ++issued) load(issued, src_row_of(issued));
  long long next_row = issued < mine ? src_row_of(issued) : 0;  // index prefetched one unit ahead

  for (long long i = 0; i < mine; ++i) {
    const int s = (int)(i % kStages);
    mbar_wait(&full[s], (uint32_t)((i / kStages) & 1));
    const long long u = first + i * step;
    const long long j = u / chunks_per_row, q = u % chunks_per_row;
    bulk_s2g(dst + j * row_bytes + q * chunk_bytes, smem + (size_t)s * kStageBytes, unit_bytes(q));
    if (mirror != nullptr) {
      bulk_s2g(mirror + stage_row[s] * row_bytes + q * chunk_bytes, smem + (size_t)s * kStageBytes,
               unit_bytes(q));
    }
    bulk_commit();
    if (i >= 1 && issued < mine) {
      bulk_wait_read<1>();  // store of unit i-1 has drained its stage -> refill it
      load(issued, next_row);
      ++issued;
      if (issued < mine) next_row = src_row_of(issued);
    }
  }
  bulk_wait<0>();
}

// ------------------------------------------------------------------ generic vector row gather
template <typename V>
__global__ void __launch_bounds__(256)
gather_rows_vec_kernel(const V* __restrict__ src, long long row_elems,
                       const long long* __restrict__ perm, long long start, long long count,
                       V* __restrict__ dst) {
  const long long total = count * row_elems;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += stride) {
    const long long j = idx / row_elems;
    const long long e = idx - j * row_elems;
    dst[idx] = __ldg(src + __ldg(perm + start + j) * row_elems + e);
  }
}

template <typename V>
int launch_vec(const void* src, long long row_bytes, const long long* perm, long long start,
               long long count, void* dst, cudaStream_t st) {
  const long long row_elems = row_bytes / (long long)sizeof(V);
  const long long total = count * row_elems;
  long long blocks = (total + 255) / 256;
  const long long cap = (long long)sm_count() * 32;
  if (blocks > cap) blocks = cap;
  gather_rows_vec_kernel<V><<<(unsigned)blocks, 256, 0, st>>>(
      reinterpret_cast<const V*>(src), row_elems, perm, start, count, reinterpret_cast<V*>(dst));
  DERL_LAUNCH_CHECK("gather_rows_vec_kernel");
  return DERL_OK;
}

inline int pow2_align(uintptr_t a, uintptr_t b, long long bytes) {
  const uintptr_t m = a | b | (uintptr_t)bytes;
  if ((m & 15) == 0) return 16;
  if ((m & 7) == 0) return 8;
  if ((m & 3) == 0) return 4;
  if ((m & 1) == 0) return 2;
  return 1;
}

// ------------------------------------------------------------------ narrow columns + moments
struct ColumnArgs {
  const uint8_t* src[DERL_MAX_COLUMNS];
  uint8_t* dst[DERL_MAX_COLUMNS];
  int row_bytes[DERL_MAX_COLUMNS];
  int unit[DERL_MAX_COLUMNS];  // copy width in bytes (1, 2, 4, 8 or 16)
  int n;
  int moments_col;
};

template <typename V>
__device__ __forceinline__ void copy_units(const uint8_t* s, uint8_t* d, int bytes) {
  const V* sv = reinterpret_cast<const V*>(s);
  V* dv = reinterpret_cast<V*>(d);
  const int n = bytes / (int)sizeof(V);
  for (int i = 0; i < n; ++i) dv[i] = __ldg(sv + i);
}

__global__ void __launch_bounds__(256)
gather_columns_kernel(const __grid_constant__ ColumnArgs a, const long long* __restrict__ perm,
                      long long start, long long count, void* workspace, double* stats) {
  __shared__ double scratch[2 * 32];
  __shared__ int flag;
  double s1 = 0.0, s2 = 0.0;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x; j < count; j += stride) {
    const long long p = __ldg(perm + start + j);
    for (int c = 0; c < a.n; ++c) {
      const int rb = a.row_bytes[c];
      const uint8_t* s = a.src[c] + p * rb;
      uint8_t* d = a.dst[c] + j * rb;
      switch (a.unit[c]) {
        case 16: copy_units<uint4>(s, d, rb); break;
        case 8: copy_units<unsigned long long>(s, d, rb); break;
        case 4: copy_units<unsigned>(s, d, rb); break;
        case 2: copy_units<unsigned short>(s, d, rb); break;
        default: copy_units<uint8_t>(s, d, rb); break;
      }
      if (c == a.moments_col) {
        const double v = (double)__ldg(reinterpret_cast<const float*>(s));
        s1 += v;
        s2 += v * v;
      }
    }
  }
  if (a.moments_col >= 0) {
    double v[2] = {s1, s2};
    block_sum<2>(v, scratch);
    if (publish_partials<2>(v, workspace, &flag)) {
      double tot[2];
      final_sum<2>(tot, workspace, scratch);
      if (threadIdx.x == 0) {
        stats[0] = tot[0];
        stats[1] = tot[1];
        stats[2] = (double)count;
      }
    }
  }
}

}  // namespace
}  // namespace derl

using namespace derl;

extern "C" {

static int gather_rows_impl(const void* src, int64_t n_src_rows, int64_t row_bytes,
                            const int64_t* perm, int64_t start, int64_t count, void* dst,
                            void* mirror, int max_ctas, void* stream) {
  DERL_REQUIRE(src && perm && dst, "gather_rows: null pointer");
  DERL_REQUIRE(row_bytes >= 1 && n_src_rows >= 1 && start >= 0 && count >= 0,
               "gather_rows: bad sizes (row_bytes=%lld rows=%lld start=%lld count=%lld)",
               (long long)row_bytes, (long long)n_src_rows, (long long)start, (long long)count);
  int rc = require_device();
  if (rc != DERL_OK) return rc;
  if (count == 0) return DERL_OK;
  cudaStream_t st = as_stream(stream);
  const long long* p = reinterpret_cast<const long long*>(perm);
  const int align = pow2_align((uintptr_t)src | (uintptr_t)mirror, (uintptr_t)dst, row_bytes);
  DERL_REQUIRE(mirror == nullptr || (align == 16 && row_bytes >= 2048),
               "gather_rows_upload: rows must be 16-byte multiples of at least 2048 bytes");
  if (align == 16 && row_bytes >= 2048) {
    const int cpr = (int)((row_bytes + kStageBytes - 1) / kStageBytes);
    // equal 16-byte-multiple chunks; the last one takes the remainder
    long long chunk = ((row_bytes + cpr - 1) / cpr + 15) & ~15ll;
    if (int rc_attr = ensure_dynamic_smem(reinterpret_cast<const void*>(gather_rows_tma_kernel), (int)kGatherSmem)) return rc_attr;
    const long long units = count * cpr;
    long long grid = sm_count();
    if (max_ctas > 0 && grid > max_ctas) grid = max_ctas;
    if (grid > units) grid = units;
    gather_rows_tma_kernel<<<(unsigned)grid, 32, kGatherSmem, st>>>(
        reinterpret_cast<const uint8_t*>(src), row_bytes, p, start, count,
        reinterpret_cast<uint8_t*>(dst), reinterpret_cast<uint8_t*>(mirror), (int)chunk, cpr);
    DERL_LAUNCH_CHECK("gather_rows_tma_kernel");
    return DERL_OK;
  }
  switch (align) {
    case 16: return launch_vec<uint4>(src, row_bytes, p, start, count, dst, st);
    case 8: return launch_vec<unsigned long long>(src, row_bytes, p, start, count, dst, st);
    case 4: return launch_vec<unsigned>(src, row_bytes, p, start, count, dst, st);
    case 2: return launch_vec<unsigned short>(src, row_bytes, p, start, count, dst, st);
    default: return launch_vec<uint8_t>(src, row_bytes, p, start, count, dst, st);
  }
}

int derl_b200_gather_rows(const void* src, int64_t n_src_rows, int64_t row_bytes,
                          const int64_t* perm, int64_t start, int64_t count, void* dst,
                          void* stream) {
  return gather_rows_impl(src, n_src_rows, row_bytes, perm, start, count, dst, nullptr, 0, stream);
}

int derl_b200_gather_rows_upload(const void* src_host, int64_t n_src_rows, int64_t row_bytes,
                                 const int64_t* perm, int64_t start, int64_t count, void* dst,
                                 void* resident, int max_ctas, void* stream) {
  DERL_REQUIRE(resident != nullptr, "gather_rows_upload: resident buffer is NULL");
  return gather_rows_impl(src_host, n_src_rows, row_bytes, perm, start, count, dst, resident,
                          max_ctas, stream);
}

int derl_b200_gather_columns(int n_columns, const void* const* src, const int64_t* row_bytes,
                             void* const* dst, const int64_t* perm, int64_t start,
                             int64_t count, int moments_col, double* stats, void* workspace,
                             size_t workspace_bytes, void* stream) {
  DERL_REQUIRE(n_columns >= 1 && n_columns <= DERL_MAX_COLUMNS,
               "gather_columns: n_columns=%d outside [1, %d]", n_columns, DERL_MAX_COLUMNS);
  DERL_REQUIRE(src && row_bytes && dst && perm && start >= 0 && count >= 0,
               "gather_columns: bad arguments");
  DERL_REQUIRE(moments_col < n_columns, "gather_columns: moments_col=%d out of range",
               moments_col);
  ColumnArgs a;
  a.n = n_columns;
  a.moments_col = moments_col < 0 ? -1 : moments_col;
  for (int c = 0; c < n_columns; ++c) {
    DERL_REQUIRE(src[c] && dst[c] && row_bytes[c] >= 1 && row_bytes[c] <= (1 << 20),
                 "gather_columns: column %d has a null pointer or row_bytes outside [1, 2^20]", c);
    a.src[c] = reinterpret_cast<const uint8_t*>(src[c]);
    a.dst[c] = reinterpret_cast<uint8_t*>(dst[c]);
    a.row_bytes[c] = (int)row_bytes[c];
    a.unit[c] = pow2_align((uintptr_t)src[c], (uintptr_t)dst[c], row_bytes[c]);
  }
  if (a.moments_col >= 0) {
    DERL_REQUIRE(a.row_bytes[a.moments_col] == 4 && a.unit[a.moments_col] >= 4,
                 "gather_columns: moments column must be float32 with 4-byte rows");
    DERL_REQUIRE(stats && workspace, "gather_columns: moments need stats and workspace");
    if (workspace_bytes < derl_b200_moments_workspace_bytes(count)) {
      set_error("gather_columns: workspace %zu B too small", workspace_bytes);
      return DERL_E_WORKSPACE;
    }
  }
  int rc = require_device();
  if (rc != DERL_OK) return rc;
  if (count == 0) return DERL_OK;
  cudaStream_t st = as_stream(stream);
  if (a.moments_col >= 0) DERL_CUDA(cudaMemsetAsync(workspace, 0, kTicketBytes, st));
  long long blocks = (count + 255) / 256;
  long long cap = (long long)sm_count() * 8;
  if (cap > kMaxReduceBlocks) cap = kMaxReduceBlocks;
  if (blocks > cap) blocks = cap;
  gather_columns_kernel<<<(unsigned)blocks, 256, 0, st>>>(
      a, reinterpret_cast<const long long*>(perm), start, count, workspace, stats);
  DERL_LAUNCH_CHECK("gather_columns_kernel");
  return DERL_OK;
}

}  // extern "C"
