// K1 — GAE / returns: reverse scan over T, one lane per environment.
//
// Arithmetic contract (bit-exact with the reference's NumPy evaluation,
// derl/runners/trajectory_transforms.py:45-63; scalar-order statement in SURVEY.md §8a):
//   c_t = reset_t ? 0.0 : gamma            (float64; (1 - reset) * gamma, :57,:59)
//   k_t = reset_t ? 0.0 : gamma * lambda   (float64; ... * lambda, :62)
//   A[T-1] = F( F(r - v) + c * last_value )                       (:46 then :53)
//   A[t]   = F( ((r[t] + c_t * v[t+1]) - v[t]) + k_t * A[t+1] )   (:58-62)
//   VT[t]  = A[t] +_f32 v[t]                                      (:63)
// F = round-to-nearest float32; every other operation is an IEEE float64 add/mul issued
// through __dadd_rn/__dmul_rn so ptxas can never contract them into DFMA.
//
// Two variants compute identical bits:
//   DIRECT  any shape; register double-buffered coalesced global loads.
//   TMA     N % 16 == 0; each warp owns a 32- or 64-env strip, [TT x W] tiles of rewards /
//           values / resets are staged through shared memory by cp.async.bulk.tensor
//           with an mbarrier ring, outputs leave through TMA stores.
#include <cuda.h>  // CUtensorMap types only; the encoder is fetched through the runtime

#include <stdlib.h>

#include <type_traits>

#include "common.cuh"

namespace derl {
namespace {

// ------------------------------------------------------------------ per-step arithmetic
template <typename RT>
__device__ __forceinline__ float gae_last_row(RT r, float v, unsigned reset, float last_value,
                                              double gamma) {
  float base;
  if constexpr (std::is_same<RT, double>::value) {
    base = __double2float_rn(__dsub_rn(r, (double)v));
  } else {
    base = __fsub_rn(r, v);
  }
  const double c = reset ? 0.0 : gamma;
  return __double2float_rn(__dadd_rn((double)base, __dmul_rn(c, (double)last_value)));
}

template <typename RT>
__device__ __forceinline__ float gae_row(RT r, float v, float v_next, float a_next,
                                         unsigned reset, double gamma, double gamma_lambda) {
  const double c = reset ? 0.0 : gamma;
  const double k = reset ? 0.0 : gamma_lambda;
  const double delta = __dsub_rn(__dadd_rn((double)r, __dmul_rn(c, (double)v_next)), (double)v);
  return __double2float_rn(__dadd_rn(delta, __dmul_rn(k, (double)a_next)));
}

// Block partial -> workspace -> last block writes {sum, sumsq, count}.
__device__ __forceinline__ void finish_stats(double s1, double s2, double count,
                                             void* workspace, double* stats, double* scratch,
                                             int* flag) {
  double v[2] = {s1, s2};
  block_sum<2>(v, scratch);
  if (publish_partials<2>(v, workspace, flag)) {
    double tot[2];
    final_sum<2>(tot, workspace, scratch);
    if (threadIdx.x == 0) {
      stats[0] = tot[0];
      stats[1] = tot[1];
      stats[2] = count;
    }
  }
}

// ------------------------------------------------------------------ DIRECT variant
template <typename RT, int U>
__global__ void __launch_bounds__(128)
gae_direct_kernel(const RT* __restrict__ rewards, const float* __restrict__ values,
                  const uint8_t* __restrict__ resets, const float* __restrict__ last_value,
                  long long T, long long N, double gamma, double gamma_lambda,
                  float* __restrict__ adv, float* __restrict__ vt, void* workspace,
                  double* stats) {
  __shared__ double scratch[2 * 32];
  __shared__ int flag;
  const long long n = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  double s1 = 0.0, s2 = 0.0;
  if (n < N) {
    const long long nchunks = (T + U - 1) / U;
    RT r_cur[U], r_nxt[U];
    float v_cur[U], v_nxt[U];
    uint8_t z_cur[U], z_nxt[U];
    {  // top chunk, possibly partial
      const long long t0 = (nchunks - 1) * U;
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const long long t = t0 + u;
        const bool ok = t < T;
        const long long i = ok ? t * N + n : n;
        r_cur[u] = __ldg(rewards + i);
        v_cur[u] = __ldg(values + i);
        z_cur[u] = __ldg(resets + i);
      }
    }
    float v_next = __ldg(last_value + n);
    float a_next = 0.f;
    for (long long c = nchunks - 1; c >= 0; --c) {
      const long long t0 = c * U;
      if (c > 0) {  // prefetch the chunk below while this one is scanned
        const long long p0 = t0 - U;
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const long long i = (p0 + u) * N + n;
          r_nxt[u] = __ldg(rewards + i);
          v_nxt[u] = __ldg(values + i);
          z_nxt[u] = __ldg(resets + i);
        }
      }
#pragma unroll
      for (int u = U - 1; u >= 0; --u) {
        const long long t = t0 + u;
        if (t < T) {
          const float v = v_cur[u];
          const float a = (t == T - 1)
                              ? gae_last_row<RT>(r_cur[u], v, z_cur[u], v_next, gamma)
                              : gae_row<RT>(r_cur[u], v, v_next, a_next, z_cur[u], gamma,
                                            gamma_lambda);
          adv[t * N + n] = a;
          vt[t * N + n] = __fadd_rn(a, v);
          s1 += (double)a;
          s2 += (double)a * (double)a;
          a_next = a;
          v_next = v;
        }
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        r_cur[u] = r_nxt[u];
        v_cur[u] = v_nxt[u];
        z_cur[u] = z_nxt[u];
      }
    }
  }
  if (stats != nullptr) {
    finish_stats(s1, s2, (double)T * (double)N, workspace, stats, scratch, &flag);
  }
}

// ------------------------------------------------------------------ TMA variant
// One warp per CTA owns a strip of W = 32*E envs (E independent chains per lane for ILP).
// Time is walked top-down in chunks of TT steps; each chunk's [TT x W] tiles of rewards /
// values / resets arrive by 2-D TMA into an S-stage mbarrier ring, results leave through a
// double-buffered pair of output tiles and TMA stores.  Interior chunks run a branch-free
// body: everything that does not depend on A[t+1] (conversions, delta_t, k_t) is computed
// for all TT steps first, so the serial chain per step is DMUL -> DADD -> F2F -> F2F only.
constexpr int kLanes = 32;

template <typename RT, int TT, int S, int E, int NW>
struct GaeTmaSmem {
  static constexpr int kW = kLanes * E * NW;
  static constexpr int kTile = TT * kW;
  static constexpr size_t r_off = 0;
  static constexpr size_t v_off = r_off + sizeof(RT) * S * kTile;
  static constexpr size_t a_off = v_off + sizeof(float) * S * kTile;
  static constexpr size_t vt_off = a_off + sizeof(float) * 2 * kTile;
  static constexpr size_t z_off = vt_off + sizeof(float) * 2 * kTile;
  static constexpr size_t bar_off = z_off + (size_t)S * kTile;
  static constexpr size_t bytes = bar_off + 8 * S;
  static constexpr uint32_t stage_tx = kTile * (sizeof(RT) + sizeof(float) + 1);
};

template <int NW>
__device__ __forceinline__ void strip_sync() {
  if constexpr (NW == 1) {
    __syncwarp();
  } else {
    __syncthreads();
  }
}

template <typename RT, int TT, int S, int E, int NW, bool kStats>
__global__ void __launch_bounds__(kLanes * NW)
gae_tma_kernel(const __grid_constant__ CUtensorMap tm_r, const __grid_constant__ CUtensorMap tm_v,
               const __grid_constant__ CUtensorMap tm_z, const __grid_constant__ CUtensorMap tm_a,
               const __grid_constant__ CUtensorMap tm_vt, const float* __restrict__ last_value,
               int T, int N, double gamma, double gamma_lambda, void* workspace, double* stats) {
  using L = GaeTmaSmem<RT, TT, S, E, NW>;
  constexpr int W = L::kW;
  extern __shared__ __align__(128) uint8_t smem[];
  RT* sr = reinterpret_cast<RT*>(smem + L::r_off);
  float* sv = reinterpret_cast<float*>(smem + L::v_off);
  float* sa = reinterpret_cast<float*>(smem + L::a_off);
  float* svt = reinterpret_cast<float*>(smem + L::vt_off);
  uint8_t* sz = smem + L::z_off;
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + L::bar_off);
  __shared__ double scratch[2 * 32];
  __shared__ int flag;

  const int lane = threadIdx.x & 31;
  const int col0 = (threadIdx.x >> 5) * (kLanes * E);  // this warp's first column in the tile
  const bool leader = threadIdx.x == 0;
  const int n0 = blockIdx.x * W;
  const int nchunks = (T + TT - 1) / TT;

  if (leader) {
    tma_prefetch_desc(&tm_r);
    tma_prefetch_desc(&tm_v);
    tma_prefetch_desc(&tm_z);
    tma_prefetch_desc(&tm_a);
    tma_prefetch_desc(&tm_vt);
#pragma unroll
    for (int s = 0; s < S; ++s) mbar_init(&full[s], 1);
    mbar_fence_init();
  }
  strip_sync<NW>();

  auto issue = [&](int chunk, int s) {
    mbar_expect_tx(&full[s], L::stage_tx);
    tma_load_2d(sr + (size_t)s * L::kTile, &tm_r, n0, chunk * TT, &full[s]);
    tma_load_2d(sv + (size_t)s * L::kTile, &tm_v, n0, chunk * TT, &full[s]);
    tma_load_2d(sz + (size_t)s * L::kTile, &tm_z, n0, chunk * TT, &full[s]);
  };
  if (leader) {
    for (int i = 0; i < S && i < nchunks; ++i) issue(nchunks - 1 - i, i);
  }

  float v_next[E], a_next[E];
  bool live[E];
#pragma unroll
  for (int e = 0; e < E; ++e) {
    const int n = n0 + col0 + e * kLanes + lane;
    live[e] = n < N;
    v_next[e] = live[e] ? __ldg(last_value + n) : 0.f;
    a_next[e] = 0.f;
  }
  double s1 = 0.0, s2 = 0.0;

  for (int i = 0; i < nchunks; ++i) {
    const int c = nchunks - 1 - i;
    const int s = i % S;
    const int o = i & 1;
    mbar_wait(&full[s], (uint32_t)((i / S) & 1));
    const RT* r_t = sr + (size_t)s * L::kTile + col0 + lane;
    const float* v_t = sv + (size_t)s * L::kTile + col0 + lane;
    const uint8_t* z_t = sz + (size_t)s * L::kTile + col0 + lane;
    float* a_t = sa + (size_t)o * L::kTile + col0 + lane;
    float* vt_t = svt + (size_t)o * L::kTile + col0 + lane;
    const int t0 = c * TT;

    if (i == 0) {
      // top chunk: may be partial (rows >= T are TMA zero fill) and holds the last row
#pragma unroll 1
      for (int tt = TT - 1; tt >= 0; --tt) {
        const int t = t0 + tt;
        if (t >= T) continue;
#pragma unroll
        for (int e = 0; e < E; ++e) {
          const int at = tt * W + e * kLanes;
          const RT r = r_t[at];
          const float v = v_t[at];
          const unsigned z = z_t[at];
          const float a = (t == T - 1)
                              ? gae_last_row<RT>(r, v, z, v_next[e], gamma)
                              : gae_row<RT>(r, v, v_next[e], a_next[e], z, gamma, gamma_lambda);
          a_t[at] = a;
          vt_t[at] = __fadd_rn(a, v);
          if (kStats && live[e]) {
            s1 += (double)a;
            s2 += (double)a * (double)a;
          }
          a_next[e] = a;
          v_next[e] = v;
        }
      }
    } else {
      float v[TT][E];
      double delta[TT][E], k[TT][E];
#pragma unroll
      for (int tt = 0; tt < TT; ++tt) {
#pragma unroll
        for (int e = 0; e < E; ++e) v[tt][e] = v_t[tt * W + e * kLanes];
      }
#pragma unroll
      for (int tt = 0; tt < TT; ++tt) {
#pragma unroll
        for (int e = 0; e < E; ++e) {
          const int at = tt * W + e * kLanes;
          const unsigned z = z_t[at];
          const double r = (double)r_t[at];
          const float vn = tt == TT - 1 ? v_next[e] : v[tt + 1][e];
          const double cz = z ? 0.0 : gamma;
          k[tt][e] = z ? 0.0 : gamma_lambda;
          delta[tt][e] = __dsub_rn(__dadd_rn(r, __dmul_rn(cz, (double)vn)), (double)v[tt][e]);
        }
      }
#pragma unroll
      for (int tt = TT - 1; tt >= 0; --tt) {
#pragma unroll
        for (int e = 0; e < E; ++e) {
          const int at = tt * W + e * kLanes;
          const float a =
              __double2float_rn(__dadd_rn(delta[tt][e], __dmul_rn(k[tt][e], (double)a_next[e])));
          a_next[e] = a;
          a_t[at] = a;
          vt_t[at] = __fadd_rn(a, v[tt][e]);
          if (kStats && live[e]) {
            s1 += (double)a;
            s2 += (double)a * (double)a;
          }
        }
      }
#pragma unroll
      for (int e = 0; e < E; ++e) v_next[e] = v[0][e];
    }
    fence_proxy_async_smem();  // own writes to the out tile -> visible to the async proxy
    strip_sync<NW>();          // every lane is done reading stage s and writing out-buffer o
    if (leader) {
      tma_store_2d(&tm_a, sa + (size_t)o * L::kTile, n0, t0);
      tma_store_2d(&tm_vt, svt + (size_t)o * L::kTile, n0, t0);
      bulk_commit();
      if (i + S < nchunks) issue(c - S, s);
      bulk_wait_read<1>();  // the store before this one has left smem: buffer o^1 is free
    }
    strip_sync<NW>();
  }
  if (leader) bulk_wait<0>();
  if (kStats) finish_stats(s1, s2, (double)T * (double)N, workspace, stats, scratch, &flag);
}

// ------------------------------------------------------------------ tensor-map encoding
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                  const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_tiled_fn() {
  static EncodeTiledFn fn = nullptr;
  if (fn == nullptr) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) ==
            cudaSuccess &&
        q == cudaDriverEntryPointSuccess) {
      fn = reinterpret_cast<EncodeTiledFn>(p);
    } else {
      cudaGetLastError();
    }
  }
  return fn;
}

// [T, N] row-major array viewed as a 2-D tensor {N (inner), T}; box {W, TT}.
// A CUtensorMap is a pure function of (address, shape, box, type): the encodings are memoised per
// thread (a PPO loop hands the same five arrays back every rollout), which takes ~45 us of driver
// calls per GAE launch off the host path — at 4096 envs x 128 steps that was 4x the kernel time.
struct StripMapKey {
  const void* base;
  long long T, N;
  int TT, W, dt;
  bool operator==(const StripMapKey& o) const {
    return base == o.base && T == o.T && N == o.N && TT == o.TT && W == o.W && dt == o.dt;
  }
};
constexpr int kStripMapSlots = 32;
struct StripMapCache {
  StripMapKey keys[kStripMapSlots];
  CUtensorMap maps[kStripMapSlots];
  int used = 0, next = 0;
};

bool make_strip_map(CUtensorMap* map, CUtensorMapDataType dt, size_t elem, const void* base,
                    long long T, long long N, int TT, int W) {
  static thread_local StripMapCache cache;
  const StripMapKey key{base, T, N, TT, W, (int)dt};
  for (int i = 0; i < cache.used; ++i) {
    if (cache.keys[i] == key) {
      *map = cache.maps[i];
      return true;
    }
  }
  EncodeTiledFn enc = encode_tiled_fn();
  if (enc == nullptr) return false;
  cuuint64_t dims[2] = {(cuuint64_t)N, (cuuint64_t)T};
  cuuint64_t strides[1] = {(cuuint64_t)N * elem};
  cuuint32_t box[2] = {(cuuint32_t)W, (cuuint32_t)TT};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(map, dt, 2, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return false;
  const int slot = cache.used < kStripMapSlots ? cache.used++ : cache.next;
  cache.next = (slot + 1) % kStripMapSlots;
  cache.keys[slot] = key;
  cache.maps[slot] = *map;
  return true;
}

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

template <typename RT, int TT, int S, int E, int NW>
int launch_tma(const void* rewards, const float* values, const uint8_t* resets,
               const float* last_value, long long T, long long N, double gamma, double gl,
               float* adv, float* vt, double* stats, void* workspace, cudaStream_t st) {
  using L = GaeTmaSmem<RT, TT, S, E, NW>;
  constexpr int W = L::kW;
  CUtensorMap tm_r, tm_v, tm_z, tm_a, tm_vt;
  const CUtensorMapDataType rdt = std::is_same<RT, double>::value
                                      ? CU_TENSOR_MAP_DATA_TYPE_FLOAT64
                                      : CU_TENSOR_MAP_DATA_TYPE_FLOAT32;
  bool ok = make_strip_map(&tm_r, rdt, sizeof(RT), rewards, T, N, TT, W) &&
            make_strip_map(&tm_v, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, values, T, N, TT, W) &&
            make_strip_map(&tm_z, CU_TENSOR_MAP_DATA_TYPE_UINT8, 1, resets, T, N, TT, W) &&
            make_strip_map(&tm_a, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, adv, T, N, TT, W) &&
            make_strip_map(&tm_vt, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, vt, T, N, TT, W);
  if (!ok) {
    set_error("cuTensorMapEncodeTiled failed for GAE strips (T=%lld N=%lld)", T, N);
    return DERL_E_CUDA;
  }
  const unsigned grid = (unsigned)((N + W - 1) / W);
  auto go = [&](auto kern) -> int {
    if (int rc_attr = ensure_dynamic_smem(reinterpret_cast<const void*>(kern), (int)L::bytes))
      return rc_attr;
    kern<<<grid, kLanes * NW, L::bytes, st>>>(tm_r, tm_v, tm_z, tm_a, tm_vt, last_value, (int)T,
                                              (int)N, gamma, gl, workspace, stats);
    DERL_LAUNCH_CHECK("gae_tma_kernel");
    return DERL_OK;
  };
  return stats != nullptr ? go(gae_tma_kernel<RT, TT, S, E, NW, true>)
                          : go(gae_tma_kernel<RT, TT, S, E, NW, false>);
}

// Tile configuration, from the sweep in profiles/r01_gae_tile_sweep.txt (B200, L2 flushed,
// GB/s of algorithmic bytes at T2048 x N65536 | T128 x N65536 | T128 x N4096):
//   cfg 0  W= 32, 1 warp,  TT16 S4   5487 | 4360 | 624    8 CTAs/SM; best for very few envs
//   cfg 1  W= 64, 1 warp x 2 chains/lane, TT8 S4   5656 | 4360 | 484
//   cfg 2  W=128, 4 warps, TT8  S4   5600 | 4636 | 545    4 CTAs/SM: one wave at N = 65536
//   cfg 5  W=128, 4 warps, TT16 S3   6190 | 4360 | 623    512-B rows per TMA box line
// Wider strips mean longer contiguous DRAM bursts per box row, which is what lifts the
// large shapes from 84 % to 94 % of the measured HBM peak.  DERL_GAE_TMA_CFG overrides.
template <typename RT>
int launch_tma_auto(const void* rewards, const float* values, const uint8_t* resets,
                    const float* last_value, long long T, long long N, double gamma, double gl,
                    float* adv, float* vt, double* stats, void* workspace, cudaStream_t st) {
#define DERL_GAE_ARGS rewards, values, resets, last_value, T, N, gamma, gl, adv, vt, stats, workspace, st
  int cfg = N < 128 ? 0 : (T <= 256 && N > 8192 ? 2 : 5);
  if (const char* env = getenv("DERL_GAE_TMA_CFG")) cfg = atoi(env);
  switch (cfg) {
    case 1: return launch_tma<RT, 8, 4, 2, 1>(DERL_GAE_ARGS);
    case 2: return launch_tma<RT, 8, 4, 1, 4>(DERL_GAE_ARGS);
    case 5: return launch_tma<RT, 16, 3, 1, 4>(DERL_GAE_ARGS);
    default: return launch_tma<RT, 16, 4, 1, 1>(DERL_GAE_ARGS);
  }
#undef DERL_GAE_ARGS
}

template <typename RT>
int launch_direct(const void* rewards, const float* values, const uint8_t* resets,
                  const float* last_value, long long T, long long N, double gamma, double gl,
                  float* adv, float* vt, double* stats, void* workspace, cudaStream_t st) {
  constexpr int U = 8;
  // Few envs: one warp per CTA so that every strip gets an SM to itself.
  const int block = N <= (long long)sm_count() * 32 * 4 ? 32 : 128;
  const unsigned grid = (unsigned)((N + block - 1) / block);
  gae_direct_kernel<RT, U><<<grid, block, 0, st>>>(
      reinterpret_cast<const RT*>(rewards), values, resets, last_value, T, N, gamma, gl, adv, vt,
      workspace, stats);
  DERL_LAUNCH_CHECK("gae_direct_kernel");
  return DERL_OK;
}

// ------------------------------------------------------------------ moments / normalise
__global__ void __launch_bounds__(256)
moments_kernel(const float* __restrict__ x, long long count, void* workspace, double* stats) {
  __shared__ double scratch[2 * 32];
  __shared__ int flag;
  double s1 = 0.0, s2 = 0.0;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += stride) {
    const double v = (double)__ldg(x + i);
    s1 += v;
    s2 += v * v;
  }
  finish_stats(s1, s2, (double)count, workspace, stats, scratch, &flag);
}

// y = (x - mean) / (std + eps): float32 elementwise like NumPy (trajectory_transforms.py:68,
// :91-92); mean/std come from float64 moments rounded once to float32.
__global__ void __launch_bounds__(256)
normalize_kernel(const float* __restrict__ x, float* __restrict__ out, long long count,
                 const double* __restrict__ stats, double epsilon) {
  const double n = stats[2];
  const double mean_d = stats[0] / n;
  double var_d = stats[1] / n - mean_d * mean_d;
  var_d = var_d > 0.0 ? var_d : 0.0;
  const float mean = (float)mean_d;
  const float denom = (float)((double)(float)sqrt(var_d) + epsilon);
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += stride) {
    out[i] = __fdiv_rn(__fsub_rn(x[i], mean), denom);
  }
}

unsigned reduce_grid(long long count, int block) {
  long long want = (count + block - 1) / block;
  long long cap = (long long)sm_count() * 8;
  if (want > cap) want = cap;
  if (want > kMaxReduceBlocks) want = kMaxReduceBlocks;
  return (unsigned)(want < 1 ? 1 : want);
}

}  // namespace
}  // namespace derl

using namespace derl;

extern "C" {

size_t derl_b200_gae_workspace_bytes(int64_t T, int64_t N) {
  (void)T;
  if (N < 1) N = 1;
  const size_t blocks = (size_t)((N + 31) / 32);
  return kTicketBytes + blocks * 2 * sizeof(double);
}

size_t derl_b200_moments_workspace_bytes(int64_t count) {
  (void)count;
  return kTicketBytes + (size_t)kMaxReduceBlocks * 2 * sizeof(double);
}

int derl_b200_gae(const void* rewards, int rewards_f64, const float* values,
                  const uint8_t* resets, const float* last_value, int64_t T, int64_t N,
                  double gamma, double lambda, float* adv, float* vt, double* stats,
                  void* workspace, size_t workspace_bytes, int variant, void* stream) {
  DERL_REQUIRE(T >= 1 && N >= 1, "gae: need T >= 1 and N >= 1 (got T=%lld N=%lld)",
               (long long)T, (long long)N);
  DERL_REQUIRE(rewards && values && resets && last_value && adv && vt, "gae: null pointer");
  DERL_REQUIRE(variant >= DERL_GAE_AUTO && variant <= DERL_GAE_TMA, "gae: bad variant %d",
               variant);
  if (stats != nullptr) {
    DERL_REQUIRE(workspace != nullptr, "gae: stats requested without a workspace");
    if (workspace_bytes < derl_b200_gae_workspace_bytes(T, N)) {
      set_error("gae: workspace %zu B < required %zu B", workspace_bytes,
                derl_b200_gae_workspace_bytes(T, N));
      return DERL_E_WORKSPACE;
    }
  }
  int rc = require_device();
  if (rc != DERL_OK) return rc;
  cudaStream_t st = as_stream(stream);
  if (stats != nullptr) DERL_CUDA(cudaMemsetAsync(workspace, 0, kTicketBytes, st));

  const bool tma_ok = (N % 16 == 0) && N >= 32 && T < (1ll << 31) && N < (1ll << 31) &&
                      aligned16(rewards) && aligned16(values) && aligned16(resets) &&
                      aligned16(adv) && aligned16(vt) && encode_tiled_fn() != nullptr;
  if (variant == DERL_GAE_TMA && !tma_ok) {
    set_error("gae: TMA variant needs N %% 16 == 0, N >= 32 and 16-byte aligned arrays "
              "(T=%lld N=%lld)", (long long)T, (long long)N);
    return DERL_E_INVALID;
  }
  const bool use_tma = variant == DERL_GAE_TMA || (variant == DERL_GAE_AUTO && tma_ok);
  const double gl = gamma * lambda;  // (1*gamma)*lambda, the reference's association (:62)
  if (use_tma) {
    return rewards_f64 ? launch_tma_auto<double>(rewards, values, resets, last_value, T, N,
                                                 gamma, gl, adv, vt, stats, workspace, st)
                       : launch_tma_auto<float>(rewards, values, resets, last_value, T, N,
                                                gamma, gl, adv, vt, stats, workspace, st);
  }
  return rewards_f64 ? launch_direct<double>(rewards, values, resets, last_value, T, N, gamma, gl,
                                             adv, vt, stats, workspace, st)
                     : launch_direct<float>(rewards, values, resets, last_value, T, N, gamma, gl,
                                            adv, vt, stats, workspace, st);
}

int derl_b200_moments(const float* x, int64_t count, double* stats, void* workspace,
                      size_t workspace_bytes, void* stream) {
  DERL_REQUIRE(count >= 1 && x && stats && workspace, "moments: bad arguments");
  if (workspace_bytes < derl_b200_moments_workspace_bytes(count)) {
    set_error("moments: workspace %zu B too small", workspace_bytes);
    return DERL_E_WORKSPACE;
  }
  int rc = require_device();
  if (rc != DERL_OK) return rc;
  cudaStream_t st = as_stream(stream);
  DERL_CUDA(cudaMemsetAsync(workspace, 0, kTicketBytes, st));
  moments_kernel<<<reduce_grid(count, 256), 256, 0, st>>>(x, count, workspace, stats);
  DERL_LAUNCH_CHECK("moments_kernel");
  return DERL_OK;
}

int derl_b200_normalize(const float* x, float* out, int64_t count, const double* stats,
                        double epsilon, void* stream) {
  DERL_REQUIRE(count >= 1 && x && out && stats, "normalize: bad arguments");
  int rc = require_device();
  if (rc != DERL_OK) return rc;
  const long long blocks = (count + 255) / 256;
  const long long cap = (long long)sm_count() * 16;
  normalize_kernel<<<(unsigned)(blocks < cap ? blocks : cap), 256, 0, as_stream(stream)>>>(
      x, out, count, stats, epsilon);
  DERL_LAUNCH_CHECK("normalize_kernel");
  return DERL_OK;
}

}  // extern "C"
