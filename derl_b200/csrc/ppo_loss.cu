// K3 — fused PPO loss, forward + backward in one launch, categorical and diagonal-Gaussian.
//
// Replaces PPOLoss.__call__ (derl/alg/ppo.py:100-108: policy_loss :31-64, value_loss :73-98),
// the torch.distributions log_prob/entropy calls behind it (derl/policies.py:42,64-66) and
// the autograd graph hanging off them — ~60 ATen kernels forward and ~80 backward in the
// reference — with a single pass over the minibatch:
//   one thread per sample; the [rows x A] logits (or loc/scale/actions) tile is staged
//   through padded shared memory so global traffic is coalesced both ways; per-sample terms
//   are float32 in torch's operation order; the means are float64, reduced by warp shuffles
//   -> block -> per-block partials -> last block in fixed order (no float atomics, so the
//   loss is run-to-run reproducible); gradients w.r.t. logits / loc / scale / values are
//   written in the same pass, already carrying the 1/B of the means.
// torch.max tie semantics (gradient split 1/2 : 1/2, torch/tools/autograd/derivatives.yaml
// `maximum`) and clamp's closed-interval pass-through mask are reproduced exactly.
#include <math_constants.h>

#include "common.cuh"
#include "ppo_terms.cuh"

namespace derl {
namespace {

// Last block: turn the ten sums into the loss and the logged scalars.
__device__ __forceinline__ void finish_loss(double (&acc)[kAcc], const LossScalars& k,
                                            bool has_policy, bool has_value, void* workspace,
                                            float* loss, float* stats, double* scratch,
                                            int* flag) {
  block_sum<kAcc>(acc, scratch);
  if (!publish_partials<kAcc>(acc, workspace, flag)) return;
  double tot[kAcc];
  final_sum<kAcc>(tot, workspace, scratch);
  if (threadIdx.x != 0) return;
  write_loss(tot, k, has_policy, has_value, loss, stats);
}

// Coalesced copy of `rows` consecutive rows of width `w` between global memory (dense) and a
// shared tile with row pitch `pitch` (odd, so that thread-per-row access is conflict-free).
template <bool kToShared>
__device__ __forceinline__ void tile_copy(float* tile, float* gmem, int rows, int w, int pitch) {
  const int n = rows * w;
  for (int e = threadIdx.x; e < n; e += blockDim.x) {
    const int r = e / w, c = e - r * w;
    if (kToShared) {
      tile[r * pitch + c] = __ldg(gmem + e);
    } else {
      gmem[e] = tile[r * pitch + c];
    }
  }
}

// ------------------------------------------------------------------ categorical head
__global__ void __launch_bounds__(128)
ppo_loss_categorical_kernel(const float* __restrict__ logits, int A,
                            const long long* __restrict__ actions,
                            const float* __restrict__ old_logp, const float* __restrict__ adv,
                            const float* __restrict__ values, const float* __restrict__ vtarg,
                            const float* __restrict__ vold, LossScalars k,
                            float* __restrict__ loss, float* __restrict__ dlogits,
                            float* __restrict__ dvalues, float* __restrict__ stats,
                            void* workspace) {
  extern __shared__ float tile[];
  __shared__ double scratch[kAcc * 32];
  __shared__ int flag;
  const bool has_policy = logits != nullptr, has_value = values != nullptr;
  const int R = blockDim.x;
  const int pitch = A | 1;
  double acc[kAcc];
#pragma unroll
  for (int i = 0; i < kAcc; ++i) acc[i] = 0.0;

  for (long long row0 = (long long)blockIdx.x * R; row0 < k.B; row0 += (long long)gridDim.x * R) {
    const int rows = (int)(k.B - row0 < R ? k.B - row0 : R);
    const long long i = row0 + threadIdx.x;
    const bool live = threadIdx.x < rows;
    if (has_policy) {
      tile_copy<true>(tile, const_cast<float*>(logits) + row0 * A, rows, A, pitch);
      __syncthreads();
      if (live) {
        float* z = tile + threadIdx.x * pitch;
        float m = -CUDART_INF_F;
        for (int c = 0; c < A; ++c) m = fmaxf(m, z[c]);
        float s = 0.f;
        for (int c = 0; c < A; ++c) s += expf(z[c] - m);
        const float lse = m + logf(s);
        const float inv_s = 1.f / s;
        int a = (int)__ldg(actions + i);
        if ((unsigned)a >= (unsigned)A) a = 0;  // memory safety only; callers validate
        const float lp = z[a] - lse;
        float h = 0.f;
        for (int c = 0; c < A; ++c) {
          const float ln = z[c] - lse;
          h -= (expf(z[c] - m) * inv_s) * fmaxf(ln, -3.4028234663852886e38f);
        }
        acc[kEnt] += (double)h;
        const float g = surrogate(lp, k.a2c ? 0.f : __ldg(old_logp + i), __ldg(adv + i), k, acc);
        const float ce = (float)k.ecoef * k.inv_b;
        for (int c = 0; c < A; ++c) {
          const float ln = z[c] - lse;
          const float p = expf(z[c] - m) * inv_s;
          z[c] = g * ((c == a ? 1.f : 0.f) - p) + ce * p * (ln + h);
        }
      }
      __syncthreads();
      tile_copy<false>(tile, dlogits + row0 * A, rows, A, pitch);
      __syncthreads();
    }
    if (has_value && live) {
      const float dv = value_term(__ldg(values + i), __ldg(vtarg + i),
                                  k.has_clip ? __ldg(vold + i) : 0.f, k, acc);
      dvalues[i] = (float)k.vcoef * k.inv_b * dv;
    }
  }
  finish_loss(acc, k, has_policy, has_value, workspace, loss, stats, scratch, &flag);
}

// ------------------------------------------------------------------ diagonal-Gaussian head
__global__ void __launch_bounds__(128)
ppo_loss_gaussian_kernel(const float* __restrict__ loc, const float* __restrict__ scale, int D,
                         const float* __restrict__ actions, const float* __restrict__ old_logp,
                         const float* __restrict__ adv, const float* __restrict__ values,
                         const float* __restrict__ vtarg, const float* __restrict__ vold,
                         LossScalars k, float* __restrict__ loss, float* __restrict__ dloc,
                         float* __restrict__ dscale, float* __restrict__ dvalues,
                         float* __restrict__ stats, void* workspace) {
  extern __shared__ float tile[];
  __shared__ double scratch[kAcc * 32];
  __shared__ int flag;
  const bool has_policy = loc != nullptr, has_value = values != nullptr;
  const int R = blockDim.x;
  const int pitch = D | 1;
  float* t_loc = tile;
  float* t_scale = tile + (size_t)R * pitch;
  float* t_act = tile + (size_t)2 * R * pitch;
  double acc[kAcc];
#pragma unroll
  for (int i = 0; i < kAcc; ++i) acc[i] = 0.0;
  const float kLogSqrt2Pi = 0.918938533204672742f;  // log(sqrt(2*pi))
  const float kHalfLog2PiE = 1.418938533204672742f; // 0.5 + 0.5*log(2*pi)

  for (long long row0 = (long long)blockIdx.x * R; row0 < k.B; row0 += (long long)gridDim.x * R) {
    const int rows = (int)(k.B - row0 < R ? k.B - row0 : R);
    const long long i = row0 + threadIdx.x;
    const bool live = threadIdx.x < rows;
    if (has_policy) {
      tile_copy<true>(t_loc, const_cast<float*>(loc) + row0 * D, rows, D, pitch);
      tile_copy<true>(t_scale, const_cast<float*>(scale) + row0 * D, rows, D, pitch);
      tile_copy<true>(t_act, const_cast<float*>(actions) + row0 * D, rows, D, pitch);
      __syncthreads();
      if (live) {
        float* mu = t_loc + threadIdx.x * pitch;
        float* sg = t_scale + threadIdx.x * pitch;
        const float* ac = t_act + threadIdx.x * pitch;
        float lp = 0.f, h = 0.f;
        for (int c = 0; c < D; ++c) {
          const float sd = sg[c], diff = ac[c] - mu[c];
          const float log_sd = logf(sd);
          lp += -(diff * diff) / (2.f * (sd * sd)) - log_sd - kLogSqrt2Pi;
          h += kHalfLog2PiE + log_sd;
        }
        acc[kEnt] += (double)h;
        const float g = surrogate(lp, k.a2c ? 0.f : __ldg(old_logp + i), __ldg(adv + i), k, acc);
        const float ce = (float)k.ecoef * k.inv_b;
        for (int c = 0; c < D; ++c) {
          const float sd = sg[c], diff = ac[c] - mu[c];
          const float inv_sd = 1.f / sd;
          const float zsc = diff * inv_sd;  // (a - mu) / sigma
          mu[c] = g * zsc * inv_sd;
          sg[c] = g * (zsc * zsc - 1.f) * inv_sd - ce * inv_sd;
        }
      }
      __syncthreads();
      tile_copy<false>(t_loc, dloc + row0 * D, rows, D, pitch);
      tile_copy<false>(t_scale, dscale + row0 * D, rows, D, pitch);
      __syncthreads();
    }
    if (has_value && live) {
      const float dv = value_term(__ldg(values + i), __ldg(vtarg + i),
                                  k.has_clip ? __ldg(vold + i) : 0.f, k, acc);
      dvalues[i] = (float)k.vcoef * k.inv_b * dv;
    }
  }
  finish_loss(acc, k, has_policy, has_value, workspace, loss, stats, scratch, &flag);
}

constexpr int kMaxTileBytes = 160 * 1024;

// rows per CTA: as many as fit the tile budget, between 32 and 128
int pick_rows(long long width, int tiles, size_t* smem_bytes) {
  const size_t pitch = (size_t)(width | 1);
  int R = 128;
  while (R > 32 && (size_t)tiles * R * pitch * sizeof(float) > (size_t)kMaxTileBytes) R >>= 1;
  *smem_bytes = (size_t)tiles * R * pitch * sizeof(float);
  return R;
}

unsigned loss_grid(long long B, int R) {
  long long blocks = (B + R - 1) / R;
  long long cap = (long long)sm_count() * 8;
  if (cap > kMaxReduceBlocks) cap = kMaxReduceBlocks;
  if (blocks > cap) blocks = cap;
  return (unsigned)(blocks < 1 ? 1 : blocks);
}

int check_common(const char* who, long long B, const void* head, const float* old_logp,
                 const float* adv, const float* values, const float* vtarg, const float* vold,
                 const void* dvalues, const float* loss, const float* stats, void* workspace,
                 size_t workspace_bytes, int a2c = 0, int has_clip = 1) {
  DERL_REQUIRE(B >= 1, "%s: B must be >= 1 (got %lld)", who, B);
  DERL_REQUIRE(head != nullptr || values != nullptr, "%s: both heads are NULL", who);
  if (head != nullptr) {
    DERL_REQUIRE(adv != nullptr, "%s: policy head needs advantages", who);
    DERL_REQUIRE(a2c || old_logp != nullptr, "%s: policy head needs old_logp", who);
  }
  if (values != nullptr) {
    DERL_REQUIRE(vtarg && dvalues, "%s: value head needs value_targets and dvalues", who);
    DERL_REQUIRE(!has_clip || vold != nullptr, "%s: clipped value loss needs old_values", who);
  } else {
    DERL_REQUIRE(dvalues == nullptr, "%s: dvalues given without values", who);
  }
  DERL_REQUIRE(loss && stats && workspace, "%s: loss/stats/workspace must not be NULL", who);
  if (workspace_bytes < derl_b200_ppo_loss_workspace_bytes(B)) {
    set_error("%s: workspace %zu B < required %zu B", who, workspace_bytes,
              derl_b200_ppo_loss_workspace_bytes(B));
    return DERL_E_WORKSPACE;
  }
  return DERL_OK;
}

// One body for the PPO and A2C entry points of each head (a2c: no ratio / no clipping).
int launch_categorical(const char* who, int a2c, const float* logits, long long B, long long A,
                       const int64_t* actions, const float* old_logp, const float* adv,
                       const float* values, const float* vtarg, const float* vold, int has_clip,
                       double clip, double vcoef, double ecoef, float* loss, float* dlogits,
                       float* dvalues, float* stats, void* workspace, size_t workspace_bytes,
                       void* stream) {
  int rc = check_common(who, B, logits, old_logp, adv, values, vtarg, vold, dvalues, loss, stats,
                        workspace, workspace_bytes, a2c, has_clip);
  if (rc != DERL_OK) return rc;
  if (logits != nullptr) {
    DERL_REQUIRE(A >= 1 && A <= 1024, "%s: A=%lld outside [1, 1024]", who, A);
    DERL_REQUIRE(actions && dlogits, "%s: actions/dlogits are NULL", who);
  } else {
    DERL_REQUIRE(dlogits == nullptr, "%s: dlogits given without logits", who);
    A = 1;
  }
  if ((rc = require_device()) != DERL_OK) return rc;
  cudaStream_t st = as_stream(stream);
  size_t smem = 0;
  const int R = pick_rows(A, 1, &smem);
  if ((rc = ensure_dynamic_smem(reinterpret_cast<const void*>(ppo_loss_categorical_kernel),
                                kMaxTileBytes)) != DERL_OK)
    return rc;
  DERL_CUDA(cudaMemsetAsync(workspace, 0, kTicketBytes, st));
  ppo_loss_categorical_kernel<<<loss_grid(B, R), R, smem, st>>>(
      logits, (int)A, reinterpret_cast<const long long*>(actions), old_logp, adv, values, vtarg,
      vold, make_scalars(B, has_clip, clip, vcoef, ecoef, a2c), loss, dlogits, dvalues, stats,
      workspace);
  DERL_LAUNCH_CHECK("ppo_loss_categorical_kernel");
  return DERL_OK;
}

int launch_gaussian(const char* who, int a2c, const float* loc, const float* scale, long long B,
                    long long D, const float* actions, const float* old_logp, const float* adv,
                    const float* values, const float* vtarg, const float* vold, int has_clip,
                    double clip, double vcoef, double ecoef, float* loss, float* dloc,
                    float* dscale, float* dvalues, float* stats, void* workspace,
                    size_t workspace_bytes, void* stream) {
  int rc = check_common(who, B, loc, old_logp, adv, values, vtarg, vold, dvalues, loss, stats,
                        workspace, workspace_bytes, a2c, has_clip);
  if (rc != DERL_OK) return rc;
  if (loc != nullptr) {
    DERL_REQUIRE(D >= 1 && D <= 256, "%s: D=%lld outside [1, 256]", who, D);
    DERL_REQUIRE(scale && actions && dloc && dscale, "%s: scale/actions/dloc/dscale are NULL", who);
  } else {
    DERL_REQUIRE(dloc == nullptr && dscale == nullptr, "%s: gradients given without loc", who);
    D = 1;
  }
  if ((rc = require_device()) != DERL_OK) return rc;
  cudaStream_t st = as_stream(stream);
  size_t smem = 0;
  const int R = pick_rows(D, 3, &smem);
  if ((rc = ensure_dynamic_smem(reinterpret_cast<const void*>(ppo_loss_gaussian_kernel),
                                kMaxTileBytes)) != DERL_OK)
    return rc;
  DERL_CUDA(cudaMemsetAsync(workspace, 0, kTicketBytes, st));
  ppo_loss_gaussian_kernel<<<loss_grid(B, R), R, smem, st>>>(
      loc, scale, (int)D, actions, old_logp, adv, values, vtarg, vold,
      make_scalars(B, has_clip, clip, vcoef, ecoef, a2c), loss, dloc, dscale, dvalues, stats,
      workspace);
  DERL_LAUNCH_CHECK("ppo_loss_gaussian_kernel");
  return DERL_OK;
}

}  // namespace


LossScalars make_scalars(long long B, int has_clip, double clip, double vcoef, double ecoef,
                         int a2c) {
  LossScalars k;
  k.B = B;
  k.a2c = a2c;
  k.has_clip = has_clip ? 1 : 0;
  k.lo = (float)(1.0 - clip);
  k.hi = (float)(1.0 + clip);
  k.vclip = (float)clip;
  k.inv_b = 1.0f / (float)B;
  k.vcoef = vcoef;
  k.ecoef = ecoef;
  return k;
}

}  // namespace derl

using namespace derl;

extern "C" {

size_t derl_b200_ppo_loss_workspace_bytes(int64_t B) {
  (void)B;
  return kTicketBytes + (size_t)kMaxReduceBlocks * kAcc * sizeof(double);
}

int derl_b200_ppo_loss_categorical(const float* logits, int64_t B, int64_t A,
                                   const int64_t* actions, const float* old_logp,
                                   const float* adv, const float* values, const float* vtarg,
                                   const float* vold, int has_clip, double clip, double vcoef,
                                   double ecoef, float* loss, float* dlogits, float* dvalues,
                                   float* stats, void* workspace, size_t workspace_bytes,
                                   void* stream) {
  return launch_categorical("ppo_loss_categorical", 0, logits, B, A, actions, old_logp, adv, values,
                            vtarg, vold, has_clip, clip, vcoef, ecoef, loss, dlogits, dvalues, stats,
                            workspace, workspace_bytes, stream);
}

int derl_b200_ppo_loss_gaussian(const float* loc, const float* scale, int64_t B, int64_t D,
                                const float* actions, const float* old_logp, const float* adv,
                                const float* values, const float* vtarg, const float* vold,
                                int has_clip, double clip, double vcoef, double ecoef, float* loss,
                                float* dloc, float* dscale, float* dvalues, float* stats,
                                void* workspace, size_t workspace_bytes, void* stream) {
  return launch_gaussian("ppo_loss_gaussian", 0, loc, scale, B, D, actions, old_logp, adv, values,
                         vtarg, vold, has_clip, clip, vcoef, ecoef, loss, dloc, dscale, dvalues,
                         stats, workspace, workspace_bytes, stream);
}

/* Advantage actor-critic loss (derl/alg/a2c.py:19-79) on the same kernels: policy term
 * -mean(log_prob * adv), unclipped value loss, same entropy term, gradients and logged scalars. */
int derl_b200_a2c_loss_categorical(const float* logits, int64_t B, int64_t A,
                                   const int64_t* actions, const float* adv, const float* values,
                                   const float* vtarg, double vcoef, double ecoef, float* loss,
                                   float* dlogits, float* dvalues, float* stats, void* workspace,
                                   size_t workspace_bytes, void* stream) {
  return launch_categorical("a2c_loss_categorical", 1, logits, B, A, actions, nullptr, adv, values,
                            vtarg, nullptr, 0, 0.0, vcoef, ecoef, loss, dlogits, dvalues, stats,
                            workspace, workspace_bytes, stream);
}

int derl_b200_a2c_loss_gaussian(const float* loc, const float* scale, int64_t B, int64_t D,
                                const float* actions, const float* adv, const float* values,
                                const float* vtarg, double vcoef, double ecoef, float* loss,
                                float* dloc, float* dscale, float* dvalues, float* stats,
                                void* workspace, size_t workspace_bytes, void* stream) {
  return launch_gaussian("a2c_loss_gaussian", 1, loc, scale, B, D, actions, nullptr, adv, values,
                         vtarg, nullptr, 0, 0.0, vcoef, ecoef, loss, dloc, dscale, dvalues, stats,
                         workspace, workspace_bytes, stream);
}

}  // extern "C"
