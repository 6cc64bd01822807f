// Per-sample terms of the PPO / A2C losses, shared by the loss kernels (ppo_loss.cu, K3) and the
// fused MLP update kernel (mlp_update.cu, K8).  Closed forms: SURVEY.md §8a (a14-a17);
// reference: derl/alg/ppo.py:31-64 (policy), :73-98 (value), torch.max / clamp tie semantics.
#pragma once

#include "common.cuh"

namespace derl {

constexpr int kAcc = 11;
enum { kPol = 0, kEnt, kVal, kAdv, kVt, kV, kVsq, kResid, kClipFrac, kKl, kVtsq };

struct LossScalars {
  long long B;
  int a2c;               // 1: advantage actor-critic policy term -log_prob * adv (derl/alg/a2c.py:31)
  int has_clip;
  float lo, hi, vclip;   // 1-clip, 1+clip, clip (rounded to float32 like torch.clamp's scalars)
  float inv_b;           // 1 / B
  double vcoef, ecoef;
};

// d loss / d log_prob for one sample (includes 1/B); accumulates the policy-side sums.
__device__ __forceinline__ float surrogate(float lp, float old_lp, float adv,
                                           const LossScalars& k, double (&acc)[kAcc]) {
  if (k.a2c) {
    acc[kPol] += (double)(-(lp * adv));
    acc[kAdv] += (double)adv;
    return -adv * k.inv_b;
  }
  const float ratio = expf(lp - old_lp);
  const float s1 = -ratio * adv;
  float pol = s1, w = 1.f;
  if (k.has_clip) {
    const float rc = fminf(fmaxf(ratio, k.lo), k.hi);
    const float s2 = -rc * adv;
    const bool in_range = ratio >= k.lo && ratio <= k.hi;
    pol = fmaxf(s1, s2);
    w = s1 > s2 ? 1.f : (s1 == s2 ? (in_range ? 1.f : 0.5f) : (in_range ? 1.f : 0.f));
    acc[kClipFrac] += in_range ? 0.0 : 1.0;
  }
  acc[kPol] += (double)pol;
  acc[kAdv] += (double)adv;
  acc[kKl] += (double)(old_lp - lp);
  return -adv * ratio * w * k.inv_b;
}

// d value_loss / d v for one sample (without vcoef/B); accumulates the value-side sums.
__device__ __forceinline__ float value_term(float v, float vt, float vold, const LossScalars& k,
                                            double (&acc)[kAcc]) {
  const float u = v - vt;
  const float l1 = u * u;
  float l = l1, dv = 2.f * u;
  if (k.has_clip) {
    const float d = v - vold;
    const float dc = fminf(fmaxf(d, -k.vclip), k.vclip);
    const float w2 = (vold + dc) - vt;
    const float l2 = w2 * w2;
    const float pass = (d >= -k.vclip && d <= k.vclip) ? 1.f : 0.f;
    l = fmaxf(l1, l2);
    dv = l1 > l2 ? 2.f * u : (l2 > l1 ? 2.f * w2 * pass : u + w2 * pass);
  }
  acc[kVal] += (double)l;
  acc[kVt] += (double)vt;
  acc[kV] += (double)v;
  acc[kVsq] += (double)v * (double)v;
  acc[kVtsq] += (double)vt * (double)vt;
  acc[kResid] += (double)l1;
  return dv;
}

LossScalars make_scalars(long long B, int has_clip, double clip, double vcoef, double ecoef,
                         int a2c = 0);

}  // namespace derl
