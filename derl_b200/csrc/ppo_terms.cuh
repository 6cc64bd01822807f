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

// The loss and the logged scalars from the minibatch sums (one thread).
__device__ __forceinline__ void write_loss(const double (&tot)[kAcc], const LossScalars& k,
                                           bool has_policy, bool has_value, float* loss,
                                           float* stats) {
  const double b = (double)k.B;
  const double pol = tot[kPol] / b, ent = tot[kEnt] / b, val = tot[kVal] / b;
  double total = 0.0;
  if (has_policy) total += pol - k.ecoef * ent;
  if (has_value) total += k.vcoef * val;
  // r_squared(targets, predictions) = 1 - mean((p - t)^2) / var_unbiased(p)   (alg/common.py:9-12);
  // PPO passes predictions = values (ppo.py:94), A2C passes predictions = value_targets (a2c.py:62)
  const double mean_v = tot[kV] / b;
  const double mean_p = (k.a2c ? tot[kVt] : tot[kV]) / b;
  const double var_v = ((k.a2c ? tot[kVtsq] : tot[kVsq]) - b * mean_p * mean_p) / (b - 1.0);
  loss[0] = (float)total;
  stats[0] = (float)total;
  stats[1] = (float)pol;
  stats[2] = (float)ent;
  stats[3] = (float)val;
  stats[4] = (float)(tot[kAdv] / b);
  stats[5] = (float)(tot[kVt] / b);
  stats[6] = (float)mean_v;
  stats[7] = (float)(1.0 - (tot[kResid] / b) / var_v);
  stats[8] = (float)(tot[kClipFrac] / b);
  stats[9] = (float)(tot[kKl] / b);
  for (int i = 10; i < DERL_LOSS_STATS; ++i) stats[i] = 0.f;
}

LossScalars make_scalars(long long B, int has_clip, double clip, double vcoef, double ecoef,
                         int a2c = 0);

}  // namespace derl
