// K8 — one launch = one whole PPO update of the MuJoCo-shaped actor-critic (two tanh MLPs
// obs -> 64 -> 64 -> {D, 1} with a state-independent log-std): every epoch, every minibatch,
//   permutation gather -> advantage normalisation -> both forward passes -> fused PPO loss
//   (diagonal Gaussian, forward + backward) -> both backward passes -> clip_grad_norm_ -> Adam.
//
// Replaces, for BASELINE configs[1] (1 env x 2048 steps, 10 epochs x 32 minibatches of 64), the
// reference's 320 trips per update around IterateWithMinibatches.run (derl/runners/onpolicy.py:
// 51-62), NormalizeAdvantages (trajectory_transforms.py:89-92), MuJoCoModel.forward
// (derl/models.py:261-271), PPOLoss (derl/alg/ppo.py:100-108), Trainer.step (derl/alg/common.py:
// 66-78: backward, clip_grad_norm_, optimizer.step) — ~150 ATen launches each.  At a minibatch
// of 64 rows and 11 k parameters none of that is bandwidth or FLOP work: the eager path spends
// 1.6 ms per optimiser step on launches and Python, a CUDA-graph replay 0.25 ms.  Here the
// whole update is one persistent CTA: parameters and their gradients live in shared memory for
// all 320 steps (row pitch 68 floats: 16-byte aligned rows whose float4 loads spread over all
// bank groups), a step is
// ~30 k cycles of fp32 FMA work, and Adam's moment vectors are the only per-step global traffic
// (88 KB, L2-resident, kept in the same padded layout as the shared parameter block so that the
// update loop is one division-free sweep with independent loads).  fp32 FMA on CUDA cores, not tensor cores: the contract is the
// reference's float32 (rtol 1e-5 on the loss), and [64 x 64 x 64] GEMMs on one SM are
// latency-, not throughput-bound.
//
// Thread roles: 512 threads; threads 0-255 work on the policy network, 256-511 on the value
// network.  A [64 x 64 x 64] GEMM is 256 threads x (4 x 4) register tile with float4 operand
// loads — along the reduction index where both operands are contiguous in it (forward),
// along the tile otherwise (dgrad, wgrad): 8 LDS.128 per 64 FMA.  The first profile (ncu:
// `not_selected` and `barrier` on top, i.e. issue-bound on one SM) had 8 scalar LDS + 8 address
// IMADs per 16 FMA and a full 64 x 64 tile for the [64 x obs] weight gradient.
#include <math_constants.h>

#include "common.cuh"
#include "ppo_terms.cuh"

namespace derl {
namespace {

constexpr int kH = 64;          // hidden width of both MLPs
constexpr int kHP = kH + 4;     // pitch of a [*, 64] array in shared memory: 16-byte aligned rows
constexpr int kRows = 64;       // rows per pass (a larger minibatch accumulates over passes)
constexpr int kThreads = 512;
constexpr int kNet = 256;       // threads per network
constexpr int kTensors = 13;    // W1 b1 W2 b2 W3 b3 (policy), the same (value), logstd
constexpr int kMaxAct = 32;

struct MlpTensor {
  float* w;
  float* m;
  float* v;
  int rows, cols, off, pitch;   // off / pitch: position inside the shared parameter block
};

// Shared-memory plan, identical on host and device (float offsets unless noted).
struct Carve {
  int OP, DP;                   // pitches of [*, obs] (multiple of 4, zero padded) and [*, act] (odd)
  int params;                   // floats of one parameter block (W or G)
  int off_w, off_g, off_x, off_h[2][3], off_adv, off_oldlp, off_vt, off_vold, off_act, off_loc,
      off_val, off_dloc, off_dscale, off_dval, off_sd, off_misc, off_idx, off_red, total;
  int t_off[kTensors], t_pitch[kTensors], t_rows[kTensors], t_cols[kTensors];

  __host__ __device__ Carve(int O, int D) {
    OP = (O + 3) & ~3;
    DP = D | 1;
    int o = 0;
    for (int net = 0; net < 2; ++net) {
      const int out = net == 0 ? D : 1;
      const int rows[6] = {kH, kH, kH, kH, out, out};
      const int cols[6] = {O, 1, kH, 1, kH, 1};
      const int pitch[6] = {OP, 1, kHP, 1, kHP, 1};
      for (int i = 0; i < 6; ++i) {
        const int t = net * 6 + i;
        t_off[t] = o;
        t_pitch[t] = pitch[i];
        t_rows[t] = rows[i];
        t_cols[t] = cols[i];
        o += (rows[i] * pitch[i] + 3) & ~3;     // every tensor starts 16-byte aligned
      }
    }
    t_off[12] = o;
    t_pitch[12] = 1;
    t_rows[12] = D;
    t_cols[12] = 1;
    o += D;
    params = (o + 3) & ~3;
    o = 0;
    off_w = o; o += params;
    off_g = o; o += params;
    off_x = o; o += kRows * OP;
    for (int net = 0; net < 2; ++net)
      for (int b = 0; b < 3; ++b) { off_h[net][b] = o; o += kRows * kHP; }
    off_adv = o; o += kRows;
    off_oldlp = o; o += kRows;
    off_vt = o; o += kRows;
    off_vold = o; o += kRows;
    off_val = o; o += kRows;
    off_dval = o; o += kRows;
    off_act = o; o += kRows * DP;
    off_loc = o; o += kRows * DP;
    off_dloc = o; o += kRows * DP;
    off_dscale = o; o += kRows * DP;
    o = (o + 3) & ~3;
    off_sd = o; o += kMaxAct;
    off_misc = o; o += 16;                       // mean, denom, clip coefficient, Adam scalars
    o = (o + 1) & ~1;
    off_idx = o; o += 2 * kRows;                 // long long[64]
    off_red = o; o += 2 * kAcc * 32;             // double[kAcc * 32]
    total = o;
  }
  __host__ __device__ size_t bytes() const { return (size_t)total * sizeof(float); }
};

struct MlpArgs {
  MlpTensor t[kTensors];
  const void* obs;
  int obs_f64, O, D;
  const float *actions, *old_logp, *adv, *vtarg, *vold;
  const long long* perm;
  long long S, mb, step0;
  int nsteps, nmb, normalize, has_clip;
  double adv_eps, clip, vcoef, ecoef, max_norm, lr, beta1, beta2, eps;
  float* losses;
  float* stats;
  float* ws_m;     // Adam moments in the PADDED shared-memory layout (Carve::params floats each):
  float* ws_v;     // element e of the workspace belongs to element e of the parameter block
};

__device__ __forceinline__ float4 ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }

// The three GEMM shapes of a 64-row pass, each by the 256 threads of one net (t = 0..255,
// tm = t / 16, tn = t % 16), 4 x 4 outputs per thread, operands read as float4.
//
// kk: C(m, n) = sum_k A[m][k] B[n][k], both operands contiguous along the reduction index
//     (forward layers: activations [row][k], weights [unit][k]).  m = tm + 16 i, n = tn + 16 j:
//     a warp reads 2 distinct A chunks and 16 B chunks in 8 distinct bank groups.  K % 4 == 0.
template <typename Epi>
__device__ __forceinline__ void gemm_kk(const float* __restrict__ A, int a_p,
                                        const float* __restrict__ B, int b_p, int K, int t,
                                        Epi epi) {
  const int tm = t >> 4, tn = t & 15;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  const float* a0 = A + tm * a_p;
  const float* b0 = B + tn * b_p;
#pragma unroll 2
  for (int k = 0; k < K; k += 4) {
    float4 a[4], b[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      a[i] = ld4(a0 + 16 * i * a_p + k);
      b[i] = ld4(b0 + 16 * i * b_p + k);
    }
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        acc[i][j] = fmaf(a[i].x, b[j].x, acc[i][j]);
        acc[i][j] = fmaf(a[i].y, b[j].y, acc[i][j]);
        acc[i][j] = fmaf(a[i].z, b[j].z, acc[i][j]);
        acc[i][j] = fmaf(a[i].w, b[j].w, acc[i][j]);
      }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) epi(tm + 16 * i, tn + 16 * j, acc[i][j]);
}

// kn: C(m, n) = sum_r A[m][r] B[r][n]: A contiguous along the reduction, B along the tile
//     (dgrad: dZ [row][unit] x W [unit][k]).  m = tm + 16 i, n = 4 tn .. 4 tn + 3.  R % 4 == 0.
template <typename Epi>
__device__ __forceinline__ void gemm_kn(const float* __restrict__ A, int a_p,
                                        const float* __restrict__ B, int b_p, int R, int t,
                                        Epi epi) {
  const int tm = t >> 4, tn = t & 15;
  float4 acc[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) acc[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  const float* a0 = A + tm * a_p;
  const float* b0 = B + 4 * tn;
#pragma unroll 2
  for (int r = 0; r < R; r += 4) {
    float4 a[4], b[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      a[i] = ld4(a0 + 16 * i * a_p + r);
      b[i] = ld4(b0 + (r + i) * b_p);
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float ar[4] = {a[i].x, a[i].y, a[i].z, a[i].w};
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        acc[i].x = fmaf(ar[e], b[e].x, acc[i].x);
        acc[i].y = fmaf(ar[e], b[e].y, acc[i].y);
        acc[i].z = fmaf(ar[e], b[e].z, acc[i].z);
        acc[i].w = fmaf(ar[e], b[e].w, acc[i].w);
      }
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) epi(tm + 16 * i, 4 * tn, acc[i]);
}

// nn: C(m, n) += sum_r A[r][m] B[r][n] over rows [r0, r1): both operands contiguous along the
//     tile (weight gradients: dZ [row][unit] x activations [row][k]).  m = 4 tm .., n = 4 tn ..
template <typename Epi>
__device__ __forceinline__ void gemm_nn(const float* __restrict__ A, int a_p,
                                        const float* __restrict__ B, int b_p, int r0, int r1,
                                        int tm, int tn, Epi epi) {
  float4 acc[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) acc[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  const float* a0 = A + 4 * tm;
  const float* b0 = B + 4 * tn;
#pragma unroll 4
  for (int r = r0; r < r1; ++r) {
    const float4 a = ld4(a0 + r * a_p), b = ld4(b0 + r * b_p);
    const float ar[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      acc[i].x = fmaf(ar[i], b.x, acc[i].x);
      acc[i].y = fmaf(ar[i], b.y, acc[i].y);
      acc[i].z = fmaf(ar[i], b.z, acc[i].z);
      acc[i].w = fmaf(ar[i], b.w, acc[i].w);
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) epi(4 * tm + i, 4 * tn, acc[i]);
}

__global__ void __launch_bounds__(kThreads, 1) ppo_mlp_update_kernel(const MlpArgs a) {
  extern __shared__ __align__(16) float sm[];
  const Carve c(a.O, a.D);
  const int tid = threadIdx.x, net = tid >> 8, t = tid & (kNet - 1);
  const int O = a.O, D = a.D, OP = c.OP, DP = c.DP;
  float* W = sm + c.off_w;
  float* G = sm + c.off_g;
  float* X = sm + c.off_x;
  float* H1 = sm + c.off_h[net][0];
  float* H2 = sm + c.off_h[net][1];
  float* Sb = sm + c.off_h[net][2];
  float* s_adv = sm + c.off_adv;
  float* s_oldlp = sm + c.off_oldlp;
  float* s_vt = sm + c.off_vt;
  float* s_vold = sm + c.off_vold;
  float* s_val = sm + c.off_val;
  float* s_dval = sm + c.off_dval;
  float* s_act = sm + c.off_act;
  float* s_loc = sm + c.off_loc;
  float* s_dloc = sm + c.off_dloc;
  float* s_dscale = sm + c.off_dscale;
  float* s_sd = sm + c.off_sd;
  float* misc = sm + c.off_misc;
  long long* s_idx = reinterpret_cast<long long*>(sm + c.off_idx);
  double* red = reinterpret_cast<double*>(sm + c.off_red);
  // this net's tensors inside a parameter block
  const int tb = net * 6;
  const int oW1 = c.t_off[tb], ob1 = c.t_off[tb + 1], oW2 = c.t_off[tb + 2], ob2 = c.t_off[tb + 3];
  const int oW3p = c.t_off[4], ob3p = c.t_off[5], oW3v = c.t_off[10], ob3v = c.t_off[11];
  const int oLs = c.t_off[12];

  // ---- parameters: global (dense) -> shared (padded); Adam moments -> padded workspace
  for (int e = tid; e < c.params; e += kThreads) {   // padding entries: defined and inert
    W[e] = 0.f;
    a.ws_m[e] = 0.f;
    a.ws_v[e] = 0.f;
  }
  __syncthreads();
  for (int k = 0; k < kTensors; ++k) {
    const int n = c.t_rows[k] * c.t_cols[k];
    for (int e = tid; e < n; e += kThreads) {
      const int r = e / c.t_cols[k], col = e - r * c.t_cols[k];
      const int so = c.t_off[k] + r * c.t_pitch[k] + col;
      W[so] = a.t[k].w[e];
      a.ws_m[so] = a.t[k].m[e];
      a.ws_v[so] = a.t[k].v[e];
    }
  }
  __syncthreads();

  LossScalars ks;
  ks.a2c = 0;
  ks.has_clip = a.has_clip;
  ks.lo = (float)(1.0 - a.clip);
  ks.hi = (float)(1.0 + a.clip);
  ks.vclip = (float)a.clip;
  ks.vcoef = a.vcoef;
  ks.ecoef = a.ecoef;
  const float kLogSqrt2Pi = 0.918938533204672742f;
  const float kHalfLog2PiE = 1.418938533204672742f;

  double pow1 = pow(a.beta1, (double)a.step0), pow2 = pow(a.beta2, (double)a.step0);
  for (int step = 0; step < a.nsteps; ++step) {
    const int epoch = step / a.nmb, j = step - epoch * a.nmb;
    const long long start = (long long)j * a.mb;
    const long long count = a.S - start < a.mb ? a.S - start : a.mb;
    const long long* rows_idx = a.perm + (long long)epoch * a.S + start;
    ks.B = count;
    ks.inv_b = 1.0f / (float)count;

    // ---- minibatch moments of the advantages (float64), gradients <- 0, std = exp(logstd)
    {
      double s[2] = {0.0, 0.0};
      if (a.normalize) {
        for (long long i = tid; i < count; i += kThreads) {
          const double x = (double)__ldg(a.adv + __ldg(rows_idx + i));
          s[0] += x;
          s[1] += x * x;
        }
      }
      if (step == 0) {
        for (int e = tid; e < c.params; e += kThreads) G[e] = 0.f;   // later steps: Adam re-zeroes
      }
      if (tid < D) s_sd[tid] = expf(W[oLs + tid]);
      block_sum<2>(s, red);
      if (tid == 0) {
        float mean = 0.f, denom = 1.f;
        if (a.normalize) {   // normalize_kernel's arithmetic (gae.cu)
          const double n = (double)count, mean_d = s[0] / n;
          double var_d = s[1] / n - mean_d * mean_d;
          var_d = var_d > 0.0 ? var_d : 0.0;
          mean = (float)mean_d;
          denom = (float)((double)(float)sqrt(var_d) + a.adv_eps);
        }
        misc[0] = mean;
        misc[1] = denom;
      }
      __syncthreads();
    }
    double acc[kAcc];
#pragma unroll
    for (int i = 0; i < kAcc; ++i) acc[i] = 0.0;

    for (long long chunk = 0; chunk < count; chunk += kRows) {
      const int rows = (int)(count - chunk < kRows ? count - chunk : kRows);
      // ---- gather this pass's rows
      if (tid < kRows) {
        const bool live = tid < rows;
        const long long idx = live ? __ldg(rows_idx + chunk + tid) : 0;
        s_idx[tid] = idx;
        float adv = 0.f, olp = 0.f, vt = 0.f, vo = 0.f;
        if (live) {
          adv = __ldg(a.adv + idx);
          if (a.normalize) adv = __fdiv_rn(__fsub_rn(adv, misc[0]), misc[1]);
          olp = __ldg(a.old_logp + idx);
          vt = __ldg(a.vtarg + idx);
          vo = a.has_clip ? __ldg(a.vold + idx) : 0.f;
        }
        s_adv[tid] = adv;
        s_oldlp[tid] = olp;
        s_vt[tid] = vt;
        s_vold[tid] = vo;
      }
      __syncthreads();
      for (int e = tid; e < kRows * OP; e += kThreads) {   // padding columns k >= O are zeros
        const int r = e / OP, k = e - r * OP;
        float x = 0.f;
        if (r < rows && k < O) {
          const long long g = s_idx[r] * O + k;
          x = a.obs_f64 ? (float)__ldg(reinterpret_cast<const double*>(a.obs) + g)
                        : __ldg(reinterpret_cast<const float*>(a.obs) + g);
        }
        X[r * OP + k] = x;
      }
      for (int e = tid; e < kRows * D; e += kThreads) {
        const int r = e / D, k = e - r * D;
        s_act[r * DP + k] = r < rows ? __ldg(a.actions + s_idx[r] * D + k) : 0.f;
      }
      __syncthreads();

      // ---- forward: H1 = tanh(X W1^T + b1), H2 = tanh(H1 W2^T + b2)
      gemm_kk(X, OP, W + oW1, OP, OP, t,
              [&](int m, int n, float v) { H1[m * kHP + n] = tanhf(v + W[ob1 + n]); });
      __syncthreads();
      gemm_kk(H1, kHP, W + oW2, kHP, kH, t,
              [&](int m, int n, float v) { H2[m * kHP + n] = tanhf(v + W[ob2 + n]); });
      __syncthreads();
      // ---- heads: loc [64, D] from the policy trunk, value [64] from the value trunk
      {
        const float* H2p = sm + c.off_h[0][1];
        const float* H2v = sm + c.off_h[1][1];
        for (int o = tid; o < kRows * (D + 1); o += kThreads) {
          const int r = o & (kRows - 1), col = o >> 6;
          const float* h = (col < D ? H2p : H2v) + r * kHP;
          const float* w = W + (col < D ? oW3p + col * kHP : oW3v);
          float z = 0.f;
#pragma unroll 4
          for (int k = 0; k < kH; k += 4) {
            const float4 hv = ld4(h + k), wv = ld4(w + k);
            z = fmaf(hv.x, wv.x, z);
            z = fmaf(hv.y, wv.y, z);
            z = fmaf(hv.z, wv.z, z);
            z = fmaf(hv.w, wv.w, z);
          }
          if (col < D) s_loc[r * DP + col] = z + W[ob3p + col];
          else s_val[r] = z + W[ob3v];
        }
      }
      __syncthreads();
      // ---- PPO loss terms, one row per thread (K3's per-sample arithmetic)
      if (tid < kRows) {
        const bool live = tid < rows;
        float* mu = s_loc + tid * DP;
        const float* ac = s_act + tid * DP;
        float* dl = s_dloc + tid * DP;
        float* ds = s_dscale + tid * DP;
        float dv = 0.f;
        if (live) {
          float lp = 0.f, h = 0.f;
          for (int k = 0; k < D; ++k) {
            const float sd = s_sd[k], diff = ac[k] - mu[k];
            const float log_sd = logf(sd);
            lp += -(diff * diff) / (2.f * (sd * sd)) - log_sd - kLogSqrt2Pi;
            h += kHalfLog2PiE + log_sd;
          }
          acc[kEnt] += (double)h;
          const float g = surrogate(lp, s_oldlp[tid], s_adv[tid], ks, acc);
          const float ce = (float)ks.ecoef * ks.inv_b;
          for (int k = 0; k < D; ++k) {
            const float sd = s_sd[k], diff = ac[k] - mu[k];
            const float inv_sd = 1.f / sd;
            const float zsc = diff * inv_sd;
            dl[k] = g * zsc * inv_sd;
            ds[k] = g * (zsc * zsc - 1.f) * inv_sd - ce * inv_sd;
          }
          dv = (float)ks.vcoef * ks.inv_b * value_term(s_val[tid], s_vt[tid], s_vold[tid], ks, acc);
        } else {
          for (int k = 0; k < D; ++k) {
            dl[k] = 0.f;
            ds[k] = 0.f;
          }
        }
        s_dval[tid] = dv;
      }
      __syncthreads();
      // ---- head gradients: W3 / b3 of both nets, logstd
      {
        const float* H2p = sm + c.off_h[0][1];
        const float* H2v = sm + c.off_h[1][1];
        const int n_w = kH * (D + 1), n_all = n_w + 2 * D + 1;
        for (int o = tid; o < n_all; o += kThreads) {
          float s = 0.f;
          if (o < n_w) {
            const int k = o & (kH - 1), col = o >> 6;
            if (col < D) {
              for (int r = 0; r < kRows; ++r) s = fmaf(s_dloc[r * DP + col], H2p[r * kHP + k], s);
              G[oW3p + col * kHP + k] += s;
            } else {
              for (int r = 0; r < kRows; ++r) s = fmaf(s_dval[r], H2v[r * kHP + k], s);
              G[oW3v + k] += s;
            }
          } else if (o < n_w + D) {
            const int col = o - n_w;
            for (int r = 0; r < kRows; ++r) s += s_dloc[r * DP + col];
            G[ob3p + col] += s;
          } else if (o < n_w + 2 * D) {
            const int col = o - n_w - D;   // d loss / d logstd = sum_rows dscale * std
            for (int r = 0; r < kRows; ++r) s += s_dscale[r * DP + col];
            G[oLs + col] += s * s_sd[col];
          } else {
            for (int r = 0; r < kRows; ++r) s += s_dval[r];
            G[ob3v] += s;
          }
        }
      }
      __syncthreads();
      // ---- dZ2 = (dOut W3) * (1 - H2^2), in place over H2
      for (int e = t; e < kRows * kH; e += kNet) {
        const int r = e >> 6, k = e & (kH - 1);
        float dh = 0.f;
        if (net == 0) {
          for (int col = 0; col < D; ++col) dh = fmaf(s_dloc[r * DP + col], W[oW3p + col * kHP + k], dh);
        } else {
          dh = s_dval[r] * W[oW3v + k];
        }
        const float h = H2[r * kHP + k];
        H2[r * kHP + k] = dh * (1.f - h * h);
      }
      __syncthreads();
      // ---- layer 2: dW2 += dZ2^T H1, db2, dZ1 = (dZ2 W2) * (1 - H1^2) -> Sb
      gemm_nn(H2, kHP, H1, kHP, 0, kRows, t >> 4, t & 15, [&](int m, int n, float4 v) {
        float4* g = reinterpret_cast<float4*>(G + oW2 + m * kHP + n);
        float4 cur = *g;
        cur.x += v.x; cur.y += v.y; cur.z += v.z; cur.w += v.w;
        *g = cur;
      });
      gemm_kn(H2, kHP, W + oW2, kHP, kH, t, [&](int m, int n, float4 v) {
        const float4 h = ld4(H1 + m * kHP + n);
        *reinterpret_cast<float4*>(Sb + m * kHP + n) =
            make_float4(v.x * (1.f - h.x * h.x), v.y * (1.f - h.y * h.y), v.z * (1.f - h.z * h.z),
                        v.w * (1.f - h.w * h.w));
      });
      if (t < kH) {
        float s = 0.f;
        for (int r = 0; r < kRows; ++r) s += H2[r * kHP + t];
        G[ob2 + t] += s;
      }
      __syncthreads();
      // ---- layer 1: dW1 += dZ1^T X (only OP / 4 column groups exist: 16 x OP/4 tiles, the 64
      // rows split over as many thread groups as fit, shared-memory atomics), db1
      {
        const int ntn = OP >> 2, tiles = 16 * ntn;
        const int groups = kNet / tiles > 0 ? kNet / tiles : 1;
        const int grp = t / tiles, id = t - grp * tiles;
        if (grp < groups) {
          const int per = (kRows + groups - 1) / groups;
          const int r0 = grp * per, r1 = r0 + per < kRows ? r0 + per : kRows;
          gemm_nn(Sb, kHP, X, OP, r0, r1, id / ntn, id % ntn, [&](int m, int n, float4 v) {
            float* g = G + oW1 + m * OP + n;
            atomicAdd(g, v.x);
            atomicAdd(g + 1, v.y);
            atomicAdd(g + 2, v.z);
            atomicAdd(g + 3, v.w);
          });
        }
      }
      if (t < kH) {
        float s = 0.f;
        for (int r = 0; r < kRows; ++r) s += Sb[r * kHP + t];
        G[ob1 + t] += s;
      }
      __syncthreads();
    }

    // ---- the minibatch loss and its logged scalars
    block_sum<kAcc>(acc, red);
    if (tid == 0) {
      write_loss(acc, ks, true, true, a.losses + step, a.stats + (size_t)step * DERL_LOSS_STATS);
    }
    // ---- clip_grad_norm_ (torch.nn.utils.clip_grad_norm_: coef = max_norm / (norm + 1e-6),
    // clamped to 1) and Adam (torch.optim.Adam, single-tensor formulas)
    double sq[1] = {0.0};
    for (int e = tid; e < c.params; e += kThreads) {   // padding entries of G are zero
      const double g = (double)G[e];
      sq[0] += g * g;
    }
    __syncthreads();   // red is reused
    block_sum<1>(sq, red);
    if (tid == 0) {
      float coef = 1.f;
      if (a.max_norm >= 0.0) {
        const float norm = (float)sqrt(sq[0]);
        coef = fminf((float)a.max_norm / (norm + 1e-6f), 1.f);
      }
      pow1 *= a.beta1;   // beta^(step0 + step + 1), carried from step to step (thread 0)
      pow2 *= a.beta2;
      const double bc1 = 1.0 - pow1, bc2 = 1.0 - pow2;
      misc[2] = coef;
      misc[3] = (float)(a.lr / bc1);       // step_size
      misc[4] = (float)sqrt(bc2);          // bias_correction2_sqrt
      a.stats[(size_t)step * DERL_LOSS_STATS + 10] = (float)sqrt(sq[0]);   // grad norm before clip
    }
    __syncthreads();
    {
      const float coef = misc[2], step_size = misc[3], bc2s = misc[4];
      const float w1 = (float)(1.0 - a.beta1), b2 = (float)a.beta2, w2 = (float)(1.0 - a.beta2);
      const float eps = (float)a.eps;
      const bool apply_clip = a.max_norm >= 0.0;
      // over the padded block: padding entries have g = m = v = 0 and stay 0
      float* __restrict__ wm = a.ws_m;
      float* __restrict__ wv = a.ws_v;
#pragma unroll 4
      for (int e = tid; e < c.params; e += kThreads) {
        float g = G[e];
        G[e] = 0.f;                                   // ready for the next step's accumulation
        if (apply_clip) g = g * coef;
        float m = wm[e], v = wv[e];
        m = m + w1 * (g - m);                         // exp_avg.lerp_(grad, 1 - beta1)
        v = v * b2 + w2 * (g * g);                    // mul_(beta2).addcmul_(grad, grad, 1 - beta2)
        const float denom = sqrtf(v) / bc2s + eps;
        W[e] = W[e] - step_size * (m / denom);        // addcdiv_(exp_avg, denom, -step_size)
        wm[e] = m;
        wv[e] = v;
      }
    }
    __syncthreads();
  }

  // ---- parameters: shared -> global
  for (int k = 0; k < kTensors; ++k) {
    const int n = c.t_rows[k] * c.t_cols[k];
    for (int e = tid; e < n; e += kThreads) {
      const int r = e / c.t_cols[k], col = e - r * c.t_cols[k];
      const int so = c.t_off[k] + r * c.t_pitch[k] + col;
      a.t[k].w[e] = W[so];
      a.t[k].m[e] = a.ws_m[so];   // written by another thread of this CTA in the Adam loop: the
      a.t[k].v[e] = a.ws_v[so];   // __syncthreads() that ends the last step orders the two
    }
  }
}

}  // namespace
}  // namespace derl

using namespace derl;

extern "C" size_t derl_b200_ppo_mlp_update_workspace_bytes(int obs_dim, int act_dim) {
  if (obs_dim < 1 || obs_dim > 64 || act_dim < 1 || act_dim > kMaxAct) return 0;
  return 2 * (size_t)Carve(obs_dim, act_dim).params * sizeof(float);
}

extern "C" size_t derl_b200_ppo_mlp_update_smem_bytes(int obs_dim, int act_dim) {
  if (obs_dim < 1 || obs_dim > 64 || act_dim < 1 || act_dim > kMaxAct) return 0;
  const size_t bytes = Carve(obs_dim, act_dim).bytes();
  return bytes <= 227 * 1024 ? bytes : 0;
}

extern "C" int derl_b200_ppo_mlp_update(
    float* const* params_dev, float* const* exp_avg_dev, float* const* exp_avg_sq_dev, int obs_dim,
    int act_dim, const void* observations_dev, int obs_f64, const float* actions_dev,
    const float* old_logp_dev, const float* advantages_dev, const float* value_targets_dev,
    const float* old_values_dev, int64_t nsamples, const int64_t* perm_dev, int64_t nepochs,
    int64_t minibatch, int normalize_advantages, double adv_epsilon, int has_clip, double cliprange,
    double value_loss_coef, double entropy_coef, double max_grad_norm, double lr, double beta1,
    double beta2, double adam_eps, int64_t adam_step, float* losses_dev, float* stats_dev,
    void* workspace_dev, size_t workspace_bytes, void* stream) {
  DERL_REQUIRE(params_dev && exp_avg_dev && exp_avg_sq_dev, "ppo_mlp_update: null tensor tables");
  DERL_REQUIRE(observations_dev && actions_dev && old_logp_dev && advantages_dev &&
                   value_targets_dev && perm_dev && losses_dev && stats_dev,
               "ppo_mlp_update: null pointer");
  DERL_REQUIRE(!has_clip || old_values_dev, "ppo_mlp_update: clipped value loss needs old values");
  DERL_REQUIRE(nsamples >= 1 && nepochs >= 1 && minibatch >= 1 && minibatch <= nsamples,
               "ppo_mlp_update: need 1 <= minibatch <= nsamples and nepochs >= 1 (got %lld, %lld, "
               "%lld)", (long long)minibatch, (long long)nsamples, (long long)nepochs);
  DERL_REQUIRE(adam_step >= 0 && lr >= 0.0 && beta1 >= 0.0 && beta1 < 1.0 && beta2 >= 0.0 &&
                   beta2 < 1.0 && adam_eps >= 0.0, "ppo_mlp_update: bad Adam hyper-parameters");
  const size_t smem = derl_b200_ppo_mlp_update_smem_bytes(obs_dim, act_dim);
  DERL_REQUIRE(smem != 0, "ppo_mlp_update: obs_dim %d / act_dim %d do not fit the shared-memory "
               "plan (two 64-64 tanh MLPs, obs_dim <= ~40)", obs_dim, act_dim);
  if (workspace_dev == nullptr ||
      workspace_bytes < derl_b200_ppo_mlp_update_workspace_bytes(obs_dim, act_dim)) {
    set_error("ppo_mlp_update: workspace %zu B too small", workspace_bytes);
    return DERL_E_WORKSPACE;
  }
  int rc = require_device();
  if (rc != DERL_OK) return rc;
  const long long nmb = (nsamples + minibatch - 1) / minibatch;
  DERL_REQUIRE(nmb * nepochs < (1ll << 30), "ppo_mlp_update: too many optimiser steps");
  MlpArgs a;
  const Carve c(obs_dim, act_dim);
  for (int k = 0; k < kTensors; ++k) {
    DERL_REQUIRE(params_dev[k] && exp_avg_dev[k] && exp_avg_sq_dev[k],
                 "ppo_mlp_update: tensor %d is NULL", k);
    a.t[k] = MlpTensor{params_dev[k], exp_avg_dev[k], exp_avg_sq_dev[k], c.t_rows[k], c.t_cols[k],
                       c.t_off[k], c.t_pitch[k]};
  }
  a.obs = observations_dev;
  a.obs_f64 = obs_f64 ? 1 : 0;
  a.O = obs_dim;
  a.D = act_dim;
  a.actions = actions_dev;
  a.old_logp = old_logp_dev;
  a.adv = advantages_dev;
  a.vtarg = value_targets_dev;
  a.vold = old_values_dev;
  a.perm = reinterpret_cast<const long long*>(perm_dev);
  a.S = nsamples;
  a.mb = minibatch;
  a.step0 = adam_step;
  a.nsteps = (int)(nmb * nepochs);
  a.nmb = (int)nmb;
  a.normalize = normalize_advantages ? 1 : 0;
  a.has_clip = has_clip ? 1 : 0;
  a.adv_eps = adv_epsilon;
  a.clip = cliprange;
  a.vcoef = value_loss_coef;
  a.ecoef = entropy_coef;
  a.max_norm = max_grad_norm;
  a.lr = lr;
  a.beta1 = beta1;
  a.beta2 = beta2;
  a.eps = adam_eps;
  a.losses = losses_dev;
  a.stats = stats_dev;
  a.ws_m = reinterpret_cast<float*>(workspace_dev);
  a.ws_v = a.ws_m + c.params;
  if (int rc_attr = ensure_dynamic_smem(reinterpret_cast<const void*>(ppo_mlp_update_kernel),
                                        (int)smem))
    return rc_attr;
  ppo_mlp_update_kernel<<<1, kThreads, smem, as_stream(stream)>>>(a);
  DERL_LAUNCH_CHECK("ppo_mlp_update_kernel");
  return DERL_OK;
}
