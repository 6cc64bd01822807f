"""ORACLE — TEST INFRASTRUCTURE ONLY (CPU restatement of the reference's PPO data path).

Nothing under derl_b200/ may import this module.  It is used by tests/ (as the checker),
by __graft_entry__.smoke() (as the checker) and by bench.py's `cpu_baseline` /
`--impl reference` legs (as the reported CPU baseline, kind "port").

Each function restates one piece of mknbv/derl in NumPy / CPU torch, following the
reference's arithmetic step by step (dtype promotion, evaluation order, reductions) so the
results are what the reference produces; citations are relative to the reference root.
The reference itself is pure Python and cannot travel to the GPU box, hence this port.

Parity is PINNED (tests/test_oracle.py):
  * GAE against the reference fixture testdata/a2c/atari/interactions.npz (bit-exact) —
    repacked as tests/golden/ref_a2c_atari_gae.npz;
  * PPOLoss forward/backward against testdata/ppo/pybullet/{interactions,grads}.npz and
    losses.npy[0] (rtol = atol = 1e-5, the reference's own tolerance, ppo_test.py:50,53) —
    repacked as tests/golden/ref_ppo_pybullet.npz;
  * everything else against outputs of the reference's own classes run in the dev
    container on seeded inputs (tests/golden/live_*.npz from tests/golden/make_golden.py),
    and live against /root/reference when that tree is present.
"""
import ctypes
import os

import numpy as np
import torch

_HERE = os.path.dirname(os.path.abspath(__file__))


# ----------------------------------------------------------------------------- GAE (NumPy)
def gae(rewards, values, resets, last_value, gamma=0.99, lambda_=0.95, normalize=False,
        epsilon=1e-8):
  """GAE.__call__ arithmetic, derl/runners/trajectory_transforms.py:42-68.

  rewards [T,...] f32|f64, values [T,...] or [T,...,1] f32, resets [T,...] bool,
  last_value broadcastable to rewards[-1] (or with a trailing 1).  Returns
  (advantages f32 shaped like squeezed values, value_targets shaped like `values`).
  """
  rewards, resets = np.asarray(rewards), np.asarray(resets)
  values_in = np.asarray(values)
  values = values_in
  if values.ndim == rewards.ndim + 1:                       # :42-43
    values = np.squeeze(values, -1)
  last_value = np.asarray(last_value)
  if np.asarray(resets[-1]).ndim < last_value.ndim:         # :51-52
    last_value = np.squeeze(last_value, -1)
  nsteps = values.shape[0]
  adv = np.zeros_like(values, dtype=np.float32)             # :45
  adv[-1] = rewards[-1] - values[-1]                        # :46
  adv[-1] += (1 - resets[-1]) * gamma * last_value          # :53
  for t in range(nsteps - 2, -1, -1):                       # :56-62
    keep = 1 - resets[t]
    delta = rewards[t] + keep * gamma * values[t + 1] - values[t]
    adv[t] = delta + keep * gamma * lambda_ * adv[t + 1]
  targets = adv + values                                    # :63
  targets = targets.reshape(targets.shape + (1,) * (values_in.ndim - targets.ndim))
  if normalize or (normalize is None and adv.size > 1):     # :67-68
    adv = (adv - adv.mean()) / (adv.std() + epsilon)
  return adv, targets


def normalize_advantages(advantages, epsilon=1e-8):
  """NormalizeAdvantages.__call__, trajectory_transforms.py:89-92 (population std)."""
  advantages = np.asarray(advantages)
  return (advantages - advantages.mean()) / (advantages.std() + epsilon)


def merge_time_batch(columns):
  """MergeTimeBatch.__call__, trajectory_transforms.py:77-81: (T,N,...) -> (T*N,...)."""
  assert columns["resets"].ndim == 2, columns["resets"].shape
  return {k: (v.reshape((-1,) + v.shape[2:]) if isinstance(v, np.ndarray) else v)
          for k, v in columns.items()}


# ----------------------------------------------------------------------------- C oracle
_LIB = None


def c_lib():
  """liboracle.so (oracle/oracle.c), built on demand with `make -C oracle`."""
  global _LIB
  if _LIB is None:
    path = os.path.join(_HERE, "liboracle.so")
    src = os.path.join(_HERE, "oracle.c")
    if not os.path.exists(path) or os.path.getmtime(path) < os.path.getmtime(src):
      import subprocess
      subprocess.run(["make", "-C", _HERE], check=True, capture_output=True)
    lib = ctypes.CDLL(path)
    i64, f64, ptr = ctypes.c_int64, ctypes.c_double, ctypes.c_void_p
    lib.derl_oracle_gae.argtypes = [ptr, ctypes.c_int, ptr, ptr, ptr, i64, i64, f64, f64, ptr, ptr]
    lib.derl_oracle_gae.restype = None
    lib.derl_oracle_normalize.argtypes = [ptr, i64, f64, ptr, ptr]
    lib.derl_oracle_normalize.restype = None
    lib.derl_oracle_gather_rows.argtypes = [ptr, i64, ptr, i64, i64, ptr]
    lib.derl_oracle_gather_rows.restype = None
    _LIB = lib
  return _LIB


def _p(arr):
  return arr.ctypes.data_as(ctypes.c_void_p)


def gae_c(rewards, values, resets, last_value, gamma=0.99, lambda_=0.95):
  """Scalar-order C statement of the same recursion on [T, N] arrays (fast at full sizes)."""
  rewards = np.ascontiguousarray(rewards)
  if rewards.dtype not in (np.float32, np.float64):
    rewards = rewards.astype(np.float64)
  values = np.ascontiguousarray(values, dtype=np.float32)
  nsteps = values.shape[0]
  nenvs = values.size // nsteps
  resets = np.ascontiguousarray(resets).astype(np.uint8)
  last_value = np.ascontiguousarray(last_value, dtype=np.float32).reshape(-1)
  assert last_value.size == nenvs and rewards.size == values.size == resets.size
  adv = np.empty((nsteps, nenvs), np.float32)
  targets = np.empty((nsteps, nenvs), np.float32)
  c_lib().derl_oracle_gae(_p(rewards), int(rewards.dtype == np.float64), _p(values), _p(resets),
                          _p(last_value), nsteps, nenvs, float(gamma), float(lambda_), _p(adv),
                          _p(targets))
  return adv, targets


def normalize_c(x, epsilon=1e-8):
  x = np.ascontiguousarray(x, dtype=np.float32).reshape(-1)
  out, scratch = np.empty_like(x), np.empty_like(x)
  c_lib().derl_oracle_normalize(_p(x), x.size, float(epsilon), _p(out), _p(scratch))
  return out


def gather_rows_c(src, perm, start, count):
  src = np.ascontiguousarray(src)
  perm = np.ascontiguousarray(perm, dtype=np.int64)
  row_bytes = src.strides[0] if src.ndim > 1 else src.itemsize
  dst = np.empty((count,) + src.shape[1:], src.dtype)
  c_lib().derl_oracle_gather_rows(_p(src), row_bytes, _p(perm), start, count, _p(dst))
  return dst


# ----------------------------------------------------------------------------- minibatches
def iterate_minibatches(columns, num_epochs=3, num_minibatches=4, shuffle_before_epoch=True):
  """IterateWithMinibatches.run for one rollout, derl/runners/onpolicy.py:43-62.

  One np.random.permutation(S) per epoch from the GLOBAL NumPy RNG (:46), every key except
  "state" re-indexed IN PLACE (:47-49, so permutations compose across epochs), then
  contiguous slices of S // num_minibatches rows (:56-62; a short tail batch if S is not
  divisible).  Yields dicts of NumPy arrays; mutates `columns` like the reference does.
  """
  for _ in range(num_epochs):
    if shuffle_before_epoch:
      size = columns["observations"].shape[0]
      order = np.random.permutation(size)
      for key in [k for k in columns if k != "state"]:
        columns[key] = columns[key][order]
    size = columns["observations"].shape[0]
    step = size // num_minibatches
    for start in range(0, size, step):
      rows = np.arange(start, min(start + step, size))
      yield {k: (v if k == "state" else v[rows]) for k, v in columns.items()}


def composed_permutations(size, num_epochs):
  """The row order each epoch sees, as indices into the ORIGINAL rollout:
  P_0 = perm_0, P_e = P_{e-1}[perm_e] (consequence of onpolicy.py:47-49 being in place)."""
  order = np.arange(size)
  out = []
  for _ in range(num_epochs):
    order = order[np.random.permutation(size)]
    out.append(order.copy())
  return out


# ----------------------------------------------------------------------------- PPO loss
def make_distribution(*dist_inputs):
  """ActorCriticPolicy.act's distribution choice, derl/policies.py:61-66."""
  if len(dist_inputs) == 1:
    return torch.distributions.Categorical(logits=dist_inputs[0])
  if len(dist_inputs) == 2:
    return torch.distributions.Independent(torch.distributions.Normal(*dist_inputs), 1)
  raise ValueError("expected one (categorical) or two (normal) distribution inputs")


def ppo_loss(dist_inputs, values, batch, cliprange=0.2, value_loss_coef=0.25,
             entropy_coef=0.01, return_parts=False):
  """PPOLoss.__call__ on CPU torch, derl/alg/ppo.py:31-64 (policy), :73-98 (value), :104.

  dist_inputs: (logits,) or (loc, scale) torch tensors (may require grad); values [B,1] or
  [B]; batch: dict with actions, log_prob, advantages, value_targets, values (NumPy or
  torch).  Uses torch.distributions exactly like the reference (third-party arithmetic).
  """
  def t(x):
    return x if isinstance(x, torch.Tensor) else torch.from_numpy(np.asarray(x))
  dist = make_distribution(*dist_inputs)
  old_log_prob, adv, actions = t(batch["log_prob"]), t(batch["advantages"]), t(batch["actions"])
  log_prob = dist.log_prob(actions)                                  # :35
  ratio = torch.exp(log_prob - old_log_prob)                         # :45
  surrogate = -ratio * adv                                           # :46
  if cliprange is not None:                                          # :47-51
    clipped = -torch.clamp(ratio, 1. - cliprange, 1. + cliprange) * adv
    surrogate = torch.max(surrogate, clipped)
  policy_loss = torch.mean(surrogate)                                # :53
  entropy = torch.mean(dist.entropy())                               # :54
  targets, old_values = t(batch["value_targets"]), t(batch["values"])
  value_loss = torch.pow(values - targets, 2)                        # :82
  if cliprange is not None:                                          # :83-87
    values_clipped = old_values + torch.clamp(values - old_values, -cliprange, cliprange)
    value_loss = torch.max(value_loss, torch.pow(values_clipped - targets, 2))
  value_loss = torch.mean(value_loss)                                # :89
  loss = (policy_loss - entropy_coef * entropy) + value_loss_coef * value_loss   # :64, :104
  if return_parts:
    pred_var = torch.pow(values.std(), 2)                            # derl/alg/common.py:9-12
    parts = dict(loss=loss, policy_loss=policy_loss, entropy=entropy, value_loss=value_loss,
                 advantages=torch.mean(adv), value_targets=torch.mean(targets),
                 value_preds=torch.mean(values),
                 r_squared=1. - torch.mean(torch.pow(values - targets, 2)) / pred_var)
    return loss, parts
  return loss


def a2c_loss(dist_inputs, values, batch, value_loss_coef=0.25, entropy_coef=0.01,
             return_parts=False):
  """A2CLoss.__call__ on CPU torch, derl/alg/a2c.py:19-79."""
  def t(x):
    return x if isinstance(x, torch.Tensor) else torch.from_numpy(np.asarray(x))
  dist = make_distribution(*dist_inputs)
  adv, targets = t(batch["advantages"]), t(batch["value_targets"])
  log_prob = dist.log_prob(t(batch["actions"]))                       # :23-24
  policy_loss = -torch.mean(log_prob * adv)                           # :32
  entropy = torch.mean(dist.entropy())                                # :33
  value_loss = torch.mean(torch.pow(values - targets, 2))             # :57
  loss = (policy_loss - entropy_coef * entropy) + value_loss_coef * value_loss   # :44, :72
  if return_parts:
    target_var = torch.pow(targets.std(), 2)   # r_squared(values, value_targets): args as in :62
    return loss, dict(loss=loss, policy_loss=policy_loss, entropy=entropy, value_loss=value_loss,
                      advantages=torch.mean(adv), value_targets=torch.mean(targets),
                      value_preds=torch.mean(values),
                      r_squared=1. - torch.mean(torch.pow(targets - values, 2)) / target_var)
  return loss


def ppo_loss_closed_form(dist_inputs, values, batch, cliprange=0.2, value_loss_coef=0.25,
                         entropy_coef=0.01):
  """Second, autograd-free statement (float64 NumPy) of the loss and of its gradients with
  respect to the distribution inputs and values — SURVEY.md §8a closed form, including
  torch.maximum's 1/2:1/2 tie split and clamp's closed pass-through interval.  Returns
  (loss, grads) with grads a tuple aligned with (*dist_inputs, values)."""
  f = lambda x: np.asarray(x.detach() if isinstance(x, torch.Tensor) else x, dtype=np.float64)
  adv, old_lp = f(batch["advantages"]), f(batch["log_prob"])
  vt = f(batch["value_targets"]).reshape(-1)
  v_old = f(batch["values"]).reshape(-1)
  v = f(values).reshape(-1)
  nb = adv.shape[0]
  if len(dist_inputs) == 1:
    z = f(dist_inputs[0])
    acts = np.asarray(batch["actions"]).astype(np.int64)
    logp = z - (np.log(np.exp(z - z.max(1, keepdims=True)).sum(1, keepdims=True))
                + z.max(1, keepdims=True))
    p = np.exp(logp)
    lp = logp[np.arange(nb), acts]
    ent = -(p * logp).sum(1)
  else:
    mu, sd = f(dist_inputs[0]), f(dist_inputs[1])
    acts = f(batch["actions"])
    lp = (-(acts - mu) ** 2 / (2 * sd ** 2) - np.log(sd) - 0.5 * np.log(2 * np.pi)).sum(1)
    ent = (0.5 + 0.5 * np.log(2 * np.pi) + np.log(sd)).sum(1)
  ratio = np.exp(lp - old_lp)
  s1 = -ratio * adv
  if cliprange is None:
    pol, w = s1, np.ones(nb)
  else:
    lo, hi = 1. - cliprange, 1. + cliprange
    s2 = -np.clip(ratio, lo, hi) * adv
    inside = (ratio >= lo) & (ratio <= hi)
    pol = np.maximum(s1, s2)
    w = np.where(s1 > s2, 1., np.where(s1 == s2, np.where(inside, 1., .5),
                                       np.where(inside, 1., 0.)))
  g = -adv * ratio * w / nb
  u = v - vt
  if cliprange is None:
    vl, dv = u ** 2, 2 * u
  else:
    d = v - v_old
    w2 = v_old + np.clip(d, -cliprange, cliprange) - vt
    passes = ((d >= -cliprange) & (d <= cliprange)).astype(np.float64)
    vl = np.maximum(u ** 2, w2 ** 2)
    dv = np.where(u ** 2 > w2 ** 2, 2 * u,
                  np.where(w2 ** 2 > u ** 2, 2 * w2 * passes, u + w2 * passes))
  loss = pol.mean() - entropy_coef * ent.mean() + value_loss_coef * vl.mean()
  dvalues = (value_loss_coef / nb * dv).reshape(np.shape(f(values)))
  if len(dist_inputs) == 1:
    onehot = np.zeros_like(p)
    onehot[np.arange(nb), acts] = 1.
    dz = g[:, None] * (onehot - p) + (entropy_coef / nb) * p * (logp + ent[:, None])
    return loss, (dz, dvalues)
  dmu = g[:, None] * (acts - mu) / sd ** 2
  dsd = g[:, None] * ((acts - mu) ** 2 - sd ** 2) / sd ** 3 - (entropy_coef / nb) / sd
  return loss, (dmu, dsd, dvalues)


# ----------------------------------------------------------------------------- models (CPU)
def _orthogonal(module):
  """orthogonal_init, derl/models.py:127-138: orthogonal weights, zero biases."""
  if hasattr(module, "weight"):
    torch.nn.init.orthogonal_(module.weight)
  if hasattr(module, "bias"):
    torch.nn.init.zeros_(module.bias)


class NatureCNN(torch.nn.Module):
  """NatureCNNModel(output_units=[A, 1]) forward, derl/models.py:102-124,198-214:
  NHWC uint8 -> NCHW float/255 -> conv 8x8/4, 4x4/2, 3x3/1 (ReLU each) -> fc 512 (no ReLU)
  -> linear heads."""

  def __init__(self, nactions):
    super().__init__()
    nn = torch.nn
    self.base = nn.Sequential(nn.Conv2d(4, 32, 8, 4), nn.ReLU(), nn.Conv2d(32, 64, 4, 2),
                              nn.ReLU(), nn.Conv2d(64, 64, 3, 1), nn.ReLU(), nn.Flatten(),
                              nn.Linear(3136, 512))
    self.heads = nn.ModuleList([nn.Linear(512, nactions), nn.Linear(512, 1)])
    self.apply(_orthogonal)

  def forward(self, obs):
    x = obs.permute(0, 3, 1, 2)
    if x.dtype == torch.uint8:
      x = x.float() / 255
    hidden = self.base(x.contiguous())
    return tuple(head(hidden) for head in self.heads)


class MuJoCoMLP(torch.nn.Module):
  """MuJoCoModel(obs_dim, [D, 1]) forward, derl/models.py:224-271: two tanh MLPs
  obs->64->64->{D,1}, state-independent logstd parameter, std = exp(logstd) per row."""

  def __init__(self, obs_dim, act_dim):
    super().__init__()
    nn = torch.nn
    mlp = lambda out: nn.Sequential(nn.Linear(obs_dim, 64), nn.Tanh(), nn.Linear(64, 64),
                                    nn.Tanh(), nn.Linear(64, out))
    self.nets = nn.ModuleList([mlp(act_dim), mlp(1)])
    self.apply(_orthogonal)
    self.logstd = nn.Parameter(torch.zeros(act_dim))

  def forward(self, obs):
    obs = obs.to(self.logstd.dtype)
    loc, values = (net(obs) for net in self.nets)
    std = torch.repeat_interleave(torch.exp(self.logstd)[None], obs.shape[0], 0)
    return loc, std, values


# ----------------------------------------------------------------------------- full update
def ppo_update(model, optimizer, rollout, last_value, *, gamma=0.99, lambda_=0.95,
               num_epochs=3, num_minibatches=4, cliprange=0.2, value_loss_coef=0.25,
               entropy_coef=0.01, max_grad_norm=0.5, batched=True):
  """One rollout through the reference's PPO pipeline on the host (the CPU baseline):
  GAE(normalize=False) -> MergeTimeBatch -> epochs x minibatches (shuffle, slice) ->
  NormalizeAdvantages -> PPOLoss -> backward -> clip_grad_norm_ -> optimizer.step
  (derl/runners/onpolicy.py:65-75, derl/alg/common.py:66-78).  `rollout` is the dict
  EnvRunner yields after np.asarray (env_runner.py:45-67).  Returns the list of losses."""
  columns = dict(rollout)
  adv, targets = gae(columns["rewards"], columns["values"], columns["resets"], last_value,
                     gamma, lambda_, normalize=False)
  columns["advantages"], columns["value_targets"] = adv, targets
  if batched:
    columns = merge_time_batch(columns)
  losses = []
  for batch in iterate_minibatches(columns, num_epochs, num_minibatches):
    batch["advantages"] = normalize_advantages(batch["advantages"])
    *dist_inputs, values = model(torch.from_numpy(batch["observations"]))
    loss = ppo_loss(dist_inputs, values, batch, cliprange, value_loss_coef, entropy_coef)
    optimizer.zero_grad()
    loss.backward()
    if max_grad_norm is not None:
      torch.nn.utils.clip_grad_norm_(model.parameters(), max_grad_norm)
    optimizer.step()
    losses.append(float(loss.detach()))
  return losses
