"""Seeded synthetic rollout source with the attribute surface of the reference's EnvRunner
(derl/runners/env_runner.py:6-69; SURVEY.md §8b "Rollout source"), for benches and tests:
no simulator is involved, the arrays have exactly the dtypes/shapes EnvRunner + np.asarray
produce (§8a a1, §8d):

  atari   observations uint8 (T,N,84,84,4), actions int64 (T,N), log_prob f32 (T,N),
          values f32 (T,N,1), rewards {-1,0,1} (T,N), resets bool (T,N),
          state.latest_observations uint8 (N,84,84,4)
  mujoco  unbatched (nenvs=None): observations f64 (T,17), actions f32 (T,6), log_prob f32
          (T,), values f32 (T,1), rewards f64 (T,), resets bool (T,),
          state.latest_observations f64 (17,)

`device="cuda"` builds the rollout directly in HBM (torch generators); `device="cpu"`
returns NumPy arrays like the reference's runner does.  No `next_observations` / `infos`.
"""
import math
import types

import numpy as np
import torch


def _space(**kw):
  return types.SimpleNamespace(**kw)


class SyntheticEnv:
  """Just enough of an env for the wrappers: spaces and `unwrapped.nenvs`."""

  def __init__(self, kind, nenvs, nactions, obs_dim, act_dim):
    self.nenvs = nenvs
    if kind == "atari":
      self.observation_space = _space(shape=(84, 84, 4), dtype=np.uint8)
      self.action_space = _space(n=nactions, shape=())
    else:
      self.observation_space = _space(shape=(obs_dim,), dtype=np.float64)
      self.action_space = _space(shape=(act_dim,), dtype=np.float32)

  @property
  def unwrapped(self):
    return self


def make_rollout(kind, horizon, nenvs, device="cuda", seed=0, nactions=4, obs_dim=17,
                 act_dim=6, rewards_dtype=None, reset_prob=None):
  """One rollout dict (see module docstring).  `nenvs=None` -> unbatched."""
  dev = torch.device(device)
  gen = torch.Generator(device=dev)
  gen.manual_seed(seed)
  lead = (horizon,) if nenvs is None else (horizon, nenvs)
  tail = () if nenvs is None else (nenvs,)
  randn = lambda shape, dtype=torch.float32: torch.randn(shape, generator=gen, device=dev,
                                                         dtype=dtype)
  rand = lambda shape: torch.rand(shape, generator=gen, device=dev)
  if kind == "atari":
    obs = torch.randint(0, 256, lead + (84, 84, 4), generator=gen, device=dev,
                        dtype=torch.uint8)
    latest = torch.randint(0, 256, tail + (84, 84, 4), generator=gen, device=dev,
                           dtype=torch.uint8)
    actions = torch.randint(0, nactions, lead, generator=gen, device=dev, dtype=torch.int64)
    log_prob = -math.log(nactions) + 0.01 * randn(lead)
    rdt = rewards_dtype or (torch.float32 if dev.type == "cuda" else torch.float64)
    rewards = (torch.sign(randn(lead)) * (rand(lead) < 0.1)).to(rdt)
    resets = rand(lead) < (0.01 if reset_prob is None else reset_prob)
  elif kind == "mujoco":
    obs = randn(lead + (obs_dim,), torch.float64)
    latest = randn(tail + (obs_dim,), torch.float64)
    actions = randn(lead + (act_dim,))
    log_prob = (-0.5 * actions ** 2 - 0.5 * math.log(2 * math.pi)).sum(-1) + 0.01 * randn(lead)
    rewards = randn(lead, rewards_dtype or torch.float64)
    resets = rand(lead) < (0.001 if reset_prob is None else reset_prob)
  else:
    raise ValueError(f"unknown rollout kind {kind!r}")
  values = randn(lead + (1,))
  rollout = dict(observations=obs, actions=actions, log_prob=log_prob, values=values,
                 rewards=rewards, resets=resets, state=dict(latest_observations=latest))
  if dev.type == "cpu":
    for key, val in rollout.items():
      if key != "state":
        rollout[key] = val.numpy()
    rollout["state"]["latest_observations"] = latest.numpy()
  return rollout


class SyntheticRolloutRunner:
  """Yields the same (or freshly seeded) synthetic rollout until `nsteps` env steps."""

  def __init__(self, policy, kind="atari", nenvs=8, horizon=128, nsteps=None, device="cuda",
               seed=0, fresh_each_rollout=False, **rollout_kwargs):
    self.policy, self.kind, self.horizon = policy, kind, horizon
    self.nsteps = None if nsteps is None else int(nsteps)
    self.device, self.seed, self.fresh = device, seed, fresh_each_rollout
    self.rollout_kwargs = rollout_kwargs
    self.env = SyntheticEnv(kind, nenvs, rollout_kwargs.get("nactions", 4),
                            rollout_kwargs.get("obs_dim", 17), rollout_kwargs.get("act_dim", 6))
    self.step_count = 0
    self._cached = None

  @property
  def nenvs(self):
    return self.env.nenvs

  def is_exhausted(self):
    return self.nsteps is not None and self.step_count >= self.nsteps

  def __len__(self):
    return self.nsteps if self.nsteps is not None else self.step_count

  def rollout(self):
    if self._cached is None or self.fresh:
      seed = self.seed + (self.step_count if self.fresh else 0)
      self._cached = make_rollout(self.kind, self.horizon, self.nenvs, self.device, seed,
                                  **self.rollout_kwargs)
    out = dict(self._cached)
    out["state"] = dict(self._cached["state"])
    return out

  def run(self, obs=None):
    while not self.is_exhausted():
      data = self.rollout()
      self.step_count += self.horizon * (self.nenvs or 1)
      yield data
