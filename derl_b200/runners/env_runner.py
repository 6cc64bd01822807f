"""Rollout producer and the wrapper base class of the drop-in boundary.

`RunnerWrapper` keeps the reference's proxying contract (derl/runners/env_runner.py:72-90):
only {env, policy, horizon, nsteps, step_count, nenvs, is_exhausted} are forwarded, `run`
is an abstract generator, `__len__` forwards, `unwrapped` is the innermost runner.
`EnvRunner` (env_runner.py:6-69) is host Python around CPU simulators — out of the
accelerated scope — and is provided so `make_ppo_runner` stays a drop-in.
"""
from abc import ABC, abstractmethod

import numpy as np
import torch

_FORWARDED = frozenset(("env", "policy", "horizon", "nsteps", "step_count", "nenvs",
                        "is_exhausted"))


class _ResidentRollout:
  """Preallocated [horizon, ...] tensors on `device`, filled one step at a time
  (SURVEY.md §8f rank 3): the rollout is born in HBM instead of being stacked from T host
  lists (`np.asarray`, derl/runners/onpolicy.py:20-27) and uploaded afterwards."""

  def __init__(self, horizon, device):
    self.horizon, self.device = horizon, torch.device(device)
    self.buffers = {}

  def put(self, key, step, value):
    arr = np.asarray(value)
    if arr.dtype == object or arr.dtype.kind in "USV":
      return False
    buf = self.buffers.get(key)
    if buf is None or buf.shape[1:] != arr.shape or buf.dtype != torch.from_numpy(arr[None]).dtype:
      buf = self.buffers[key] = torch.empty((self.horizon,) + arr.shape,
                                            dtype=torch.from_numpy(arr[None]).dtype,
                                            device=self.device)
    buf[step].copy_(torch.from_numpy(np.ascontiguousarray(arr)), non_blocking=True)
    return True


class EnvRunner:
  """Steps `env` with `policy` for `horizon` steps per yielded rollout.

  `resident_device` (extension, default None = reference behaviour): write every numeric
  per-step value straight into preallocated device tensors; `next_observations` is then not
  collected (nothing on the PPO path reads it) and `infos` stays a host list.
  """

  def __init__(self, env, policy, horizon, nsteps=None, time_limit=None, resident_device=None):
    self.resident_device = resident_device
    self.env, self.policy, self.horizon = env, policy, horizon
    self.nsteps = int(nsteps)
    if time_limit is not None and getattr(env.unwrapped, "nenvs", None) is not None:
      raise TypeError("batched envs are not supported for time_limit "
                      f"not equal to None, got env={env}, time_limit={time_limit}")
    self.time_limit = time_limit
    self.step_count = 0
    self.episode_length = 0

  @property
  def nenvs(self):
    return getattr(self.env.unwrapped, "nenvs", None)

  def is_exhausted(self):
    return self.nsteps is not None and self.step_count >= self.nsteps

  def __len__(self):
    return self.nsteps if self.nsteps is not None else self.step_count

  def run(self, obs=None):
    if obs is None:
      obs = self.env.reset()
      self.episode_length = 0
    while not self.is_exhausted():
      rollout = {}
      resident = None
      if self.resident_device is not None:
        resident = _ResidentRollout(self.horizon, self.resident_device)

      def put(key, val, step):
        if resident is not None and key == "next_observations":
          return
        if resident is not None and resident.put(key, step, val):
          rollout[key] = resident.buffers[key]
        else:
          rollout.setdefault(key, []).append(val)

      for step in range(self.horizon):
        act = self.policy.act(obs)
        put("observations", obs, step)
        if "actions" not in act:
          raise ValueError("result of policy.act must contain 'actions' "
                           f"but has keys {list(act.keys())}")
        for key, val in act.items():
          put(key, val, step)
        next_obs, reward, done, info = self.env.step(act["actions"])
        self.episode_length += 1
        put("rewards", reward, step)
        put("resets", done, step)
        put("infos", info, step)
        put("next_observations", next_obs, step)
        # batched envs auto-reset; a single env is reset here (env_runner.py:58-65)
        if self.nenvs is None and (done or self.episode_length == self.time_limit):
          obs = self.env.reset()
          self.episode_length = 0
        else:
          obs = next_obs
      rollout["state"] = dict(latest_observations=obs)
      self.step_count += self.horizon * (self.nenvs or 1)
      yield rollout


class RunnerWrapper(ABC):
  """Base of TransformInteractions / IterateWithMinibatches."""

  def __init__(self, runner):
    self.runner = runner
    self.unwrapped = getattr(runner, "unwrapped", runner)

  def __getattr__(self, attr):
    if attr not in _FORWARDED:
      raise AttributeError(f"'{self.__class__.__name__}' has no attribute '{attr}'")
    return getattr(self.runner, attr)

  def __len__(self):
    return len(self.runner)

  @abstractmethod
  def run(self, obs=None):
    """Generator of interaction dicts."""
