"""Annealing variables (reference: derl/anneal.py — AnnealingVariable :14-43,
LinearAnneal :65-86).  Same API and values; `LinearAnneal.step_to(n)` is closed-form instead
of the reference's one-tensor-per-env-step Python loop (:32-38), which costs ~13 us per env
step (~7 s per 4096x128 rollout) and would dominate the update once the
kernels are fast.  The value lives in a 0-d tensor shared with the optimizer (:76-81 of
derl/factory/ppo.py) and is updated in place, so CUDA-graph captures stay valid.
"""
from abc import ABC, abstractmethod
import re

import torch

from . import summary


def camel2snake(string):
  sub = re.sub("(.)([A-Z][a-z]+)", r"\1_\2", string)
  return re.sub("([a-z0-9])([A-Z])", r"\1_\2", sub).lower()


class AnnealingVariable(ABC):
  """Variable the value of which changes after each step."""

  def __init__(self, name=None):
    self.name = name or camel2snake(self.__class__.__name__)
    self.step_count = 0

  @abstractmethod
  def get_tensor(self):
    """Tensor that changes after each call to step."""

  def get_current_value(self):
    return self.get_tensor().clone()

  @abstractmethod
  def step(self):
    """Update the value of the variable."""

  def step_to(self, val):
    if val < self.step_count:
      raise ValueError(f"val={val} cannot be smaller than "
                       f"self.step_count={self.step_count}")
    for _ in range(val - self.step_count):
      self.step()

  def summarize(self, global_step):
    summary.add_scalar(f"anneal/{self.name}", self.get_tensor(), global_step=global_step)


class LinearAnneal(AnnealingVariable):
  """start -> end linearly over nsteps, clamped (reference :65-86)."""

  def __init__(self, start, nsteps, end=0., name=None, device=None):
    super().__init__(name)
    self.start, self.nsteps, self.end = start, nsteps, end
    self.tensor = torch.tensor(self.start, device=device)

  def get_tensor(self):
    return self.tensor

  def _value_at(self, count):
    value = self.start + (self.end - self.start) * (count / self.nsteps)
    return min(max(value, min(self.start, self.end)), max(self.start, self.end))

  def _set(self, count):
    self.step_count = count
    self.tensor.fill_(self._value_at(count))

  def step(self):
    self._set(self.step_count + 1)
    return self.get_current_value()

  def step_to(self, val):
    if val < self.step_count:
      raise ValueError(f"val={val} cannot be smaller than "
                       f"self.step_count={self.step_count}")
    if val != self.step_count:
      self._set(val)
