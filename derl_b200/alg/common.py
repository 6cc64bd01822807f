"""Loss / Trainer / Alg: the consumer side of the hot path (reference: derl/alg/common.py —
r_squared :9-12, total_norm :15-20, Loss :23-45, Trainer :48-78, Alg :81-106).

One addition: `Trainer(grad_sync=...)`, a callable run between `loss.backward()` and the
gradient clipping (reference :70 / :71) — the insertion point of the NCCL gradient
all-reduce when the rollout is sharded along the env axis (derl_b200.parallel).
"""
from abc import ABC, abstractmethod

import numpy as np
import torch

from .. import summary


def r_squared(targets, predictions):
  """Coefficient of determination with torch's Bessel-corrected std (reference :9-12)."""
  variance = torch.pow(predictions.std(), 2)
  return 1. - torch.mean(torch.pow(predictions - targets, 2)) / variance


def total_norm(tensors, norm_type=2):
  """Norm of the tensors as if concatenated into one vector."""
  if norm_type == float("inf"):
    return max(t.abs().max() for t in tensors)
  return sum(t.norm(norm_type) ** norm_type for t in tensors) ** (1. / norm_type)


class Loss(ABC):
  """Algorithm loss function bound to a model."""

  def __init__(self, model, name=None):
    self.model = model
    if name is None:
      name = self.__class__.__name__
      name = name[:-len("Loss")] if name.endswith("Loss") else name
      name = name.lower()
    self.name = name
    self.call_count = 0

  @property
  def device(self):
    return next(self.model.parameters()).device

  def torch_from_numpy(self, arr):
    """NumPy array (or tensor) -> tensor on the model's device."""
    if isinstance(arr, torch.Tensor):
      return arr if arr.device == self.device else arr.to(self.device)
    return torch.from_numpy(np.ascontiguousarray(arr)).to(device=self.device)

  @abstractmethod
  def __call__(self, data):
    """Computes and returns loss value on given data."""


def _split_rows(data, size):
  """Row chunks of a minibatch dict ("state" passes through), for gradient accumulation."""
  nrows = data["observations"].shape[0]
  for lo in range(0, nrows, size):
    yield {k: (v if k == "state" else v[lo:lo + size]) for k, v in data.items()}, \
        min(size, nrows - lo) / nrows


class Trainer:
  """loss -> zero_grad -> backward -> [grad_sync] -> clip -> anneal -> optimizer.step.

  Extensions over the reference (both default to the reference's behaviour):
    grad_sync    callable(model) run right after backward (NCCL gradient all-reduce);
    micro_batch  rows per forward/backward pass: a minibatch larger than this is processed
                 in row chunks whose losses are weighted by their share of the minibatch, so
                 the accumulated gradient equals the full-minibatch mean's gradient while
                 activation memory stays bounded (131072 x 84x84x4 frames do not fit one
                 cuDNN call).  The returned loss is the weighted sum (== the minibatch mean).
  """

  def __init__(self, optimizer, anneals=None, max_grad_norm=None, grad_sync=None,
               micro_batch=None):
    self.optimizer = optimizer
    self.anneals = anneals or []
    self.max_grad_norm = max_grad_norm
    self.grad_sync = grad_sync
    self.micro_batch = micro_batch
    self.step_count = 0

  def preprocess_gradients(self, parameters, tag):
    grad_norm = None
    parameters = list(parameters)
    if self.max_grad_norm is not None:
      grad_norm = torch.nn.utils.clip_grad_norm_(parameters, self.max_grad_norm)
    if summary.should_record():
      if grad_norm is None:
        grad_norm = total_norm(p.grad for p in parameters if p.grad is not None)
      summary.add_scalar(tag, grad_norm, global_step=self.step_count)

  def _backward(self, alg, data):
    nrows = data["observations"].shape[0] if "observations" in data else 0
    # a flat gradient buffer (grad_sync) must keep its views alive: zero in place then
    keep_buffers = self.grad_sync is not None
    arm = getattr(self.grad_sync, "arm", None)   # overlapped all-reduce: see derl_b200.parallel
    if not self.micro_batch or nrows <= self.micro_batch:
      loss = alg.loss(data)
      self.optimizer.zero_grad(set_to_none=not keep_buffers)
      if arm is not None:
        arm()
      loss.backward()
      return loss
    self.optimizer.zero_grad(set_to_none=not keep_buffers)
    total = None
    chunks = list(_split_rows(data, self.micro_batch))
    for i, (chunk, weight) in enumerate(chunks):
      part = alg.loss(chunk) * weight
      if arm is not None and i == len(chunks) - 1:
        arm()                                    # this backward completes the gradients
      part.backward()
      total = part.detach() if total is None else total + part.detach()
    return total

  def step(self, alg, data):
    loss = self._backward(alg, data)
    if self.grad_sync is not None:
      self.grad_sync(alg.model)
    self.preprocess_gradients(alg.model.parameters(), f"{alg.name}/grad_norm")
    for anneal in self.anneals:
      if summary.should_record():
        anneal.summarize(alg.runner.step_count)
      anneal.step_to(alg.runner.step_count)
    self.optimizer.step()
    self.step_count += 1
    return loss


class Alg:
  """Generic learning algorithm specified by its loss function."""

  def __init__(self, runner, trainer, loss_fn, name=None):
    self.runner = runner
    self.model = self.runner.policy.model
    self.trainer = trainer
    self.loss_fn = loss_fn
    self.name = name if name is not None else self.__class__.__name__.lower()

  def loss(self, data):
    return self.loss_fn(data)

  def step(self, data):
    return self.trainer.step(self, data)

  def learn(self, progress=True):
    """Performs learning until the runner is exhausted (reference :100-106).  `progress=False`
    (extension) drops the tqdm bar."""
    from tqdm import tqdm
    with tqdm(total=len(self.runner), disable=not progress) as pbar:
      for data in self.runner.run():
        pbar.update(self.runner.step_count - pbar.n)
        self.step(data)
