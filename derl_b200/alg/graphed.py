"""CUDA-graphed training step.

The reference's `Trainer.step` (derl/alg/common.py:66-78) issues, per minibatch, ~60 forward
and ~80 backward ATen ops plus clipping and Adam from Python; for the small default configs
(Atari 8 envs x 128 steps -> minibatches of 256; MuJoCo 2048 steps -> minibatches of 64) the
GPU finishes each kernel long before Python has launched the next one.  `GraphedTrainer`
captures the whole step — policy forward, fused PPO loss, backward, clip_grad_norm_,
optimizer step — once per minibatch signature into CUDA graphs and replays them: per
minibatch the host issues one multi-tensor copy into the static inputs and one or two graph
launches.  Same arithmetic, same kernels, same order as the eager `Trainer`.

Requirements: a capturable optimizer (`torch.optim.Adam(..., capturable=True)`; the learning
rate may be a device tensor, e.g. `LinearAnneal(..., device="cuda").get_tensor()`, which
`step_to` updates in place between replays).  With `grad_sync` (NCCL all-reduce) the step
is split into two graphs around the eager collective.
"""
import torch

from .. import summary
from .common import Trainer

# what PPOLoss reads from a minibatch (everything else is not copied into the graph inputs)
LOSS_KEYS = ("observations", "actions", "log_prob", "advantages", "value_targets", "values")


class _Captured:
  def __init__(self):
    self.inputs, self.loss = None, None
    self.forward_backward, self.update = None, None


class GraphedTrainer(Trainer):
  """Drop-in `Trainer` that replays CUDA graphs after `warmup` eager steps per signature."""

  def __init__(self, optimizer, anneals=None, max_grad_norm=None, grad_sync=None, warmup=3,
               keys=LOSS_KEYS):
    super().__init__(optimizer, anneals=anneals, max_grad_norm=max_grad_norm,
                     grad_sync=grad_sync)
    self.warmup = warmup
    self.keys = keys
    self._seen = {}
    self._graphs = {}
    self.replays = 0

  @staticmethod
  def _signature(data, keys):
    return tuple((k, tuple(data[k].shape), data[k].dtype) for k in keys if k in data)

  def _params(self, alg):
    return [p for p in alg.model.parameters() if p.requires_grad]

  def _capture(self, alg, sig, data):
    """Record (not run) the step on static input buffers shaped like `data`."""
    cap = _Captured()
    cap.inputs = {k: torch.empty_like(data[k]) for k, _, _ in sig}
    for k, t in cap.inputs.items():
      t.copy_(data[k])
    params = self._params(alg)
    keep = self.grad_sync is not None
    pool = torch.cuda.graph_pool_handle()
    # grads are re-created inside the capture (from the graph's private pool) unless a flat
    # all-reduce buffer owns them, in which case they are zeroed in place inside the graph
    self.optimizer.zero_grad(set_to_none=not keep)
    count = alg.loss_fn.call_count
    cap.forward_backward = torch.cuda.CUDAGraph()
    with torch.cuda.graph(cap.forward_backward, pool=pool):
      if keep:
        self.optimizer.zero_grad(set_to_none=False)
      cap.loss = alg.loss(cap.inputs)
      cap.loss.backward()
      if self.grad_sync is None:
        self._update(params)
    if self.grad_sync is not None:
      cap.update = torch.cuda.CUDAGraph()
      with torch.cuda.graph(cap.update, pool=pool):
        self._update(params)
    alg.loss_fn.call_count = count   # tracing is not a call; replays count below
    return cap

  def _update(self, params):
    if self.max_grad_norm is not None:
      torch.nn.utils.clip_grad_norm_(params, self.max_grad_norm, foreach=True)
    self.optimizer.step()

  def step(self, alg, data):
    tensors_ok = all(isinstance(data.get(k), torch.Tensor) and data[k].is_cuda
                     for k in self.keys if k in data)
    if summary.should_record() or not tensors_ok:
      return super().step(alg, data)     # logging steps and host data take the eager path
    sig = self._signature(data, self.keys)
    seen = self._seen.get(sig, 0)
    self._seen[sig] = seen + 1
    if seen < self.warmup:
      return super().step(alg, data)     # initialises optimizer state and library plans
    for anneal in self.anneals:          # in-place update of the lr tensor the graph reads
      anneal.step_to(alg.runner.step_count)
    cap = self._graphs.get(sig)
    if cap is None:
      cap = self._graphs[sig] = self._capture(alg, sig, data)
    else:
      torch._foreach_copy_([cap.inputs[k] for k, _, _ in sig], [data[k] for k, _, _ in sig])
    cap.forward_backward.replay()
    if cap.update is not None:
      self.grad_sync(alg.model)
      cap.update.replay()
    self.replays += 1
    alg.loss_fn.call_count += 1
    self.step_count += 1
    return cap.loss.detach().clone()
