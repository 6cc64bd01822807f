"""Whole-update fast path for the MuJoCo-shaped actor-critic (BASELINE configs[1]).

`PPO.learn()` on the default pipeline (`ppo_runner_wrap` + `Trainer` + Adam) with a
`MuJoCoModel` of two 64-64 tanh MLPs spends its time on ~150 tiny launches per optimiser
step, 320 steps per rollout (derl/runners/onpolicy.py:51-62, derl/alg/common.py:66-78).  When
`plan()` recognises exactly that pipeline, one rollout's update — every epoch and minibatch —
is ONE launch of the persistent kernel K8 (`torch.ops.derl_b200.ppo_mlp_update`,
csrc/mlp_update.cu): same minibatch membership (the global NumPy RNG is consumed exactly like
IterateWithMinibatches does), same normalisation, loss, clipping and Adam arithmetic, the
optimizer's own state tensors updated in place.  Anything it does not recognise (other
models, optimisers, wrappers, summaries being recorded) takes the ordinary per-minibatch path.
"""
import numpy as np
import torch
from torch import nn

from .. import _lib, ops, summary  # noqa: F401
from ..models import MLP, MuJoCoModel
from ..runners.onpolicy import IterateWithMinibatches, TransformInteractions
from ..runners.trajectory_transforms import NormalizeAdvantages
from .common import Trainer

_K = torch.ops.derl_b200


def _mlp_tensors(mlp, in_features, out_features):
  """[W1, b1, W2, b2, W3, b3] of a Linear-Tanh-Linear-Tanh-Linear stack of width 64, or None."""
  layers = list(mlp.children())
  if len(layers) != 5 or not all(isinstance(layers[i], nn.Linear) for i in (0, 2, 4)) \
      or not all(type(layers[i]) is nn.Tanh for i in (1, 3)):
    return None
  shapes = [(64, in_features), (64, 64), (out_features, 64)]
  tensors = []
  for layer, shape in zip((layers[0], layers[2], layers[4]), shapes):
    if tuple(layer.weight.shape) != shape or layer.bias is None:
      return None
    tensors += [layer.weight, layer.bias]
  return tensors


class FusedMLPUpdate:
  """The recognised pipeline, bound to one `Alg`; `run(rollout)` performs one whole update."""

  def __init__(self, alg, outer, iterate, inner, normalize, tensors, obs_dim, act_dim):
    self.alg, self.outer, self.iterate, self.inner = alg, outer, iterate, inner
    self.normalize = normalize
    self.tensors, self.obs_dim, self.act_dim = tensors, obs_dim, act_dim

  @staticmethod
  def plan(alg):
    """A FusedMLPUpdate when `alg` is the stock PPO pipeline on a fusable model, else None."""
    from .ppo import PPOLoss
    model, trainer, loss_fn = alg.model, alg.trainer, alg.loss_fn
    if type(trainer) is not Trainer or trainer.grad_sync is not None or trainer.micro_batch:
      return None
    if type(loss_fn) is not PPOLoss or loss_fn.policy.model is not model \
        or getattr(loss_fn.policy, "distribution", None) is not None:
      return None
    outer = alg.runner
    normalize = None
    if type(outer) is TransformInteractions:
      if len(outer.transforms) != 1 or type(outer.transforms[0]) is not NormalizeAdvantages \
          or outer.transforms[0].group is not None:
        return None
      normalize = outer.transforms[0]
      iterate = outer.runner
    else:
      iterate = outer
    if type(iterate) is not IterateWithMinibatches or iterate.fused_gather:
      return None
    inner = iterate.runner
    if type(model) is not MuJoCoModel or len(model.module_list) != 2:
      return None
    if not all(type(m) is MLP for m in model.module_list):
      return None
    first = model.module_list[0][0]
    obs_dim, act_dim = first.in_features, model.logstd.numel()
    policy_t = _mlp_tensors(model.module_list[0], obs_dim, act_dim)
    value_t = _mlp_tensors(model.module_list[1], obs_dim, 1)
    if policy_t is None or value_t is None:
      return None
    tensors = policy_t + value_t + [model.logstd]
    if not all(t.is_cuda and t.dtype == torch.float32 and t.is_contiguous() and t.requires_grad
               for t in tensors):
      return None
    if _lib.load().derl_b200_ppo_mlp_update_smem_bytes(obs_dim, act_dim) == 0:
      return None
    opt = trainer.optimizer
    if type(opt) is not torch.optim.Adam or len(opt.param_groups) != 1:
      return None
    group = opt.param_groups[0]
    if group.get("amsgrad") or group.get("maximize") or group.get("weight_decay", 0) != 0 \
        or group.get("differentiable") or group.get("decoupled_weight_decay"):
      return None
    if {id(p) for p in group["params"]} != {id(t) for t in tensors} \
        or len(group["params"]) != len(tensors):
      return None
    return FusedMLPUpdate(alg, outer, iterate, inner, normalize, tensors, obs_dim, act_dim)

  # ------------------------------------------------------------------------------- one rollout
  def _columns(self, rollout):
    """The six flat float columns the kernel reads, or None when the rollout is not the shape
    this path handles (then the caller uses the per-minibatch path for it)."""
    need = ("observations", "actions", "log_prob", "values", "advantages", "value_targets")
    if not all(isinstance(rollout.get(k), torch.Tensor) and rollout[k].is_cuda for k in need):
      return None
    obs = rollout["observations"]
    if obs.dim() != 2 or obs.shape[1] != self.obs_dim or not obs.is_contiguous() \
        or obs.dtype not in (torch.float32, torch.float64):
      return None
    size = obs.shape[0]
    actions = rollout["actions"]
    if tuple(actions.shape) != (size, self.act_dim) or actions.dtype != torch.float32:
      return None
    cols = [obs, actions.contiguous()]
    for key in ("log_prob", "advantages", "value_targets", "values"):
      col = rollout[key]
      if col.dtype != torch.float32 or col.numel() != size:
        return None
      cols.append(col.reshape(size).contiguous())
    return cols

  def _adam_state(self):
    """exp_avg / exp_avg_sq tensor lists (created like torch.optim.Adam creates them) and the
    number of steps taken so far."""
    opt = self.alg.trainer.optimizer
    group = opt.param_groups[0]
    on_device = bool(group.get("capturable") or group.get("fused"))
    exp_avg, exp_avg_sq, step = [], [], None
    for p in self.tensors:
      state = opt.state[p]
      if len(state) == 0:
        state["step"] = torch.zeros((), dtype=torch.float32,
                                    device=p.device if on_device else "cpu")
        state["exp_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
        state["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
      exp_avg.append(state["exp_avg"])
      exp_avg_sq.append(state["exp_avg_sq"])
      if step is None:
        step = int(state["step"])
    return exp_avg, exp_avg_sq, step

  def run(self, rollout):
    """One PPO update on `rollout` (as the GAE stage yields it); returns the [nsteps] losses, or
    None when this rollout has to take the per-minibatch path."""
    cols = self._columns(rollout)
    iterate, trainer, alg = self.iterate, self.alg.trainer, self.alg
    if cols is None or summary.should_record():
      return None
    obs, actions, log_prob, advantages, value_targets, values = cols
    size = obs.shape[0]
    mbsize = size // iterate.num_minibatches   # 0 -> range() step 0 ValueError in the generic path
    if mbsize < 1:
      return None
    orders, order = [], None
    for _ in range(iterate.num_epochs):          # the RNG draws of IterateWithMinibatches.run
      if iterate.shuffle_before_epoch:
        draw = np.random.permutation(size)
        order = draw if order is None else order[draw]
      elif order is None:
        order = np.arange(size)
      orders.append(order)
    perm = torch.from_numpy(np.ascontiguousarray(np.concatenate(orders), dtype=np.int64)).to(
        obs.device)
    for anneal in trainer.anneals:
      anneal.step_to(alg.runner.step_count)
    group = trainer.optimizer.param_groups[0]
    exp_avg, exp_avg_sq, step = self._adam_state()
    loss_fn = alg.loss_fn
    with torch.no_grad():
      losses, stats = _K.ppo_mlp_update(
          self.tensors, exp_avg, exp_avg_sq, obs, actions, log_prob, advantages, value_targets,
          values, perm, iterate.num_epochs, mbsize, self.normalize is not None,
          float(self.normalize.epsilon) if self.normalize is not None else 0.0,
          loss_fn.cliprange, float(loss_fn.value_loss_coef), float(loss_fn.entropy_coef),
          trainer.max_grad_norm, float(group["lr"]), float(group["betas"][0]),
          float(group["betas"][1]), float(group["eps"]), step)
    nsteps = losses.numel()
    for p in self.tensors:
      trainer.optimizer.state[p]["step"] += nsteps
      p.grad = None
    trainer.step_count += nsteps
    loss_fn.call_count += nsteps
    loss_fn.last_stats = stats[-1]
    return losses
