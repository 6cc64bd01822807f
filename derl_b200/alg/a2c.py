"""Advantage actor-critic loss on the fused loss kernels (reference:
derl/alg/a2c.py — A2CLoss :7-79, A2C :82-91).  Same constructor, entry points, errors and
logged scalars; one launch of torch.ops.derl_b200.a2c_loss_* computes the loss, the scalars
and the gradients w.r.t. the head outputs (policy term -mean(log_prob * adv), unclipped
value loss, entropy bonus).  GAE with lambda = 1 and optional whole-rollout normalisation
(derl/factory/a2c.py:53-61) is the K1 kernel, whose golden vector is the A2C fixture.
"""
import torch

from .. import ops, summary  # noqa: F401
from .common import Alg, Loss
from .ppo import STAT, _head_inputs

_K = torch.ops.derl_b200


class A2CLoss(Loss):
  """Advantage Actor Critic loss."""

  def __init__(self, policy, value_loss_coef=0.25, entropy_coef=0.01, name=None):
    super().__init__(model=policy.model, name=name)
    self.policy = policy
    self.value_loss_coef = value_loss_coef
    self.entropy_coef = entropy_coef
    self.last_stats = None

  def _f32(self, arr):
    t = self.torch_from_numpy(arr)
    return (t if t.dtype == torch.float32 else t.float()).contiguous()

  def _policy_args(self, trajectory, act):
    actions = self.torch_from_numpy(trajectory["actions"])
    advantages = self._f32(trajectory["advantages"])
    kind, *head = _head_inputs(act["distribution"])
    head = [h if h.dtype == torch.float32 else h.float() for h in head]
    log_prob_shape = head[0].shape[:-1]
    if log_prob_shape != advantages.shape:
      raise ValueError("trajectory has mismatched shapes: "
                       f"log_prob.shape={log_prob_shape} "
                       f"advantages.shape={advantages.shape}")
    width = head[0].shape[-1]
    head = [h.reshape(-1, width).contiguous() for h in head]
    actions = actions.reshape(-1).long().contiguous() if kind == "categorical" \
        else actions.reshape(-1, width).float().contiguous()
    return kind, head, actions, advantages.reshape(-1)

  def _value_args(self, trajectory, act):
    values = act["values"]
    value_targets = self._f32(trajectory["value_targets"])
    if values.shape != value_targets.shape:
      raise ValueError("trajectory has mismatched shapes "
                       f"values.shape={values.shape} "
                       f"value_targets.shape={value_targets.shape}")
    return (values if values.dtype == torch.float32 else values.float()).contiguous(), value_targets

  def _fused(self, kind, head, values, actions, advantages, value_targets, value_loss_coef):
    if kind == "gaussian":
      loss, _, _, _, stats = _K.a2c_loss_gaussian(head[0], head[1], values, actions, advantages,
                                                  value_targets, float(value_loss_coef),
                                                  float(self.entropy_coef))
    else:
      loss, _, _, stats = _K.a2c_loss_categorical(head[0] if head else None, values, actions,
                                                  advantages, value_targets,
                                                  float(value_loss_coef), float(self.entropy_coef))
    self.last_stats = stats.detach()
    return loss

  def _log(self, keys):
    for key in keys:
      summary.add_scalar(f"{self.name}/{key}", self.last_stats[STAT[key]],
                         global_step=self.call_count)

  def policy_loss(self, trajectory, act=None):
    if act is None:
      act = self.policy.act(trajectory, training=True)
    kind, head, actions, advantages = self._policy_args(trajectory, act)
    loss = self._fused(kind, head, None, actions, advantages, None, 0.)
    if summary.should_record():
      self._log(("advantages", "entropy", "policy_loss"))
    return loss

  def value_loss(self, trajectory, act=None):
    if act is None:
      act = self.policy.act(trajectory, training=True)
    values, value_targets = self._value_args(trajectory, act)
    loss = self._fused(None, None, values, None, None, value_targets, 1.)
    if summary.should_record():
      self._log(("value_targets", "value_preds", "value_loss", "r_squared"))
    return loss

  def __call__(self, data):
    act = self.policy.act(data, training=True)
    kind, head, actions, advantages = self._policy_args(data, act)
    values, value_targets = self._value_args(data, act)
    loss = self._fused(kind, head, values, actions, advantages, value_targets,
                       self.value_loss_coef)
    if summary.should_record():
      self._log(("advantages", "entropy", "policy_loss", "value_targets", "value_preds",
                 "value_loss", "r_squared"))
      summary.add_scalar(f"{self.name}/loss", loss, global_step=self.call_count)
    self.call_count += 1
    return loss


class A2C(Alg):
  """Advantage Actor Critic: runner + trainer + A2CLoss."""

  def __init__(self, runner, trainer, value_loss_coef=0.25, entropy_coef=0.01, name=None):
    loss_fn = A2CLoss(runner.policy, value_loss_coef=value_loss_coef,
                      entropy_coef=entropy_coef, name=name)
    super().__init__(runner, trainer, loss_fn, name=name)
