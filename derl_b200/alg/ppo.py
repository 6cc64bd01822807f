"""PPO loss on the fused sm_100a kernel (reference: derl/alg/ppo.py — PPOLoss :8-108,
PPO :111-123; Schulman et al. 2017).

`PPOLoss(policy, cliprange, value_loss_coef, entropy_coef, name)` and its three entry
points (`__call__`, `policy_loss`, `value_loss`) keep the reference's signatures, errors and
logged scalars.  What differs is the execution: instead of ~60 ATen ops forward and ~80
backward through torch.distributions, one launch of torch.ops.derl_b200.ppo_loss_* reads
the head outputs once and writes the loss, the logged scalars and d loss / d(logits | loc,
scale, values); autograd only scales those saved gradients by the incoming grad.
"""
import torch

from .. import ops, summary  # noqa: F401
from .common import Alg, Loss

_K = torch.ops.derl_b200
# indices into the stats vector written by the kernel (include/derl_b200.h)
STAT = dict(loss=0, policy_loss=1, entropy=2, value_loss=3, advantages=4, value_targets=5,
            value_preds=6, r_squared=7, clip_fraction=8, approx_kl=9)


def _head_inputs(distribution):
  """("categorical", logits) or ("gaussian", loc, scale) from whatever act() returned."""
  if hasattr(distribution, "logits") and not hasattr(distribution, "base_dist"):
    return ("categorical", distribution.logits)
  base = getattr(distribution, "base_dist", distribution)
  if hasattr(base, "loc") and hasattr(base, "scale"):
    return ("gaussian", base.loc, base.scale)
  raise ValueError(f"unsupported distribution {type(distribution).__name__}: expected a "
                   "categorical (logits) or diagonal normal (loc, scale) head")


class PPOLoss(Loss):
  """Clipped-surrogate + clipped-value + entropy loss."""

  def __init__(self, policy, cliprange=0.2, value_loss_coef=0.25, entropy_coef=0.01, name=None):
    super().__init__(model=policy.model, name=name)
    self.policy = policy
    self.cliprange = cliprange
    self.value_loss_coef = value_loss_coef
    self.entropy_coef = entropy_coef
    self.last_stats = None  # device tensor f32[16] of the latest call (see STAT)

  # ----------------------------------------------------------------- argument marshalling
  def _f32(self, arr):
    t = self.torch_from_numpy(arr)
    return (t if t.dtype == torch.float32 else t.float()).contiguous()

  def _policy_args(self, trajectory, act):
    if "advantages" not in trajectory:
      raise ValueError("trajectory does not contain 'advantages'")
    old_log_prob = self._f32(trajectory["log_prob"])
    advantages = self._f32(trajectory["advantages"])
    actions = self.torch_from_numpy(trajectory["actions"])
    kind, *head = _head_inputs(act["distribution"])
    head = [h if h.dtype == torch.float32 else h.float() for h in head]
    # shape of distribution.log_prob(actions): batch dims of the head
    log_prob_shape = head[0].shape[:-1]
    if kind == "gaussian" and tuple(actions.shape) != tuple(head[0].shape):
      raise ValueError("trajectory has mismatched shapes: "
                       f"actions.shape={tuple(actions.shape)} loc.shape={tuple(head[0].shape)}")
    if kind == "categorical" and tuple(actions.shape) != tuple(log_prob_shape):
      raise ValueError("trajectory has mismatched shapes: "
                       f"actions.shape={tuple(actions.shape)} "
                       f"log_prob.shape={tuple(log_prob_shape)}")
    if log_prob_shape != old_log_prob.shape:
      raise ValueError("trajectory has mismatched shapes: "
                       f"log_prob.shape={log_prob_shape} "
                       f"old_log_prob.shape={old_log_prob.shape}")
    if log_prob_shape != advantages.shape:
      raise ValueError("trajectory has mismatched shapes: "
                       f"log_prob.shape={log_prob_shape} "
                       f"advantages.shape={advantages.shape}")
    width = head[0].shape[-1]
    head = [h.reshape(-1, width).contiguous() for h in head]
    if kind == "categorical":
      actions = actions.reshape(-1).long().contiguous()
    else:
      actions = actions.reshape(-1, width).float().contiguous()
    return kind, head, actions, old_log_prob.reshape(-1), advantages.reshape(-1)

  def _value_args(self, trajectory, act):
    if "value_targets" not in trajectory:
      raise ValueError("trajectory does not contain 'value_targets'")
    value_targets = self._f32(trajectory["value_targets"])
    old_values = self._f32(trajectory["values"])
    values = act["values"]
    if values.shape != value_targets.shape:
      raise ValueError("trajectory has mismatched shapes "
                       f"values.shape={values.shape} "
                       f"value_targets.shape={value_targets.shape}")
    values = (values if values.dtype == torch.float32 else values.float()).contiguous()
    return values, value_targets, old_values

  def _fused(self, kind, head, values, actions, old_log_prob, advantages, value_targets,
             old_values, value_loss_coef):
    """One launch of the fused kernel; `head`/`values` may be None to skip that side."""
    if kind == "gaussian":
      loc, scale = head
      loss, _, _, _, stats = _K.ppo_loss_gaussian(
          loc, scale, values, actions, old_log_prob, advantages, value_targets, old_values,
          self.cliprange, float(value_loss_coef), float(self.entropy_coef))
    else:  # categorical head, or the value-only call (no head at all)
      logits = head[0] if head else None
      loss, _, _, stats = _K.ppo_loss_categorical(
          logits, values, actions, old_log_prob, advantages, value_targets, old_values,
          self.cliprange, float(value_loss_coef), float(self.entropy_coef))
    self.last_stats = stats.detach()
    return loss

  def _log(self, tag_prefix, keys):
    for key in keys:
      summary.add_scalar(f"{tag_prefix}/{key}", self.last_stats[STAT[key]],
                         global_step=self.call_count)

  # ----------------------------------------------------------------- reference entry points
  def policy_loss(self, trajectory, act=None):
    """mean max(-r*A, -clip(r)*A) - entropy_coef * mean entropy (reference :24-64)."""
    if act is None:
      act = self.policy.act(trajectory, training=True)
    kind, head, actions, old_log_prob, advantages = self._policy_args(trajectory, act)
    loss = self._fused(kind, head, None, actions, old_log_prob, advantages, None, None, 0.)
    if summary.should_record():
      self._log(self.name, ("advantages", "policy_loss", "entropy"))
    return loss

  def value_loss(self, trajectory, act=None):
    """mean max((v-vt)^2, (clipped v - vt)^2) (reference :66-98)."""
    if act is None:
      act = self.policy.act(trajectory, training=True)
    values, value_targets, old_values = self._value_args(trajectory, act)
    loss = self._fused(None, None, values, None, None, None, value_targets, old_values, 1.)
    if summary.should_record():
      self._log("ppo", ("value_loss", "value_targets", "value_preds", "r_squared"))
    return loss

  def __call__(self, data):
    act = self.policy.act(data, training=True)
    kind, head, actions, old_log_prob, advantages = self._policy_args(data, act)
    values, value_targets, old_values = self._value_args(data, act)
    loss = self._fused(kind, head, values, actions, old_log_prob, advantages, value_targets,
                       old_values, self.value_loss_coef)
    if summary.should_record():
      self._log(self.name, ("advantages", "policy_loss", "entropy"))
      self._log("ppo", ("value_loss", "value_targets", "value_preds", "r_squared"))
      summary.add_scalar("ppo/loss", loss, global_step=self.call_count)
    self.call_count += 1
    return loss


class PPO(Alg):
  """Proximal Policy Optimization: runner + trainer + PPOLoss."""

  def __init__(self, runner, trainer, cliprange=0.2, value_loss_coef=0.25, entropy_coef=0.01,
               name=None):
    loss_fn = PPOLoss(runner.policy, cliprange=cliprange, value_loss_coef=value_loss_coef,
                      entropy_coef=entropy_coef, name=name)
    super().__init__(runner, trainer, loss_fn, name=name)

  fused_update = True   # class-wide switch: let learn() use the whole-update kernel (K8)
  last_losses = None

  def learn(self, progress=True):
    """Reference semantics (derl/alg/common.py:100-106: step on every minibatch until the
    runner is exhausted).  When the pipeline is the stock one on a fusable model
    (alg/fused_mlp.py) each rollout's epochs x minibatches run as one launch of the persistent
    update kernel instead of ~150 launches per minibatch."""
    from .fused_mlp import FusedMLPUpdate
    plan = FusedMLPUpdate.plan(self) if self.fused_update else None
    if plan is None:
      return super().learn(progress=progress)
    from tqdm import tqdm
    with tqdm(total=len(self.runner), disable=not progress) as pbar:
      for rollout in plan.inner.run():
        pbar.update(self.runner.step_count - pbar.n)
        self.last_losses = plan.run(rollout)   # [nsteps] device tensor of this update's losses
        if self.last_losses is None:           # e.g. summaries being recorded: ordinary path
          for data in plan.iterate.minibatches(rollout):
            if plan.normalize is not None:
              plan.normalize(data)
            self.step(data)
