"""Actor-critic policy feeding the fused PPO loss.

Reference: derl/policies.py — Policy :11-32, ActorCriticPolicy :45-80.  `act` keeps its
signature and both modes: rollout mode samples, evaluates log-prob and hands NumPy arrays
back (:76-80); training mode returns {"distribution", "values"} (:74-75).  The distribution
objects below carry the raw head outputs (`logits` or `loc`/`scale`) so that PPOLoss can
pass them straight to the fused sm_100a kernel; `log_prob` / `entropy` / `sample` remain
available with torch.distributions' formulas for any other caller.
"""
from abc import ABC, abstractmethod
import math

import torch


class Policy(ABC):
  """RL policy (typically wraps a torch.nn.Module)."""

  def is_recurrent(self):
    return False

  def get_state(self):
    return None

  def reset(self):
    pass

  @abstractmethod
  def act(self, inputs, state=None, update_state=True, training=False):
    """Returns a dict of all policy outputs (see reference docstring, :25-32)."""


class CategoricalHead:
  """Categorical(logits=...) semantics of torch.distributions on raw logits."""

  def __init__(self, logits):
    self.logits = logits

  def _log_softmax(self):
    return self.logits - torch.logsumexp(self.logits, dim=-1, keepdim=True)

  def log_prob(self, actions):
    return self._log_softmax().gather(-1, actions.long().unsqueeze(-1)).squeeze(-1)

  def entropy(self):
    logp = self._log_softmax()
    return -(logp.exp() * logp.clamp(min=torch.finfo(logp.dtype).min)).sum(-1)

  def sample(self):
    probs = torch.softmax(self.logits, dim=-1)
    flat = torch.multinomial(probs.reshape(-1, probs.shape[-1]), 1, True)
    return flat.reshape(probs.shape[:-1])


class DiagNormalHead:
  """Independent(Normal(loc, scale), 1) semantics of torch.distributions."""

  def __init__(self, loc, scale):
    self.loc, self.scale = loc, scale

  def log_prob(self, actions):
    var = self.scale ** 2
    per_dim = (-((actions - self.loc) ** 2) / (2 * var) - self.scale.log()
               - math.log(math.sqrt(2 * math.pi)))
    return per_dim.sum(-1)

  def entropy(self):
    return (0.5 + 0.5 * math.log(2 * math.pi) + torch.log(self.scale)).sum(-1)

  def sample(self):
    with torch.no_grad():
      return torch.normal(self.loc, self.scale)


def _np(tensor):
  return tensor.cpu().detach().numpy()


class ActorCriticPolicy(Policy):
  """model(observations) -> (*distribution_inputs, values)."""

  def __init__(self, model, distribution=None):
    self.model = model
    self.distribution = distribution

  def act(self, inputs, state=None, update_state=True, training=False):
    _ = update_state
    if state is not None:
      raise NotImplementedError()
    observations = inputs["observations"] if training else inputs
    if training:
      return self._heads(observations)
    # rollout mode: nothing differentiates through these outputs (the reference detaches them on
    # the way to NumPy, policies.py:76-80), so no graph — frames and activations are not kept
    # alive for a backward that never comes
    with torch.no_grad():
      out = self._heads(observations)
      actions = out["distribution"].sample()
      log_prob = out["distribution"].log_prob(actions)
    return {"actions": _np(actions), "log_prob": _np(log_prob), "values": _np(out["values"])}

  def value_tensor(self, observations):
    """Critic values of `observations` as a tensor on the model's device, no graph, no host
    round trip (extension: what GAE needs for its bootstrap, derl/runners/
    trajectory_transforms.py:47-50, without the `.cpu().numpy()` synchronisation of `act`)."""
    with torch.no_grad():
      return self.model(observations)[-1]

  def _heads(self, observations):
    *dist_inputs, values = self.model(observations)
    if self.distribution is not None:
      distribution = self.distribution(*dist_inputs)
    elif len(dist_inputs) == 1:
      distribution = CategoricalHead(dist_inputs[0])
    elif len(dist_inputs) == 2:
      distribution = DiagNormalHead(*dist_inputs)
    else:
      raise ValueError(f"model has {len(dist_inputs)} "
                       "outputs to create a distribution, "
                       "expected a single output for categorical "
                       "and two outputs for normal distributions")
    return {"distribution": distribution, "values": values}
