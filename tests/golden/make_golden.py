"""Generates the golden vectors under tests/golden/ from the LIVE reference.

Run in the dev container only (needs the read-only reference tree at /root/reference or
$DERL_REF; `gym` / `atari_py` are replaced by the import-only stubs in tests/_stubs):

    python tests/golden/make_golden.py

Two families of files are written:
  ref_*.npz   re-packed golden vectors that the reference's OWN tests hold for this path
              (testdata/a2c/atari/interactions.npz, testdata/ppo/pybullet/*), stripped to
              the arrays the PPO data path needs so they stay small;
  live_*.npz  inputs + outputs of the reference's unmodified classes (GAE,
              IterateWithMinibatches, NormalizeAdvantages, PPOLoss, Trainer) on seeded
              synthetic inputs, covering what the reference's tests never exercise
              (resets, lambda < 1, normalize, ragged minibatches, categorical loss, ties).
The GPU box has no reference tree: tests there read only these files.
"""
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get("DERL_REF", "/root/reference")
sys.path.insert(0, os.path.join(REPO, "tests", "_stubs"))
sys.path.insert(0, REF)

import derl  # noqa: E402  (the reference)
import derl.summary as ref_summary  # noqa: E402

ref_summary.stop_recording()


def save(name, **arrays):
  path = os.path.join(HERE, name)
  np.savez_compressed(path, **arrays)
  print(f"{name}: {os.path.getsize(path)} bytes, keys={sorted(arrays)}")


class ConstPolicy:
  """policy.act(...)["values"] == a fixed bootstrap value array."""

  def __init__(self, last_value, model=None):
    self.last_value, self.model = last_value, model

  def act(self, inputs, state=None, update_state=True, training=False):
    return {"values": self.last_value}

  def is_recurrent(self):
    return False


class ArrayRunner:
  """Rollout source yielding prepared arrays (the surface EnvRunner exposes)."""

  def __init__(self, rollouts, policy, nenvs, horizon):
    self.rollouts, self.policy, self.horizon = rollouts, policy, horizon
    self.env = types.SimpleNamespace(nenvs=nenvs)
    self.env.unwrapped = self.env
    self.nenvs, self.nsteps, self.step_count = nenvs, 10 ** 9, 0

  def run(self, obs=None):
    for rollout in self.rollouts:
      self.step_count += self.horizon * (self.nenvs or 1)
      yield {k: (dict(v) if k == "state" else np.array(v)) for k, v in rollout.items()}


# ----------------------------------------------------------------------------- ref_*
def repack_reference_fixtures():
  with np.load(os.path.join(REF, "testdata/a2c/atari/interactions.npz"), allow_pickle=True) as f:
    # 40 = T5 x N8 merged time-major; A2C factory: gamma .99, lambda 1, normalize False
    save("ref_a2c_atari_gae.npz", rewards=f["rewards"].reshape(5, 8),
         resets=f["resets"].reshape(5, 8), values=f["values"].reshape(5, 8, 1),
         advantages=f["advantages"].reshape(5, 8),
         value_targets=f["value_targets"].reshape(5, 8, 1), gamma=0.99, lambda_=1.0)
  with np.load(os.path.join(REF, "testdata/ppo/pybullet/interactions.npz"),
               allow_pickle=True) as f:
    batch = {k: f[k] for k in ("observations", "actions", "log_prob", "values", "rewards",
                               "resets", "advantages", "value_targets")}
  with np.load(os.path.join(REF, "testdata/ppo/pybullet/grads.npz")) as f:
    grads = {k: f[k] for k in f.files}
  losses = np.load(os.path.join(REF, "testdata/ppo/pybullet/losses.npy"))
  # mujoco PPO defaults (derl/factory/ppo.py:35-49): cliprange .2, vcoef .25, ecoef 0
  save("ref_ppo_pybullet.npz", loss0=losses[0], cliprange=0.2, value_loss_coef=0.25,
       entropy_coef=0.0, **batch, **grads)


# ----------------------------------------------------------------------------- live GAE
def gae_case(rng, nsteps, nenvs, rdtype, reset_prob, gamma, lambda_, normalize, scale=1.0,
             special=None):
  lead = (nsteps,) if nenvs is None else (nsteps, nenvs)
  rewards = (rng.randn(*lead) * scale).astype(rdtype)
  values = (rng.randn(*lead, 1) * scale).astype(np.float32)
  resets = rng.rand(*lead) < reset_prob
  last_value = (rng.randn(*((1,) if nenvs is None else (nenvs, 1))) * scale).astype(np.float32)
  if special == "reset_last_row":
    resets[-1] = True
  if special == "all_reset":
    resets[:] = True
  if special == "clip_rewards":
    rewards = np.sign(rewards) * (rng.rand(*lead) < 0.3)
    rewards = rewards.astype(rdtype)
  traj = dict(rewards=rewards, values=values, resets=resets,
              state=dict(latest_observations=np.zeros(1)))
  adv, targets = derl.GAE(ConstPolicy(last_value), gamma=gamma, lambda_=lambda_,
                          normalize=normalize)(traj)
  return dict(rewards=rewards, values=values, resets=resets, last_value=last_value,
              gamma=gamma, lambda_=lambda_,
              normalize=-1 if normalize is None else int(normalize),
              advantages=adv, value_targets=targets)


def live_gae():
  rng = np.random.RandomState(1234)
  specs = [
      (16, 48, np.float64, .1, .99, .95, False, 1., None),
      (7, 33, np.float32, .2, .99, .95, True, 1., None),
      (1, 16, np.float64, .5, .99, .95, False, 1., None),
      (40, None, np.float64, .1, .99, .95, False, 1., None),       # unbatched (MuJoCo)
      (12, 32, np.float64, .0, .9, 1., False, 1., "all_reset"),
      (12, 20, np.float32, .05, .99, .95, False, 1., "reset_last_row"),
      (33, 64, np.float64, .02, .999, .97, None, 100., None),       # large magnitudes
      (128, 8, np.float64, .01, .99, .95, False, 1., "clip_rewards"),  # Atari defaults
      (9, 5, np.float32, .3, .5, .0, False, 1e-3, None),            # lambda = 0
  ]
  out = {}
  for i, spec in enumerate(specs):
    for key, val in gae_case(rng, *spec).items():
      out[f"c{i}_{key}"] = val
  out["ncases"] = len(specs)
  save("live_gae.npz", **out)


# ----------------------------------------------------------------------------- live minibatches
def live_minibatches():
  """Which rows each minibatch holds and its normalised advantages, under np.random.seed."""
  out = {}
  specs = [(8, 6, 3, 4), (10, 5, 2, 4), (5, 3, 3, 2)]  # S = 48, 50 (ragged: +1 short), 15 (ragged)
  for i, (nsteps, nenvs, epochs, nmb) in enumerate(specs):
    rng = np.random.RandomState(100 + i)
    size = nsteps * nenvs
    rollout = dict(
        observations=rng.randint(0, 256, (nsteps, nenvs, 4, 3, 2)).astype(np.uint8),
        ids=np.arange(size, dtype=np.int64).reshape(nsteps, nenvs),
        actions=rng.randint(0, 4, (nsteps, nenvs)).astype(np.int64),
        log_prob=rng.randn(nsteps, nenvs).astype(np.float32),
        values=rng.randn(nsteps, nenvs, 1).astype(np.float32),
        rewards=rng.randn(nsteps, nenvs),
        resets=rng.rand(nsteps, nenvs) < .1,
        state=dict(latest_observations=np.zeros((nenvs, 4, 3, 2), np.uint8)))
    last_value = rng.randn(nenvs, 1).astype(np.float32)
    policy = ConstPolicy(last_value)
    runner = derl.ppo_runner_wrap(ArrayRunner([rollout], policy, nenvs, nsteps),
                                  num_epochs=epochs, num_minibatches=nmb)
    np.random.seed(7 + i)
    ids, advs, sizes = [], [], []
    for batch in runner.run():
      ids.append(batch["ids"])
      advs.append(batch["advantages"])
      sizes.append(batch["ids"].shape[0])
      assert np.array_equal(batch["observations"],
                            rollout["observations"].reshape(size, 4, 3, 2)[batch["ids"]])
    for key, val in rollout.items():
      if key != "state":
        out[f"c{i}_{key}"] = val
    out[f"c{i}_last_value"] = last_value
    out[f"c{i}_seed"] = 7 + i
    out[f"c{i}_epochs"], out[f"c{i}_nmb"] = epochs, nmb
    out[f"c{i}_mb_sizes"] = np.asarray(sizes)
    out[f"c{i}_mb_ids"] = np.concatenate(ids)
    out[f"c{i}_mb_advantages"] = np.concatenate(advs)
  out["ncases"] = len(specs)
  save("live_minibatches.npz", **out)


# ----------------------------------------------------------------------------- live PPO loss
class HeadPolicy:
  """act(training=True) returns distributions built from given leaf tensors."""

  def __init__(self, dist_inputs, values):
    self.dist_inputs, self.values = dist_inputs, values
    self.model = torch.nn.Linear(1, 1)  # only to give Loss.device a parameter (CPU)

  def act(self, inputs, state=None, update_state=True, training=False):
    if len(self.dist_inputs) == 1:
      dist = torch.distributions.Categorical(logits=self.dist_inputs[0])
    else:
      dist = torch.distributions.Independent(torch.distributions.Normal(*self.dist_inputs), 1)
    return {"distribution": dist, "values": self.values}


class ScalarLog:
  def __init__(self):
    self.scalars = {}

  def add_scalar(self, tag, value, global_step=None):
    self.scalars[tag] = float(value)


def loss_case(rng, kind, nbatch, width, cliprange, vcoef, ecoef):
  old_log_prob = rng.randn(nbatch).astype(np.float32) * .3 - 1.
  adv = rng.randn(nbatch).astype(np.float32)
  adv[::7] = 0.                                   # torch.max ties on the policy side
  values = rng.randn(nbatch, 1).astype(np.float32)
  old_values = (values + rng.randn(nbatch, 1) * .3).astype(np.float32)
  targets = (values + rng.randn(nbatch, 1)).astype(np.float32)
  old_values[::5] = values[::5]                   # d == 0: u^2 == w^2 ties on the value side
  if kind == "categorical":
    head = [(rng.randn(nbatch, width) * 2).astype(np.float32)]
    actions = rng.randint(0, width, nbatch).astype(np.int64)
  else:
    head = [rng.randn(nbatch, width).astype(np.float32),
            np.exp(rng.randn(nbatch, width) * .3).astype(np.float32)]
    actions = (head[0] + rng.randn(nbatch, width) * head[1]).astype(np.float32)
  batch = dict(actions=actions, log_prob=old_log_prob, advantages=adv, value_targets=targets,
               values=old_values)
  leaves = [torch.tensor(h, requires_grad=True) for h in head]
  v_leaf = torch.tensor(values, requires_grad=True)
  loss_fn = derl.PPOLoss(HeadPolicy(leaves, v_leaf), cliprange=cliprange,
                         value_loss_coef=vcoef, entropy_coef=ecoef)
  log = ScalarLog()
  ref_summary.set_writer(log)
  ref_summary.start_recording()
  loss = loss_fn(batch)
  ref_summary.stop_recording()
  loss.backward()
  out = dict(batch)
  out.update(kind=kind, cliprange=-1. if cliprange is None else cliprange, vcoef=vcoef,
             ecoef=ecoef, pred_values=values, loss=loss.detach().numpy(),
             dvalues=v_leaf.grad.numpy())
  for i, (h, leaf) in enumerate(zip(head, leaves)):
    out[f"head{i}"], out[f"dhead{i}"] = h, leaf.grad.numpy()
  for tag, val in log.scalars.items():
    out["log_" + tag.replace("/", "_")] = np.float32(val)
  # the two public partial entry points on fresh leaves
  for name in ("policy_loss", "value_loss"):
    leaves2 = [torch.tensor(h, requires_grad=True) for h in head]
    v2 = torch.tensor(values, requires_grad=True)
    fn = derl.PPOLoss(HeadPolicy(leaves2, v2), cliprange=cliprange, value_loss_coef=vcoef,
                      entropy_coef=ecoef)
    out[name] = getattr(fn, name)(batch).detach().numpy()
  return out


def live_ppo_loss():
  rng = np.random.RandomState(4321)
  specs = [("categorical", 257, 6, .2, .25, .01), ("categorical", 64, 4, .1, .25, .01),
           ("categorical", 100, 18, None, .5, .0), ("categorical", 33, 1, .2, .25, .01),
           ("gaussian", 257, 6, .2, .25, .0), ("gaussian", 50, 3, .3, 1., .02),
           ("gaussian", 31, 1, None, .25, .01), ("gaussian", 64, 17, .2, .25, .001)]
  out = {}
  for i, spec in enumerate(specs):
    for key, val in loss_case(rng, *spec).items():
      out[f"c{i}_{key}"] = val
  out["ncases"] = len(specs)
  save("live_ppo_loss.npz", **out)


def live_a2c_loss():
  """A2CLoss forward/backward + logged scalars (derl/alg/a2c.py) on both heads."""
  rng = np.random.RandomState(777)
  out = {}
  specs = [("categorical", 129, 6, .5, .01), ("gaussian", 77, 3, .25, .02),
           ("categorical", 40, 18, 1., 0.)]
  for i, (kind, nbatch, width, vcoef, ecoef) in enumerate(specs):
    adv = rng.randn(nbatch).astype(np.float32)
    values = rng.randn(nbatch, 1).astype(np.float32)
    targets = (values + rng.randn(nbatch, 1)).astype(np.float32)
    if kind == "categorical":
      head = [(rng.randn(nbatch, width) * 2).astype(np.float32)]
      actions = rng.randint(0, width, nbatch).astype(np.int64)
    else:
      head = [rng.randn(nbatch, width).astype(np.float32),
              np.exp(rng.randn(nbatch, width) * .3).astype(np.float32)]
      actions = (head[0] + rng.randn(nbatch, width) * head[1]).astype(np.float32)
    batch = dict(actions=actions, advantages=adv, value_targets=targets)
    leaves = [torch.tensor(h, requires_grad=True) for h in head]
    v_leaf = torch.tensor(values, requires_grad=True)
    loss_fn = derl.A2CLoss(HeadPolicy(leaves, v_leaf), value_loss_coef=vcoef, entropy_coef=ecoef)
    log = ScalarLog()
    ref_summary.set_writer(log)
    ref_summary.start_recording()
    loss = loss_fn(batch)
    ref_summary.stop_recording()
    loss.backward()
    case = dict(batch, kind=kind, vcoef=vcoef, ecoef=ecoef, pred_values=values,
                loss=loss.detach().numpy(), dvalues=v_leaf.grad.numpy())
    for j, (h, leaf) in enumerate(zip(head, leaves)):
      case[f"head{j}"], case[f"dhead{j}"] = h, leaf.grad.numpy()
    for tag, val in log.scalars.items():
      case["log_" + tag.replace("/", "_")] = np.float32(val)
    for key, val in case.items():
      out[f"c{i}_{key}"] = val
  out["ncases"] = len(specs)
  save("live_a2c_loss.npz", **out)


# ----------------------------------------------------------------------------- live full update
def live_update(name, kind):
  """Seeded model + synthetic rollouts through the reference's whole PPO pipeline:
  ppo_runner_wrap -> PPOLoss -> Trainer(Adam eps 1e-5, max_grad_norm .5).step."""
  torch.manual_seed(0)
  rng = np.random.RandomState(55)
  if kind == "mujoco":
    nsteps, nenvs, epochs, nmb, obs_dim, act_dim = 32, None, 2, 4, 5, 2
    model = derl.MuJoCoModel(obs_dim, [act_dim, 1])
    hp = dict(cliprange=.2, value_loss_coef=.25, entropy_coef=0.)
    lr = 3e-4
  else:
    nsteps, nenvs, epochs, nmb, nact = 6, 4, 2, 2, 4
    model = derl.NatureCNNModel([nact, 1])
    hp = dict(cliprange=.1, value_loss_coef=.25, entropy_coef=.01)
    lr = 2.5e-4
  model.to("cpu")
  policy = derl.ActorCriticPolicy(model)
  rollouts = []
  for _ in range(2):
    if kind == "mujoco":
      r = dict(observations=rng.randn(nsteps, obs_dim),
               actions=rng.randn(nsteps, act_dim).astype(np.float32),
               log_prob=(rng.randn(nsteps) * .1 - 2.8).astype(np.float32),
               values=(rng.randn(nsteps, 1) * .1).astype(np.float32),
               rewards=rng.randn(nsteps), resets=rng.rand(nsteps) < .1,
               state=dict(latest_observations=rng.randn(obs_dim)))
    else:
      r = dict(observations=rng.randint(0, 256, (nsteps, nenvs, 84, 84, 4)).astype(np.uint8),
               actions=rng.randint(0, nact, (nsteps, nenvs)).astype(np.int64),
               log_prob=(rng.randn(nsteps, nenvs) * .05 - np.log(nact)).astype(np.float32),
               values=(rng.randn(nsteps, nenvs, 1) * .1).astype(np.float32),
               rewards=np.sign(rng.randn(nsteps, nenvs)) * (rng.rand(nsteps, nenvs) < .3),
               resets=rng.rand(nsteps, nenvs) < .1,
               state=dict(latest_observations=rng.randint(
                   0, 256, (nenvs, 84, 84, 4)).astype(np.uint8)))
    rollouts.append(r)
  source = ArrayRunner(rollouts, policy, nenvs, nsteps)
  runner = derl.ppo_runner_wrap(source, num_epochs=epochs, num_minibatches=nmb)
  optimizer = torch.optim.Adam(model.parameters(), lr=lr, eps=1e-5)
  trainer = derl.Trainer(optimizer, max_grad_norm=.5)
  alg = derl.PPO(runner, trainer, **hp)
  np.random.seed(11)
  losses = [float(alg.step(batch).detach()) for batch in runner.run()]
  out = dict(losses=np.asarray(losses, np.float32), seed=11, epochs=epochs, nmb=nmb, lr=lr,
             nrollouts=len(rollouts), nenvs=-1 if nenvs is None else nenvs, **hp)
  for i, r in enumerate(rollouts):
    for key, val in r.items():
      if key == "state":
        out[f"r{i}_latest_observations"] = val["latest_observations"]
      else:
        out[f"r{i}_{key}"] = val
  final = torch.cat([p.detach().reshape(-1) for p in model.parameters()])
  out["final_param_sum"] = np.float64(final.double().sum())
  out["final_param_abs_sum"] = np.float64(final.double().abs().sum())
  if kind == "mujoco":
    for k, v in model.state_dict().items():
      out["final_" + k] = v.numpy()
  save(name, **out)


def main():
  repack_reference_fixtures()
  live_gae()
  live_minibatches()
  live_ppo_loss()
  live_a2c_loss()
  live_update("live_update_mujoco.npz", "mujoco")
  live_update("live_update_atari.npz", "atari")


if __name__ == "__main__":
  main()
