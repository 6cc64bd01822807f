"""Differential tests against the LIVE reference (dev container only: needs /root/reference or
$DERL_REF; skipped on the GPU box, where the committed golden vectors stand in).  CPU only:
they pin the oracle and the host-side mirror, the CUDA path is compared with the oracle in
tests/test_gpu_parity.py."""
import numpy as np
import pytest
import torch

import derl_b200 as d
from oracle import derl_oracle as O


class ConstPolicy:
  def __init__(self, last_value):
    self.last_value, self.model = last_value, None

  def act(self, inputs, state=None, update_state=True, training=False):
    return {"values": self.last_value}

  def is_recurrent(self):
    return False


@pytest.mark.parametrize("seed", range(12))
def test_oracle_gae_equals_reference_on_random_rollouts(ref, seed):
  rng = np.random.RandomState(seed)
  nsteps = int(rng.randint(1, 70))
  batched = seed % 3 != 0
  lead = (nsteps, int(rng.randint(1, 40))) if batched else (nsteps,)
  rdtype = (np.float64, np.float32)[seed % 2]
  rewards = (rng.standard_normal(lead) * 10 ** rng.uniform(-2, 2)).astype(rdtype)
  values = rng.standard_normal(lead + (1,)).astype(np.float32)
  resets = rng.random(lead) < rng.uniform(0, .5)
  last_value = rng.standard_normal((lead[1], 1) if batched else (1,)).astype(np.float32)
  gamma, lam = float(rng.uniform(.5, 1)), float(rng.uniform(0, 1))
  for normalize in (False, True, None):
    traj = dict(rewards=rewards, values=values, resets=resets,
                state=dict(latest_observations=None))
    want_a, want_vt = ref.GAE(ConstPolicy(last_value), gamma=gamma, lambda_=lam,
                              normalize=normalize)(traj)
    a, vt = O.gae(rewards, values, resets, last_value, gamma, lam, normalize=normalize)
    np.testing.assert_array_equal(a, want_a)
    np.testing.assert_array_equal(vt, want_vt)
  a_c, vt_c = O.gae_c(rewards.reshape(nsteps, -1), values.reshape(nsteps, -1),
                      resets.reshape(nsteps, -1), last_value, gamma, lam)
  a, vt = O.gae(rewards, values, resets, last_value, gamma, lam, normalize=False)
  np.testing.assert_array_equal(a_c.reshape(a.shape), a)
  np.testing.assert_array_equal(vt_c.reshape(vt.shape), vt)
  if a.size > 1:
    np.testing.assert_array_equal(O.normalize_c(a), ref.NormalizeAdvantages and
                                  O.normalize_advantages(a.reshape(-1)))


def test_reference_gae_errors_match_the_drop_in(ref):
  z = np.zeros((4, 2), np.float32)
  cases = [dict(advantages=z), dict(value_targets=z),
           dict(rewards=z, resets=z > 0, values=np.zeros((4, 2, 2), np.float32))]
  for traj in cases:
    with pytest.raises(ValueError) as want:
      ref.GAE(ConstPolicy(z))(dict(traj))
    with pytest.raises(ValueError) as got:
      d.GAE(ConstPolicy(z))(dict(traj))
    assert str(got.value) == str(want.value)


@pytest.mark.parametrize("kind", ["atari", "mujoco"])
def test_seeded_models_are_identical_to_the_reference(ref, kind):
  """Same construction order -> same RNG consumption -> identical initial weights, identical
  state_dict keys and parameter order (grads.npz fixtures index parameters by position)."""
  def build(pkg):
    torch.manual_seed(0)
    model = pkg.NatureCNNModel([6, 1]) if kind == "atari" else pkg.MuJoCoModel(17, [6, 1])
    return model.to("cpu")
  ours, theirs = build(d), build(ref)
  assert list(ours.state_dict()) == list(theirs.state_dict())
  for (k, a), b in zip(ours.state_dict().items(), theirs.state_dict().values()):
    assert torch.equal(a, b), k
  assert [tuple(p.shape) for p in ours.parameters()] == [tuple(p.shape) for p in theirs.parameters()]
  obs = torch.randint(0, 256, (3, 84, 84, 4), dtype=torch.uint8) if kind == "atari" \
      else torch.randn(3, 17, dtype=torch.float64)
  for a, b in zip(ours(obs), theirs(obs)):
    assert torch.allclose(a, b, rtol=1e-4, atol=1e-6)  # NHWC-strided vs NCHW conv paths
  # the oracle's plain modules load the same parameters positionally
  cpu = O.NatureCNN(6) if kind == "atari" else O.MuJoCoMLP(17, 6)
  assert [tuple(p.shape) for p in cpu.parameters()] == [tuple(p.shape) for p in ours.parameters()]


def test_policy_heads_equal_torch_distributions(ref):
  torch.manual_seed(3)
  logits = torch.randn(50, 7)
  acts = torch.randint(0, 7, (50,))
  want = torch.distributions.Categorical(logits=logits)
  head = d.policies.CategoricalHead(logits)
  assert torch.allclose(head.log_prob(acts), want.log_prob(acts), atol=1e-6)
  assert torch.allclose(head.entropy(), want.entropy(), atol=1e-6)
  loc, scale = torch.randn(50, 4), torch.rand(50, 4) + .1
  cont = torch.randn(50, 4)
  want = torch.distributions.Independent(torch.distributions.Normal(loc, scale), 1)
  head = d.policies.DiagNormalHead(loc, scale)
  assert torch.allclose(head.log_prob(cont), want.log_prob(cont), atol=1e-5)
  assert torch.allclose(head.entropy(), want.entropy(), atol=1e-6)
  # rollout-mode act(): same keys, same values as the reference policy on the same model + RNG
  torch.manual_seed(0)
  model = d.MuJoCoModel(5, [3, 1]).to("cpu")
  obs = torch.randn(9, 5)
  torch.manual_seed(1)
  ours = d.ActorCriticPolicy(model).act(obs)
  torch.manual_seed(1)
  theirs = ref.ActorCriticPolicy(model).act(obs)
  assert list(ours) == list(theirs)
  for k in ours:
    np.testing.assert_allclose(ours[k], theirs[k], rtol=1e-6, atol=1e-6)


def test_linear_anneal_equals_reference(ref):
  for start, end, nsteps, target in ((2.5e-4, 0., 500, 321), (1., 5., 40, 100), (3e-4, 0., 7, 7)):
    ours, theirs = d.LinearAnneal(start, nsteps, end), ref.LinearAnneal(start, nsteps, end)
    assert ours.name == theirs.name == "linear_anneal"
    ours.step_to(target)
    theirs.step_to(target)
    np.testing.assert_allclose(float(ours.get_tensor()), float(theirs.get_tensor()), rtol=1e-6)
    assert ours.step_count == theirs.step_count == target
    np.testing.assert_allclose(float(ours.step()), float(theirs.step()), rtol=1e-6)


def test_runner_wrapper_contract_equals_reference(ref):
  class Source:
    env, policy, horizon, nsteps, step_count, nenvs, other = 1, 2, 3, 4, 5, 6, 7

    def is_exhausted(self):
      return True

    def __len__(self):
      return 4

    def run(self, obs=None):
      yield dict(observations=[np.zeros(2)], state={})

  for pkg in (ref, d):
    w = pkg.TransformInteractions(Source(), [], asarray=False)
    assert [getattr(w, k) for k in ("env", "policy", "horizon", "nsteps", "step_count", "nenvs")] \
        == [1, 2, 3, 4, 5, 6]
    assert w.is_exhausted() and len(w) == 4 and isinstance(w.unwrapped, Source)
    with pytest.raises(AttributeError):
      w.other
    assert next(w.run())["state"] == {}
  assert d.IterateWithMinibatches(Source()).num_epochs == ref.IterateWithMinibatches(Source()).num_epochs == 3
  assert d.IterateWithMinibatches(Source()).num_minibatches == 4
  loss_ref, loss_d = ref.PPOLoss.__init__.__defaults__, d.PPOLoss.__init__.__defaults__
  assert loss_ref == loss_d == (0.2, 0.25, 0.01, None)
  assert ref.GAE.__init__.__defaults__ == d.GAE.__init__.__defaults__[:4] == (0.99, 0.95, None, 1e-8)


def test_oracle_full_update_equals_reference_pipeline(ref):
  """oracle.ppo_update == the reference's ppo_runner_wrap + PPO + Trainer on a fresh seeded
  case (not one of the committed golden files)."""
  torch.manual_seed(5)
  rng = np.random.RandomState(5)
  nsteps, nenvs, nact = 8, 3, 5
  rollout = dict(observations=rng.randint(0, 256, (nsteps, nenvs, 84, 84, 4)).astype(np.uint8),
                 actions=rng.randint(0, nact, (nsteps, nenvs)).astype(np.int64),
                 log_prob=(rng.standard_normal((nsteps, nenvs)) * .05 - 1.6).astype(np.float32),
                 values=(rng.standard_normal((nsteps, nenvs, 1)) * .1).astype(np.float32),
                 rewards=rng.standard_normal((nsteps, nenvs)), resets=rng.random((nsteps, nenvs)) < .2,
                 state=dict(latest_observations=rng.randint(0, 256, (nenvs, 84, 84, 4)).astype(np.uint8)))
  model_ref = ref.NatureCNNModel([nact, 1]).to("cpu")
  model_cpu = O.NatureCNN(nact)
  model_cpu.load_state_dict(dict(zip(model_cpu.state_dict().keys(), model_ref.state_dict().values())))

  env = type("E", (), {"nenvs": nenvs, "unwrapped": property(lambda s: s)})()
  ref_policy = ref.ActorCriticPolicy(model_ref)

  class Source:
    horizon, step_count = 8, 0

    def __init__(self):
      self.env, self.policy, self.nsteps, self.nenvs = env, ref_policy, 10, nenvs

    def run(self, obs=None):
      yield {k: (dict(v) if k == "state" else np.array(v)) for k, v in rollout.items()}

  runner = ref.ppo_runner_wrap(Source(), num_epochs=2, num_minibatches=3)
  alg = ref.PPO(runner, ref.Trainer(torch.optim.Adam(model_ref.parameters(), lr=1e-3, eps=1e-5),
                                    max_grad_norm=.5), cliprange=.1)
  np.random.seed(9)
  want = [float(alg.step(b).detach()) for b in runner.run()]
  np.random.seed(9)
  with torch.no_grad():
    last_value = model_cpu(torch.from_numpy(rollout["state"]["latest_observations"]))[-1].numpy()
  cols = {k: v for k, v in rollout.items() if k != "state"}
  got = O.ppo_update(model_cpu, torch.optim.Adam(model_cpu.parameters(), lr=1e-3, eps=1e-5), cols,
                     last_value, num_epochs=2, num_minibatches=3, cliprange=.1)
  np.testing.assert_allclose(got, want, rtol=1e-6, atol=1e-7)
