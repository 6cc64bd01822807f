"""Host-side mirror of the reference interface: wrappers, errors, schedules, models.  CPU."""
import numpy as np
import pytest
import torch

import derl_b200 as d
from derl_b200.runners.env_runner import RunnerWrapper


class Source:
  def __init__(self):
    self.env, self.policy, self.horizon, self.nsteps = "env", "policy", 5, 100
    self.step_count, self.nenvs, self.secret = 7, 3, "hidden"

  def is_exhausted(self):
    return False

  def __len__(self):
    return 100

  def run(self, obs=None):
    yield {}


class Passthrough(RunnerWrapper):
  def run(self, obs=None):
    yield from self.runner.run(obs)


def test_runner_wrapper_proxies_only_the_whitelist():
  w = Passthrough(Passthrough(Source()))
  assert (w.env, w.policy, w.horizon, w.nsteps, w.step_count, w.nenvs) == \
      ("env", "policy", 5, 100, 7, 3)
  assert w.is_exhausted() is False and len(w) == 100
  assert isinstance(w.unwrapped, Source)
  with pytest.raises(AttributeError, match="has no attribute 'secret'"):
    w.secret


def test_gae_errors_come_before_any_kernel():
  gae = d.GAE(policy=None)
  z = np.zeros((4, 2), np.float32)
  with pytest.raises(ValueError, match="cannot contain 'advantages'"):
    gae(dict(advantages=z))
  with pytest.raises(ValueError, match="cannot contain 'value_targets'"):
    gae(dict(value_targets=z))
  with pytest.raises(ValueError, match="or have last dimension of size 1"):
    gae(dict(rewards=z, resets=z > 0, values=np.zeros((4, 2, 2), np.float32)))
  with pytest.raises(ValueError, match="or have last dimension of size 1"):
    gae(dict(rewards=z, resets=z > 0, values=np.zeros((4,), np.float32)))


def test_merge_time_batch_is_a_view():
  traj = dict(resets=torch.zeros(4, 3, dtype=torch.bool), observations=torch.zeros(4, 3, 5, 2),
              values=np.zeros((4, 3, 1), np.float32), state=dict(x=1))
  d.MergeTimeBatch()(traj)
  assert traj["observations"].shape == (12, 5, 2) and traj["values"].shape == (12, 1)
  assert traj["resets"].shape == (12,) and traj["state"] == dict(x=1)
  with pytest.raises(AssertionError):
    d.MergeTimeBatch()(dict(resets=torch.zeros(4)))


def test_linear_anneal_closed_form_equals_stepwise_loop():
  for start, end, nsteps in ((2.5e-4, 0., 1000), (1., 3., 50), (0.1, 0.1, 10)):
    fast = d.LinearAnneal(start, nsteps, end)
    for target in (0, 1, 17, 500, 1000, 1500):
      fast.step_to(target)
      # the reference's per-step update (derl/anneal.py:77-86), iterated `target` times
      value = start
      for count in range(1, target + 1):
        value = min(max(start + (end - start) * (count / nsteps), min(start, end)),
                    max(start, end))
      assert fast.step_count == target
      np.testing.assert_allclose(float(fast.get_tensor()), np.float32(value), rtol=1e-7)
    with pytest.raises(ValueError, match="cannot be smaller"):
      fast.step_to(3)
  shared = d.LinearAnneal(1e-3, 10)
  tensor = shared.get_tensor()
  shared.step()
  assert float(tensor) == pytest.approx(9e-4)  # in place: the optimizer keeps seeing it


def test_summary_gate_contract():
  s = d.summary
  s.start_recording()
  assert s.should_record()
  s.set_writer(None)
  with pytest.raises(ValueError, match="summary.writer cannot be None"):
    s.add_scalar("a", 1.0)
  seen = []

  class W:
    def add_scalar(self, *a, **k):
      seen.append((a, k))
  s.set_writer(W())
  s.add_scalar("tag", 2.0, global_step=3)
  assert seen == [(("tag", 2.0), dict(global_step=3))]
  s.set_recording(False)
  assert not s.should_record()
  s.set_writer(None)


def test_policy_known_answers_from_reference_tests():
  """derl/policies_test.py:11-29 constants (seed 0, CPU): Gaussian head reproduces actions,
  log-prob and value; categorical head reproduces value and the log-prob of action 3 (the
  sampled action itself depends on torch's multinomial stream, which changed after 1.5)."""
  torch.manual_seed(0)
  model = d.MuJoCoModel(3, (2, 1)).to("cpu")
  act = d.ActorCriticPolicy(model).act(torch.randn(3))
  assert list(act.keys()) == ["actions", "log_prob", "values"]
  np.testing.assert_allclose(act["actions"], [-1.7938228, 1.0464325], rtol=1e-6)
  np.testing.assert_allclose(act["log_prob"], -3.7467263, rtol=1e-6)
  np.testing.assert_allclose(act["values"], [-0.18482158], rtol=1e-6)

  torch.manual_seed(0)
  model = d.NatureCNNModel((6, 1)).to("cpu")
  policy = d.ActorCriticPolicy(model)
  obs = torch.rand(84, 84, 4)
  out = policy.act({"observations": obs}, training=True)
  np.testing.assert_allclose(out["values"].detach().numpy(), [0.257305294], rtol=1e-5)
  logp3 = out["distribution"].log_prob(torch.tensor(3))
  np.testing.assert_allclose(float(logp3), -1.80754196, rtol=1e-5)


def test_model_state_dict_layout():
  cnn = d.NatureCNNModel([4, 1]).to("cpu")
  keys = list(cnn.state_dict())
  assert keys == ["base.conv-0.weight", "base.conv-0.bias", "base.conv-1.weight",
                  "base.conv-1.bias", "base.conv-2.weight", "base.conv-2.bias",
                  "base.linear.weight", "base.linear.bias", "output_layers.0.weight",
                  "output_layers.0.bias", "output_layers.1.weight", "output_layers.1.bias"]
  assert sum(p.numel() for p in cnn.parameters()) == 1686693
  mlp = d.MuJoCoModel(17, [6, 1]).to("cpu")
  assert next(iter(mlp.state_dict())) == "logstd"
  assert sum(p.numel() for p in mlp.parameters()) == 11085
  loc, std, values = mlp(torch.zeros(5, 17))
  assert loc.shape == (5, 6) and std.shape == (5, 6) and values.shape == (5, 1)
  loc1, std1, v1 = mlp(torch.zeros(17))
  assert loc1.shape == (6,) and v1.shape == (1,)
  assert torch.all(cnn.output_layers[0].bias == 0)
  w = cnn.base[0].weight.detach().reshape(32, -1)
  np.testing.assert_allclose((w @ w.T).numpy(), np.eye(32), atol=1e-5)  # orthogonal rows


def test_ppo_loss_value_errors_before_dispatch():
  model = d.MuJoCoModel(3, [2, 1]).to("cpu")
  loss = d.PPOLoss(d.ActorCriticPolicy(model))
  assert loss.name == "ppo" and loss.call_count == 0
  obs = np.zeros((4, 3))
  base = dict(observations=obs, actions=np.zeros((4, 2), np.float32),
              log_prob=np.zeros(4, np.float32), values=np.zeros((4, 1), np.float32))
  with pytest.raises(ValueError, match="does not contain 'advantages'"):
    loss.policy_loss(dict(base))
  with pytest.raises(ValueError, match="does not contain 'value_targets'"):
    loss.value_loss(dict(base, advantages=np.zeros(4, np.float32)))
  with pytest.raises(ValueError, match="mismatched shapes"):
    loss.policy_loss(dict(base, advantages=np.zeros((4, 1), np.float32)))
  with pytest.raises(ValueError, match="mismatched shapes"):
    loss.value_loss(dict(base, value_targets=np.zeros(4, np.float32)))


def test_minibatch_iterator_requires_resident_rollout():
  class OneShot(Source):
    def run(self, obs=None):
      yield dict(observations=np.zeros((8, 2)), state={})
  it = d.IterateWithMinibatches(OneShot(), num_epochs=1, num_minibatches=2)
  with pytest.raises(TypeError, match="resident on the GPU"):
    next(it.run())


def test_synthetic_rollout_shapes_match_env_runner_layout():
  r = d.make_rollout("atari", 4, 3, device="cpu", seed=1)
  assert r["observations"].shape == (4, 3, 84, 84, 4) and r["observations"].dtype == np.uint8
  assert r["actions"].dtype == np.int64 and r["values"].shape == (4, 3, 1)
  assert r["rewards"].dtype == np.float64 and set(np.unique(r["rewards"])) <= {-1., 0., 1.}
  assert r["resets"].dtype == np.bool_ and r["state"]["latest_observations"].shape == (3, 84, 84, 4)
  m = d.make_rollout("mujoco", 16, None, device="cpu", seed=1)
  assert m["observations"].shape == (16, 17) and m["observations"].dtype == np.float64
  assert m["actions"].shape == (16, 6) and m["values"].shape == (16, 1)
  assert m["rewards"].shape == (16,) and m["state"]["latest_observations"].shape == (17,)
  src = d.SyntheticRolloutRunner(policy=None, kind="mujoco", nenvs=None, horizon=16, nsteps=32,
                                 device="cpu")
  assert len(list(src.run())) == 2 and src.step_count == 32 and src.is_exhausted()


class CountingEnv:
  """4 batched toy envs: observation = step counter, episode ends every 3 steps."""
  nenvs = 4

  def __init__(self):
    self.t = 0

  @property
  def unwrapped(self):
    return self

  def reset(self):
    self.t = 0
    return np.full((4, 2), self.t, np.float32)

  def step(self, actions):
    self.t += 1
    obs = np.full((4, 2), self.t, np.float32)
    return obs, np.arange(4, dtype=np.float64) + self.t, np.full(4, self.t % 3 == 0), [{}] * 4


class EchoPolicy:
  def act(self, obs, state=None, update_state=True, training=False):
    return dict(actions=np.zeros(4, np.int64), log_prob=np.zeros(4, np.float32),
                values=np.asarray(obs)[:, :1] * 0.5)

  def is_recurrent(self):
    return False


def test_env_runner_list_and_resident_modes_agree():
  """EnvRunner(resident_device=...) fills preallocated [T, N, ...] tensors step by step and
  yields the same rollout as the reference-style per-step lists (minus next_observations)."""
  plain = next(d.EnvRunner(CountingEnv(), EchoPolicy(), horizon=5, nsteps=40).run())
  runner = d.EnvRunner(CountingEnv(), EchoPolicy(), horizon=5, nsteps=40, resident_device="cpu")
  gen = runner.run()
  resident = next(gen)
  assert runner.step_count == 20 and not runner.is_exhausted()
  assert "next_observations" in plain and "next_observations" not in resident
  assert isinstance(resident["infos"], list) and len(resident["infos"]) == 5
  for key in ("observations", "actions", "log_prob", "values", "rewards", "resets"):
    want = np.asarray(plain[key])
    assert isinstance(resident[key], torch.Tensor)
    np.testing.assert_array_equal(resident[key].numpy(), want)
    assert resident[key].numpy().dtype == want.dtype
  np.testing.assert_array_equal(resident["state"]["latest_observations"],
                                plain["state"]["latest_observations"])
  second = next(gen)
  assert float(second["observations"][0, 0, 0]) == 5.0 and runner.is_exhausted()
  with pytest.raises(ValueError, match="must contain 'actions'"):
    class NoActions(EchoPolicy):
      def act(self, obs, **kw):
        return dict(values=np.zeros((4, 1), np.float32))
    next(d.EnvRunner(CountingEnv(), NoActions(), 2, 10).run())


def test_row_selection_is_a_lazy_descriptor():
  """RowSelection (fused minibatch gather, runners/row_selection.py): shape / slicing / row
  indices are pure bookkeeping; materialising needs the CUDA gather kernel (no CPU path)."""
  from derl_b200.runners.row_selection import RowSelection, dense
  source = torch.arange(10 * 6, dtype=torch.uint8).reshape(10, 2, 3)
  perm = torch.tensor([3, 1, 4, 1, 5, 9, 2, 6])
  sel = RowSelection(source, perm, 2, 5)
  assert sel.shape == (5, 2, 3) and len(sel) == 5 and sel.ndim == 3 and sel.dtype == torch.uint8
  assert sel.size(0) == 5 and sel.dim() == 3 and sel.is_contiguous() and not sel.is_cuda
  assert sel.rows.tolist() == [4, 1, 5, 9, 2]
  assert sel[()] is sel
  part = sel[1:4]
  assert isinstance(part, RowSelection) and part.rows.tolist() == [1, 5, 9] and part.source is source
  assert sel[3:].rows.tolist() == [9, 2] and len(sel[7:]) == 0
  assert "perm[2:7]" in repr(sel)
  with pytest.raises(NotImplementedError):   # gather_rows is registered for CUDA only
    sel.materialize()
  with pytest.raises(NotImplementedError):
    sel[::2]
  with pytest.raises(ValueError):
    RowSelection(source, perm, 5, 4)
  with pytest.raises(TypeError):
    RowSelection(source, perm.int(), 0, 2)
  with pytest.raises(TypeError):
    RowSelection(source.permute(1, 0, 2), perm, 0, 2)
  assert dense(source) is source


@pytest.mark.parametrize("axis", [0, 1, -1])
@pytest.mark.parametrize("indices", [2, -1, [0, 2, 2, 1], [-1, 0], [[0, 1], [2, -3]], []])
def test_take_follows_np_take_for_tensors_and_arrays(axis, indices):
  """derl/runners/trajectory_transforms.py:95-103: np.take for every key but "state".  Tensor
  values (CPU here, index_select; CUDA rows via the gather kernel in the GPU suite) must give
  what np.take gives for the same arrays: scalar index drops the axis, negative indices wrap,
  index arrays of any rank replace the axis by their shape."""
  rng = np.random.RandomState(0)
  obs = rng.randint(0, 255, size=(5, 3, 4)).astype(np.uint8)
  rew = rng.randn(5, 3, 4)
  state = dict(latest_observations=obs[0])
  as_np = dict(observations=obs.copy(), rewards=rew.copy(), state=state)
  as_t = dict(observations=torch.from_numpy(obs.copy()), rewards=torch.from_numpy(rew.copy()),
              state=state)
  take = d.Take(indices, axis=axis)
  take(as_np)
  take(as_t)
  for key in ("observations", "rewards"):
    want = np.take({"observations": obs, "rewards": rew}[key], indices, axis=axis)
    np.testing.assert_array_equal(as_np[key], want)
    assert tuple(as_t[key].shape) == want.shape
    np.testing.assert_array_equal(as_t[key].numpy(), want)
  assert as_np["state"] is state and as_t["state"] is state


def test_take_rejects_out_of_range_indices_like_numpy():
  """np.take raises IndexError for an index outside [-n, n); the device gather kernels do no
  bounds checks, so Take validates the (host) indices before any launch."""
  x = np.arange(12.).reshape(3, 4)
  for bad, axis in ((3, 0), (-4, 0), ([0, 4], 1), ([-5], 1)):
    with pytest.raises(IndexError):
      np.take(x, bad, axis=axis)
    with pytest.raises(IndexError, match="out of bounds"):
      d.Take(bad, axis=axis)(dict(x=torch.from_numpy(x)))
  with pytest.raises(IndexError, match="axis 2 is out of bounds"):
    d.Take([0], axis=2)(dict(x=torch.from_numpy(x)))
  with pytest.raises(TypeError):
    d.Take([0.5], axis=0)(dict(x=torch.from_numpy(x)))


def test_take_default_axis_selects_envs():
  """Default axis=1 is the env axis of a [T, N, ...] rollout (the reference's use: keep a subset
  of environments)."""
  traj = dict(rewards=torch.arange(12.).reshape(4, 3), resets=np.zeros((4, 3), bool))
  d.Take([2, 0])(traj)
  assert traj["rewards"].tolist() == [[2., 0.], [5., 3.], [8., 6.], [11., 9.]]
  assert traj["resets"].shape == (4, 2)


@pytest.mark.parametrize("kind,nenvs", [("atari", 3), ("mujoco", None)])
def test_bench_reference_arm_generates_the_same_rollout_without_importing_the_package(kind, nenvs):
  """bench.py's reference arm must not import derl_b200 (no CUDA library in that process), so it
  carries its own copy of the seeded host rollout generator: the copy must stay equal to
  derl_b200.make_rollout(device="cpu") array for array."""
  import importlib.util
  import os
  root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
  spec = importlib.util.spec_from_file_location("_bench_under_test", os.path.join(root, "bench.py"))
  bench = importlib.util.module_from_spec(spec)
  spec.loader.exec_module(bench)
  theirs = bench.host_rollout(kind, 5, nenvs, seed=7)
  ours = d.make_rollout(kind, 5, nenvs, device="cpu", seed=7)
  assert set(theirs) == set(ours)
  for key in ours:
    if key == "state":
      np.testing.assert_array_equal(theirs[key]["latest_observations"],
                                    ours[key]["latest_observations"])
    else:
      assert theirs[key].dtype == ours[key].dtype, key
      np.testing.assert_array_equal(theirs[key], ours[key])
  # and the arm's CPU path runs on it (oracle port here unless the reference tree is present)
  sec, how = bench.cpu_update_seconds(kind, nenvs, 5, 1, 1, steps=1, warmup=0)
  assert sec > 0 and how in ("live", "port")


def test_fused_heads_plumbing_defers_and_restores_the_trunk_bias():
  """NatureCNNModel._fused_heads asks the trunk to leave its last bias out (K9 adds it inside the
  heads kernel); off the GPU the same code path must add it back and use the nn.Linear heads.
  The deferred bias is an nn.Parameter handed over without registering it a second time."""
  import torch.nn as nn

  class Trunk(nn.Module):
    defer_linear_bias, deferred_bias = False, None

    def __init__(self):
      super().__init__()
      self.linear = nn.Linear(12, 512)

    def forward(self, x):
      if self.defer_linear_bias:
        self.__dict__["deferred_bias"] = self.linear.bias
        return nn.functional.linear(x, self.linear.weight, None)
      return self.linear(x)

  torch.manual_seed(0)
  model = d.NatureCNNModel([3, 1]).cpu()
  model.base = Trunk()
  x = torch.randn(5, 12)
  want = [layer(model.base(x)) for layer in model.output_layers]
  got = model._fused_heads(x)
  assert model.base.defer_linear_bias is False and model.base.deferred_bias is None
  assert [n for n, _ in model.base.named_parameters()] == ["linear.weight", "linear.bias"]
  for a, b in zip(got, want):
    np.testing.assert_allclose(a.detach().numpy(), b.detach().numpy(), rtol=1e-5, atol=1e-6)
  assert not model._heads_fusable()   # CPU parameters: the library heads
