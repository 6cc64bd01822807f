"""World-size-2 tests of the env-axis data-parallel host logic on the gloo backend (CPU).
The data path itself has no collective; the exchange steps are the flat gradient all-reduce
and the 3-double moments all-reduce (derl_b200/parallel.py)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from derl_b200 import parallel


def free_port():
  with socket.socket() as s:
    s.bind(("127.0.0.1", 0))
    return s.getsockname()[1]


def _worker(rank, world, port, results):
  os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank),
                    WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
  got_rank, got_world, _ = parallel.init_from_env(backend="gloo")
  assert (got_rank, got_world) == (rank, world)
  torch.manual_seed(0)
  model = torch.nn.Sequential(torch.nn.Linear(6, 5), torch.nn.Tanh(), torch.nn.Linear(5, 1))
  sync = parallel.GradientAllReduce(model)
  # global batch of 8 samples, each rank owns 4: mean over the shard, then averaged grads
  g = torch.Generator().manual_seed(1)
  x, y = torch.randn(8, 6, generator=g), torch.randn(8, 1, generator=g)
  lo, hi = parallel.env_shard(8, rank, world)
  opt = torch.optim.SGD(model.parameters(), lr=0.1)
  opt.zero_grad(set_to_none=False)
  ((model(x[lo:hi]) - y[lo:hi]) ** 2).mean().backward()
  sync(model)
  got = [p.grad.clone() for p in model.parameters()]   # views into the flat buffer
  # single-process answer on the full batch
  torch.manual_seed(0)
  ref = torch.nn.Sequential(torch.nn.Linear(6, 5), torch.nn.Tanh(), torch.nn.Linear(5, 1))
  ((ref(x) - y) ** 2).mean().backward()
  ok_grad = all(torch.allclose(g, p.grad, rtol=1e-5, atol=1e-7)
                for g, p in zip(got, ref.parameters()))
  ok_grad = ok_grad and sum(g.numel() for g in got) == sync.flat.numel()
  views_ok = all(p.grad.data_ptr() >= sync.flat.data_ptr() for p in model.parameters())
  # moments all-reduce: every shard normalises with global statistics
  adv = torch.arange(8, dtype=torch.float64)[lo:hi]
  moments = torch.stack([adv.sum(), (adv * adv).sum(), torch.tensor(float(hi - lo),
                                                                    dtype=torch.float64)])
  dist.all_reduce(moments)
  full = torch.arange(8, dtype=torch.float64)
  ok_moments = torch.allclose(moments, torch.stack([full.sum(), (full * full).sum(),
                                                    torch.tensor(8., dtype=torch.float64)]))
  results[rank] = (ok_grad, views_ok, ok_moments)
  dist.destroy_process_group()


def test_gradient_allreduce_and_moments_world2():
  world, port = 2, free_port()
  manager = mp.Manager()
  results = manager.dict()
  mp.spawn(_worker, args=(world, port, results), nprocs=world, join=True)
  assert dict(results) == {0: (True, True, True), 1: (True, True, True)}


def test_env_shard_and_rollout_slicing():
  assert parallel.env_shard(32768, 3, 8) == (12288, 16384)
  with pytest.raises(ValueError, match="divisible"):
    parallel.env_shard(10, 0, 4)
  rollout = dict(observations=np.arange(4 * 6 * 2).reshape(4, 6, 2), rewards=np.zeros((4, 6)),
                 state=dict(latest_observations=np.arange(6)))
  part = parallel.shard_rollout(rollout, 1, 3)
  assert part["observations"].shape == (4, 2, 2) and part["rewards"].shape == (4, 2)
  np.testing.assert_array_equal(part["observations"], rollout["observations"][:, 2:4])
  np.testing.assert_array_equal(part["state"]["latest_observations"], [2, 3])


def test_flat_gradient_views_keep_the_parameter_layout():
  """channels_last conv weights get channels_last gradient views into the flat buffer."""
  model = torch.nn.Sequential(torch.nn.Conv2d(4, 8, 3), torch.nn.Flatten(), torch.nn.Linear(8, 2))
  model.to(memory_format=torch.channels_last)
  sync = parallel.GradientAllReduce(model)
  for p in model.parameters():
    assert p.grad.stride() == p.stride() and p.grad.shape == p.shape
  model(torch.randn(5, 4, 3, 3)).sum().backward()
  want = [p.grad.clone() for p in model.parameters()]
  opt = torch.optim.Adam(model.parameters(), lr=1e-3)
  opt.zero_grad(set_to_none=False)
  assert float(sync.flat.abs().sum()) == 0.0
  model(torch.randn(5, 4, 3, 3)).sum().backward()
  opt.step()
  assert all(p.grad.data_ptr() >= sync.flat.data_ptr() for p in model.parameters())
  assert sum(w.numel() for w in want) == sync.flat.numel()


def test_single_process_is_a_noop():
  for key in ("WORLD_SIZE", "RANK", "LOCAL_RANK"):
    os.environ.pop(key, None)
  assert parallel.init_from_env() == (0, 1, 0)
  model = torch.nn.Linear(3, 2)
  sync = parallel.GradientAllReduce(model)
  model(torch.ones(1, 3)).sum().backward()
  before = sync.flat.clone()
  sync(model)
  assert torch.equal(sync.flat, before) and before.abs().sum() > 0
