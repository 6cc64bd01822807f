"""Env-axis data parallelism on real GPUs (NCCL): numeric equivalence of dp-2 with one GPU fed
the union minibatches (SURVEY.md §7 "DP equivalence", §8e; insertion point of the gradient
all-reduce: derl/alg/common.py:70-71).  Needs >= 2 GPUs: `gpurun --gpus 2 -- pytest -m gpu
tests/test_multi_gpu.py`; skipped on a single-GPU box."""
import os
import socket

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

NENVS, HORIZON, EPOCHS, NMB = 8, 6, 2, 2
HP = dict(cliprange=0.1, value_loss_coef=0.25, entropy_coef=0.01)


def _free_port():
  with socket.socket() as s:
    s.bind(("127.0.0.1", 0))
    return s.getsockname()[1]


def _host_rollout():
  rng = np.random.RandomState(7)
  return dict(
      observations=rng.randint(0, 256, (HORIZON, NENVS, 84, 84, 4)).astype(np.uint8),
      actions=rng.randint(0, 4, (HORIZON, NENVS)).astype(np.int64),
      log_prob=(rng.randn(HORIZON, NENVS) * .05 - np.log(4)).astype(np.float32),
      values=(rng.randn(HORIZON, NENVS, 1) * .1).astype(np.float32),
      rewards=np.sign(rng.randn(HORIZON, NENVS)) * (rng.rand(HORIZON, NENVS) < .3),
      resets=rng.rand(HORIZON, NENVS) < .1,
      state=dict(latest_observations=rng.randint(0, 256, (NENVS, 84, 84, 4)).astype(np.uint8)))


class _Source:
  def __init__(self, rollout, policy, nenvs):
    self.rollout, self.policy, self.nenvs, self.horizon = rollout, policy, nenvs, HORIZON
    self.env = type("E", (), {"nenvs": nenvs, "unwrapped": property(lambda s: s)})()
    self.nsteps, self.step_count = 10 ** 9, 0

  def run(self, obs=None):
    self.step_count += self.horizon * self.nenvs
    yield {k: (dict(v) if k == "state" else v) for k, v in self.rollout.items()}


def _worker(rank, world, port, results, overlap):
  os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank),
                    WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
  import torch.distributed as dist
  import derl_b200 as d
  from derl_b200 import parallel
  from derl_b200.runners.onpolicy import gather_minibatch
  d.summary.stop_recording()
  parallel.init_from_env("nccl")
  device = torch.device("cuda", rank)
  torch.backends.cudnn.allow_tf32 = False
  torch.backends.cuda.matmul.allow_tf32 = False
  full = _host_rollout()

  def make_alg(rollout, nenvs, sync_factory, group, micro_batch):
    torch.manual_seed(0)   # identical initial weights everywhere
    model = d.NatureCNNModel([4, 1])
    policy = d.ActorCriticPolicy(model)
    runner = d.TransformInteractions(_Source(rollout, policy, nenvs),
                                     [d.GAE(policy, normalize=False), d.MergeTimeBatch()])
    runner = d.IterateWithMinibatches(runner, EPOCHS, NMB)
    runner = d.TransformInteractions(runner, [d.NormalizeAdvantages(group=group)])
    optimizer = torch.optim.Adam(model.parameters(), lr=2.5e-4, eps=1e-5)
    trainer = d.Trainer(optimizer, max_grad_norm=.5, grad_sync=sync_factory(model),
                        micro_batch=micro_batch)
    return d.PPO(runner, trainer, **HP), model, runner

  # ---- dp-2: this rank owns envs [4 rank, 4 rank + 4), local permutations (seed differs per rank)
  shard = parallel.shard_rollout(full, rank, world)
  alg, model, runner = make_alg(shard, NENVS // world,
                                lambda m: parallel.GradientAllReduce(m, overlap=overlap),
                                dist.group.WORLD, micro_batch=5)   # ragged micro-batches of 12 rows
  np.random.seed(100 + rank)
  state = np.random.get_state()
  losses = [alg.step(batch).item() for batch in runner.run()]
  # the permutations IterateWithMinibatches drew (same RNG stream, replayed)
  np.random.set_state(state)
  size = HORIZON * NENVS // world
  orders, order = [], None
  for _ in range(EPOCHS):
    draw = np.random.permutation(size)
    order = draw if order is None else order[draw]
    orders.append(order)
  gathered = [None] * world
  dist.all_gather_object(gathered, orders)
  params = torch.cat([p.detach().reshape(-1) for p in model.parameters()])
  other = [torch.empty_like(params) for _ in range(world)]
  dist.all_gather(other, params)
  replicas_equal = all(torch.equal(other[0], o) for o in other)

  # ---- one GPU fed the union minibatches (rank 0 only): same rows, global mean
  ok, err = True, 0.0
  if rank == 0:
    ref_alg, ref_model, _ = make_alg(full, NENVS, lambda m: None, None, micro_batch=None)
    rollout = {k: (dict(v) if k == "state" else torch.from_numpy(np.ascontiguousarray(v)).to(device))
               for k, v in full.items()}
    d.GAE(ref_alg.loss_fn.policy, normalize=False)(rollout)
    d.MergeTimeBatch()(rollout)
    per = NENVS // world
    mbsize = size // NMB
    ref_losses = []
    for e in range(EPOCHS):
      for j in range(NMB):
        rows = []
        for r in range(world):
          local = gathered[r][e][j * mbsize:(j + 1) * mbsize]          # t * per + n_local
          rows.append((local // per) * NENVS + r * per + local % per)  # t * NENVS + n_global
        perm = torch.from_numpy(np.concatenate(rows)).to(device)
        batch = gather_minibatch(rollout, perm, 0, perm.numel())
        d.NormalizeAdvantages()(batch)
        ref_losses.append(ref_alg.step(batch).item())
    want = torch.cat([p.detach().reshape(-1) for p in ref_model.parameters()])
    err = (params - want).abs().max().item()
    moved = (want - torch.cat([p.detach().reshape(-1) for p in
                               make_alg(full, NENVS, lambda m: None, None, None)[1].parameters()])
             ).abs().max().item()
    # mean of the two shard losses == loss of the union minibatch (equal shard sizes)
    all_losses = [None] * world
    dist.all_gather_object(all_losses, losses)
    mean_losses = np.mean(np.asarray(all_losses), axis=0)
    loss_err = float(np.abs(mean_losses - np.asarray(ref_losses)).max())
    ok = err <= 2e-6 and loss_err <= 2e-6 and moved > 1e-4
    results[0] = (ok, replicas_equal, err, loss_err, moved)
  else:
    all_losses = [None] * world
    dist.all_gather_object(all_losses, losses)
    results[rank] = (True, replicas_equal, 0.0, 0.0, 0.0)
  dist.barrier()
  dist.destroy_process_group()


@pytest.mark.parametrize("overlap", [True, False])
def test_dp2_equals_one_gpu_on_the_union_minibatches(overlap):
  """Two ranks, 4 envs each, local permutations, NCCL gradient all-reduce (bucketed and
  overlapped with backward, or one blocking call) + all-reduced advantage moments + ragged
  micro-batches: after 4 optimiser steps both replicas hold bit-equal parameters, and they equal
  (<= 2e-6 absolute, float32 summation order) the parameters of ONE process that consumed, at
  every step, the union of the two ranks' minibatch rows with a global mean — the semantics the
  reference's single-process Trainer.step has."""
  if torch.cuda.device_count() < 2:
    pytest.skip("needs 2 GPUs (gpurun --gpus 2)")
  import torch.multiprocessing as mp
  manager = mp.Manager()
  results = manager.dict()
  mp.spawn(_worker, args=(2, _free_port(), results, overlap), nprocs=2, join=True)
  ok, replicas_equal, err, loss_err, moved = results[0]
  print(f"\n[dp2 overlap={overlap}] max |param diff| {err:.2e}, max |loss diff| {loss_err:.2e}, "
        f"parameters moved by {moved:.2e}")
  assert replicas_equal and results[1][1], "replicas diverged"
  assert ok, (err, loss_err, moved)
