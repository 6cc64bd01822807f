"""Env-axis data parallelism for the PPO update: one process per GPU, NCCL over NVLink.

The reference has no distributed code at all; the path shards naturally
along the environment axis (§8e): GAE is independent per env, samples are independent in
the loss.  Couplings are means only, so the exchange steps are
  (1) one all-reduce of the flat gradient buffer after each minibatch backward
      (reference insertion point: between derl/alg/common.py:70 and :71), averaged over
      ranks, so clip-by-global-norm and Adam then act identically on every rank;
  (2) a 3-double all-reduce of {sum, sumsq, count} for advantage normalisation
      (NormalizeAdvantages(group=...)), so shards normalise with global statistics.
There is no data-path collective: observations never leave the GPU that owns their envs.
"""
import os

import torch
import torch.distributed as dist


def init_from_env(backend=None):
  """Initialise torch.distributed from torchrun's env (RANK/WORLD_SIZE/LOCAL_RANK/MASTER_*).
  Returns (rank, world_size, local_rank); a no-op single-process answer without torchrun."""
  world = int(os.environ.get("WORLD_SIZE", "1"))
  rank = int(os.environ.get("RANK", "0"))
  local = int(os.environ.get("LOCAL_RANK", "0"))
  if world > 1 and not dist.is_initialized():
    if backend is None:
      backend = "nccl" if torch.cuda.is_available() else "gloo"
    if backend == "nccl":
      torch.cuda.set_device(local)
      dist.init_process_group(backend, device_id=torch.device("cuda", local))
    else:
      dist.init_process_group(backend)
  return rank, world, local


def env_shard(nenvs, rank, world):
  """[lo, hi) slice of the env axis owned by `rank` (equal shards; nenvs % world == 0)."""
  if nenvs % world != 0:
    raise ValueError(f"nenvs={nenvs} must be divisible by world size {world} so that every "
                     "rank takes equal minibatches (mean-of-means == global mean)")
  per = nenvs // world
  return rank * per, (rank + 1) * per


def shard_rollout(rollout, rank, world, env_axis=1):
  """Slice every [T, N, ...] array of a rollout dict (and state.latest_observations [N, ...])
  down to this rank's envs."""
  out = {}
  for key, val in rollout.items():
    if key == "state":
      state = dict(val)
      obs = state.get("latest_observations")
      if obs is not None:
        lo, hi = env_shard(obs.shape[0], rank, world)
        state["latest_observations"] = obs[lo:hi]
      out[key] = state
    else:
      lo, hi = env_shard(val.shape[env_axis], rank, world)
      index = [slice(None)] * val.ndim
      index[env_axis] = slice(lo, hi)
      out[key] = val[tuple(index)]
  return out


class GradientAllReduce:
  """`Trainer(grad_sync=GradientAllReduce(model))`: flat-buffer gradient averaging.

  All parameter gradients are views into ONE contiguous buffer, so each minibatch issues a
  single all-reduce (6.75 MB for NatureCNN: latency-bound, NVLS in-switch reduction when
  NCCL selects it) on the compute stream right after backward.
  """

  def __init__(self, model, group=None):
    self.group = group
    self.world = dist.get_world_size(group) if dist.is_initialized() else 1
    params = [p for p in model.parameters() if p.requires_grad]
    total = sum(p.numel() for p in params)
    ref = params[0]
    self.flat = torch.zeros(total, dtype=ref.dtype, device=ref.device)
    offset = 0
    for p in params:
      # same strides as the parameter (channels_last conv weights stay channels_last): autograd's
      # gradient-layout contract and the fused optimizers require grad.layout == param.layout
      p.grad = torch.as_strided(self.flat, p.size(), p.stride(), storage_offset=offset)
      offset += p.numel()
    self.params = params

  def __call__(self, model=None):
    if self.world == 1:
      return
    for p in self.params:  # zero_grad(set_to_none=True) may have detached a view
      if p.grad is None or p.grad.data_ptr() < self.flat.data_ptr() or \
          p.grad.data_ptr() >= self.flat.data_ptr() + self.flat.numel() * self.flat.element_size():
        raise RuntimeError("gradient left the flat buffer: call optimizer.zero_grad("
                           "set_to_none=False) when using GradientAllReduce")
    dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=self.group)
    self.flat.div_(self.world)
