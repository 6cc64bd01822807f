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
  """`Trainer(grad_sync=GradientAllReduce(model))`: flat-buffer gradient averaging, overlapped
  with the tail of the backward pass.

  All parameter gradients are views into ONE contiguous buffer, laid out in REVERSE registration
  order — the order in which backward finishes them — and cut into `buckets` contiguous ranges.
  While armed (the `Trainer` arms it before the backward that completes the minibatch's
  gradient: the only one, or the last micro-batch's) a bucket is all-reduced on a side stream
  the moment its last gradient has been accumulated, so the big early-finishing layers (for
  NatureCNN the 3136x512 linear is 95 % of the 6.75 MB) travel over NVLink while the conv
  layers are still in backward; `__call__` (reference insertion point: between
  derl/alg/common.py:70 and :71) launches whatever is left and makes the compute stream wait.
  The average is taken inside the collective (ReduceOp.AVG) — no separate division pass.
  """

  def __init__(self, model, group=None, bucket_bytes=1 << 20, overlap=True):
    self.group = group
    self.world = dist.get_world_size(group) if dist.is_initialized() else 1
    params = [p for p in model.parameters() if p.requires_grad]
    order = list(reversed(params))               # approximately the order backward finishes them
    total = sum(p.numel() for p in params)
    ref = params[0]
    self.flat = torch.zeros(total, dtype=ref.dtype, device=ref.device)
    self.buckets = []                            # [start, stop, number of params, pending]
    self._bucket_of = {}
    offset, start, count = 0, 0, 0
    for p in order:
      # same strides as the parameter (channels_last conv weights stay channels_last): autograd's
      # gradient-layout contract and the fused optimizers require grad.layout == param.layout
      p.grad = torch.as_strided(self.flat, p.size(), p.stride(), storage_offset=offset)
      self._bucket_of[id(p)] = len(self.buckets)
      offset += p.numel()
      count += 1
      if (offset - start) * self.flat.element_size() >= bucket_bytes:
        self.buckets.append([start, offset, count, count])
        start, count = offset, 0
    if count:
      self.buckets.append([start, offset, count, count])
    self.params = params
    self.overlap = bool(overlap) and self.flat.is_cuda
    self.armed = False
    self._launched = []
    self._stream = None
    self._op = dist.ReduceOp.AVG if (self.flat.is_cuda and dist.is_initialized()
                                     and dist.get_backend(group) == "nccl") else dist.ReduceOp.SUM
    if self.overlap and self.world > 1:
      self._stream = torch.cuda.Stream(device=self.flat.device)
      for p in params:
        p.register_post_accumulate_grad_hook(self._on_grad)

  # ------------------------------------------------------------------ overlap machinery
  def arm(self):
    """The next backward completes the minibatch's gradients: reduce buckets as they finish."""
    if self._stream is None:
      return
    self.armed = True
    self._launched = [False] * len(self.buckets)
    for bucket in self.buckets:
      bucket[3] = bucket[2]

  def _reduce(self, index, stream):
    start, stop = self.buckets[index][:2]
    view = self.flat[start:stop]
    with torch.cuda.stream(stream):
      dist.all_reduce(view, op=self._op, group=self.group)
      if self._op == dist.ReduceOp.SUM:
        view.div_(self.world)

  def _on_grad(self, param):
    if not self.armed:
      return
    index = self._bucket_of[id(param)]
    bucket = self.buckets[index]
    bucket[3] -= 1
    if bucket[3] == 0 and not self._launched[index]:
      self._launched[index] = True
      self._stream.wait_stream(torch.cuda.current_stream(self.flat.device))
      self._reduce(index, self._stream)

  def __call__(self, model=None):
    if self.world == 1:
      return
    for p in self.params:  # zero_grad(set_to_none=True) may have detached a view
      if p.grad is None or p.grad.data_ptr() < self.flat.data_ptr() or \
          p.grad.data_ptr() >= self.flat.data_ptr() + self.flat.numel() * self.flat.element_size():
        raise RuntimeError("gradient left the flat buffer: call optimizer.zero_grad("
                           "set_to_none=False) when using GradientAllReduce")
    if self._stream is None or not self.armed:
      dist.all_reduce(self.flat, op=self._op, group=self.group)
      if self._op == dist.ReduceOp.SUM:
        self.flat.div_(self.world)
      return
    current = torch.cuda.current_stream(self.flat.device)
    for index, done in enumerate(self._launched):   # parameters without a gradient this step
      if not done:
        self._launched[index] = True
        self._stream.wait_stream(current)
        self._reduce(index, self._stream)
    current.wait_stream(self._stream)
    self.armed = False
