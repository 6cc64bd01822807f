"""CUDA-graphed training step.

The reference's `Trainer.step` (derl/alg/common.py:66-78) issues, per minibatch, ~60 forward
and ~80 backward ATen ops plus clipping and Adam from Python; for the small default configs
(Atari 8 envs x 128 steps -> minibatches of 256; MuJoCo 2048 steps -> minibatches of 64) the
GPU finishes each kernel long before Python has launched the next one.  `GraphedTrainer`
captures the whole step — policy forward, fused PPO loss, backward, clip_grad_norm_,
optimizer step — once per minibatch signature into CUDA graphs and replays them: per
minibatch the host issues one multi-tensor copy into the static inputs and one or two graph
launches.  Same arithmetic, same kernels, same order as the eager `Trainer`.

Micro-batched minibatches (`micro_batch=`, the 131 072-row minibatches of the large Atari
configuration) replay one graph per row chunk (forward + loss x its share of the minibatch +
backward, accumulating into the static gradient buffers) and one update graph.  The chunk
graph's inputs are the narrow columns (24 B per row, copied) and the OBSERVATIONS BY INDEX: with
`IterateWithMinibatches(fused_gather=True)` a chunk's observations are a `RowSelection`, so the
static input is its int64 row vector (8 B per row) and the stem kernels gather the frames in
place — no frame is copied to feed a graph.  A dense observation tensor is copied like any
other input (925 MB per 32 768-frame chunk: correct, but use the fused gather).

Requirements: a capturable optimizer (`torch.optim.Adam(..., capturable=True)`; the learning
rate may be a device tensor, e.g. `LinearAnneal(..., device="cuda").get_tensor()`, which
`step_to` updates in place between replays).  With `grad_sync` (NCCL all-reduce) the step
is split into two graphs around the eager collective.
"""
import torch

from .. import summary
from ..runners.row_selection import RowSelection
from .common import Trainer

# what PPOLoss reads from a minibatch (everything else is not copied into the graph inputs)
LOSS_KEYS = ("observations", "actions", "log_prob", "advantages", "value_targets", "values")


class _Captured:
  def __init__(self):
    self.inputs, self.loss = None, None
    self.forward_backward, self.update = None, None


class GraphedTrainer(Trainer):
  """Drop-in `Trainer` that replays CUDA graphs after `warmup` eager steps per signature."""

  def __init__(self, optimizer, anneals=None, max_grad_norm=None, grad_sync=None, warmup=3,
               keys=LOSS_KEYS, micro_batch=None):
    super().__init__(optimizer, anneals=anneals, max_grad_norm=max_grad_norm,
                     grad_sync=grad_sync, micro_batch=micro_batch)
    self._chunk_graphs = {}
    self._update_graph = None
    self._total = None
    self._grads = None
    self.warmup = warmup
    self.keys = keys
    self._seen = {}
    self._graphs = {}
    self.replays = 0

  @staticmethod
  def _signature(data, keys):
    return tuple((k, tuple(data[k].shape), data[k].dtype) for k in keys if k in data)

  def _params(self, alg):
    return [p for p in alg.model.parameters() if p.requires_grad]

  def _capture(self, alg, sig, data):
    """Record (not run) the step on static input buffers shaped like `data`."""
    cap = _Captured()
    cap.inputs = {k: torch.empty_like(data[k]) for k, _, _ in sig}
    for k, t in cap.inputs.items():
      t.copy_(data[k])
    params = self._params(alg)
    keep = self.grad_sync is not None
    pool = torch.cuda.graph_pool_handle()
    # grads are re-created inside the capture (from the graph's private pool) unless a flat
    # all-reduce buffer owns them, in which case they are zeroed in place inside the graph
    self.optimizer.zero_grad(set_to_none=not keep)
    count = alg.loss_fn.call_count
    cap.forward_backward = torch.cuda.CUDAGraph()
    with torch.cuda.graph(cap.forward_backward, pool=pool):
      if keep:
        self.optimizer.zero_grad(set_to_none=False)
      cap.loss = alg.loss(cap.inputs)
      cap.loss.backward()
      if self.grad_sync is None:
        self._update(params)
    if self.grad_sync is not None:
      cap.update = torch.cuda.CUDAGraph()
      with torch.cuda.graph(cap.update, pool=pool):
        self._update(params)
    alg.loss_fn.call_count = count   # tracing is not a call; replays count below
    return cap

  def _update(self, params):
    if self.max_grad_norm is not None:
      torch.nn.utils.clip_grad_norm_(params, self.max_grad_norm, foreach=True)
    self.optimizer.step()

  # ------------------------------------------------------------------ micro-batched replay
  @staticmethod
  def _chunk_signature(chunk, keys, weight):
    sig = []
    for k in keys:
      if k not in chunk:
        continue
      v = chunk[k]
      if isinstance(v, RowSelection):
        sig.append((k, "rows", tuple(v.shape), v.source.data_ptr()))
      else:
        sig.append((k, tuple(v.shape), v.dtype))
    return tuple(sig) + (round(float(weight), 12),)

  def _capture_chunk(self, alg, chunk, weight):
    """Record forward + weighted loss + backward of one row chunk on static inputs; gradients
    accumulate into the parameters' existing (static) .grad buffers."""
    cap = _Captured()
    cap.inputs = {}
    for k in self.keys:
      if k not in chunk:
        continue
      v = chunk[k]
      if isinstance(v, RowSelection):
        rows = v.rows.clone()                     # the graph's input: indices, not frames
        cap.inputs[k] = RowSelection(v.source, rows, 0, rows.numel())
      else:
        cap.inputs[k] = v.clone()
    count = alg.loss_fn.call_count
    cap.forward_backward = torch.cuda.CUDAGraph()
    with torch.cuda.graph(cap.forward_backward, pool=self._pool):
      part = alg.loss(cap.inputs) * weight
      part.backward()
      self._total.add_(part.detach())
    alg.loss_fn.call_count = count
    return cap

  @staticmethod
  def _load_chunk(cap, chunk):
    dst, src = [], []
    for k, static in cap.inputs.items():
      v = chunk[k]
      if isinstance(static, RowSelection):
        dst.append(static.perm)
        src.append(v.rows)
      else:
        dst.append(static)
        src.append(v)
    torch._foreach_copy_(dst, src)

  def _step_chunked(self, alg, data):
    from .common import _split_rows
    params = self._params(alg)
    if self._total is None:
      self._pool = torch.cuda.graph_pool_handle()
      self._total = torch.zeros((), dtype=torch.float32, device=params[0].device)
    chunks = list(_split_rows(data, self.micro_batch))
    sigs = [self._chunk_signature(c, self.keys, w) for c, w in chunks]
    # eager warm-up is per SHAPE (optimizer state, library plans); a new rollout tensor (another
    # source pointer in the signature) only needs a new capture
    shapes = [tuple(x[:3] if isinstance(x, tuple) and len(x) == 4 and x[1] == "rows" else x
                    for x in sig) for sig in sigs]
    seen = min(self._seen.get(shape, 0) for shape in shapes)
    for shape in set(shapes):
      self._seen[shape] = self._seen.get(shape, 0) + 1
    if seen < self.warmup:
      return Trainer.step(self, alg, data)      # eager: optimizer state, library plans, grads
    for anneal in self.anneals:
      anneal.step_to(alg.runner.step_count)
    if self._grads is None:                      # static gradient buffers the graphs write to
      self._grads = [p.grad if p.grad is not None else torch.zeros_like(p) for p in params]
    for p, g in zip(params, self._grads):        # an eager step in between may have dropped them
      if p.grad is not g:
        p.grad = g
    torch._foreach_zero_(self._grads)
    self._total.zero_()
    for (chunk, weight), sig in zip(chunks, sigs):
      cap = self._chunk_graphs.get(sig)
      if cap is None:
        cap = self._chunk_graphs[sig] = self._capture_chunk(alg, chunk, weight)
        # capture only records: load the same inputs and run it
      self._load_chunk(cap, chunk)
      cap.forward_backward.replay()
      alg.loss_fn.call_count += 1
    if self.grad_sync is not None:
      self.grad_sync(alg.model)
    if self._update_graph is None:
      self._update_graph = torch.cuda.CUDAGraph()
      with torch.cuda.graph(self._update_graph, pool=self._pool):
        self._update(params)
    self._update_graph.replay()
    self.replays += 1
    self.step_count += 1
    return self._total.clone()

  def step(self, alg, data):
    obs = data.get("observations")
    tensors_ok = all(isinstance(data.get(k), (torch.Tensor, RowSelection)) and data[k].is_cuda
                     for k in self.keys if k in data)
    if summary.should_record() or not tensors_ok:
      return super().step(alg, data)     # logging steps and host data take the eager path
    if self.micro_batch and obs is not None and obs.shape[0] > self.micro_batch:
      return self._step_chunked(alg, data)
    if isinstance(obs, RowSelection):
      data = dict(data, observations=obs.materialize())
    sig = self._signature(data, self.keys)
    seen = self._seen.get(sig, 0)
    self._seen[sig] = seen + 1
    if seen < self.warmup:
      return super().step(alg, data)     # initialises optimizer state and library plans
    for anneal in self.anneals:          # in-place update of the lr tensor the graph reads
      anneal.step_to(alg.runner.step_count)
    cap = self._graphs.get(sig)
    if cap is None:
      cap = self._graphs[sig] = self._capture(alg, sig, data)
    else:
      torch._foreach_copy_([cap.inputs[k] for k, _, _ in sig], [data[k] for k, _, _ in sig])
    cap.forward_backward.replay()
    if cap.update is not None:
      self.grad_sync(alg.model)
      cap.update.replay()
    self.replays += 1
    alg.loss_fn.call_count += 1
    self.step_count += 1
    return cap.loss.detach().clone()
