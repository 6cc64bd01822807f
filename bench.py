#!/usr/bin/env python
"""bench.py — PPO update throughput of the B200-native data path (and its CPU reference arm).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...   # the reference's CPU path (port)

A *step* is one pass of the hot path over one synthetic rollout (BASELINE.json configs[2],
weak-scaled: 4096 envs x 128 steps of 84x84x4 uint8 frame stacks PER GPU; at 8 GPUs this is
configs[3]'s 32768 envs):  bootstrap value -> GAE kernel -> 4 epochs x 4 minibatches of
[permutation gather -> advantage normalisation -> NatureCNN forward -> fused PPO loss
fwd+bwd -> network backward -> (NCCL gradient all-reduce) -> clip_grad_norm -> Adam].
Nothing is skipped and the numbers come from the public API (ppo_runner_wrap + PPO.step).

value  = samples consumed by the update (T*N*epochs, all ranks) / device time, rollout
         already resident in HBM when the timed region starts.
e2e    = same, rollout starting in pinned HOST memory (NumPy arrays, as the reference's
         EnvRunner hands them over): the H2D upload of the whole rollout and the D2H of
         bootstrap values + losses are inside the timed region.
roofline = the dominant hand-written kernel (TMA row gather), algorithmic bytes / CUDA-event
         time inside the timed region, against MEASURED_PEAKS.json.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np
import torch

REPO = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, REPO)

OBS_ROW_BYTES = 84 * 84 * 4
# PPO atari defaults of the reference (derl/factory/ppo.py:20-34); epochs x minibatches per
# BASELINE.json configs[0] ("4 epochs x 4 minibatches")
HP = dict(gamma=0.99, lambda_=0.95, cliprange=0.1, value_loss_coef=0.25, entropy_coef=0.01,
          max_grad_norm=0.5, lr=2.5e-4, eps=1e-5)


def parse_args():
  p = argparse.ArgumentParser()
  p.add_argument("--gpus", type=int, default=1)
  p.add_argument("--steps", type=int, default=3)
  p.add_argument("--warmup", type=int, default=3)
  p.add_argument("--impl", choices=("ours", "reference"), default="ours")
  p.add_argument("--envs-per-gpu", type=int, default=4096)
  p.add_argument("--horizon", type=int, default=128)
  p.add_argument("--epochs", type=int, default=4)
  p.add_argument("--minibatches", type=int, default=4)
  p.add_argument("--nactions", type=int, default=4)
  p.add_argument("--micro-batch", type=int, default=65536)   # rows per forward/backward pass
  p.add_argument("--net", choices=("tf32", "fp32", "bf16"), default="tf32",
                 help="tensor-core mode of the cuDNN/cuBLAS policy network (parameters fp32)")
  p.add_argument("--cpu-envs", type=int, default=32,
                 help="envs of the bounded CPU-baseline sample (same horizon/epochs/minibatches)")
  p.add_argument("--no-s2d-hidden", action="store_true",
                 help="A/B switch: keep the 4x4/2 conv strided (cuDNN strided dgrad)")
  p.add_argument("--torch-profile", default=None,
                 help="diagnostic: write a torch.profiler kernel table of one extra step here")
  p.add_argument("--no-e2e", action="store_true")
  p.add_argument("--no-alt", action="store_true",
                 help="skip the informational bf16-autocast-network measurement")
  p.add_argument("--no-cpu-baseline", action="store_true")
  p.add_argument("--no-gae-sweep", action="store_true",
                 help="skip the GAE kernel sweep (BASELINE configs[4]; runs by default at N=1)")
  p.add_argument("--no-small-configs", action="store_true",
                 help="skip BASELINE configs[0] / configs[1] (Atari 8x128, MuJoCo 1x2048)")
  p.add_argument("--alt-steps", type=int, default=5,
                 help="timed steps of each informational alt_* measurement (capped by --steps)")
  return p.parse_args()


def ncu_traffic(kernel_prefix):
  """DRAM bytes per launch of a kernel from the committed `ncu --set full` capture
  (profiles/r01_kernel_traffic.json, written from tools/profile_kernels.py at the same
  131072 x 28224-byte minibatch size), or None."""
  path = os.path.join(REPO, "profiles", "r01_kernel_traffic.json")
  if not os.path.exists(path):
    return None
  with open(path) as f:
    table = json.load(f)
  for name, entry in table.items():
    if kernel_prefix in name:
      return entry["dram_read_bytes"] + entry["dram_write_bytes"]
  return None


def peaks():
  path = os.path.join(REPO, "MEASURED_PEAKS.json")
  if os.path.exists(path):
    with open(path) as f:
      return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
  return 6650.0, "fallback (B200_PROFILING.md)"


# ------------------------------------------------------------------------------ clocks
class ClockSampler:
  """nvidia-smi sampled every 200 ms during the timed region (B200_PROFILING.md recipe)."""
  FIELDS = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,"
            "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
            "clocks_event_reasons.sw_power_cap")

  def __init__(self, index):
    self.index, self.proc, self.lines = index, None, []

  def __enter__(self):
    try:
      self.proc = subprocess.Popen(
          ["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.FIELDS}",
           "--format=csv,noheader,nounits", "-lms", "200"],
          stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
      self.thread = threading.Thread(target=self._read, daemon=True)
      self.thread.start()
    except OSError:
      self.proc = None
    return self

  def _read(self):
    for line in self.proc.stdout:
      self.lines.append(line.strip())

  def __exit__(self, *exc):
    if self.proc is not None:
      time.sleep(0.25)
      self.proc.terminate()
      self.thread.join(timeout=2)

  def summary(self):
    sm, mx, reasons = [], [], set()
    names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
    for line in self.lines:
      parts = [x.strip() for x in line.split(",")]
      if len(parts) != 6:
        continue
      try:
        sm.append(float(parts[0]))
        mx.append(float(parts[1]))
      except ValueError:
        continue
      for name, flag in zip(names, parts[2:]):
        if flag.lower().startswith("active"):
          reasons.add(name)
    if not sm:
      return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
    return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)),
            "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------ CPU arm
def host_rollout(kind, horizon, nenvs, seed=0, nactions=4, obs_dim=17, act_dim=6):
  """Seeded synthetic rollout as NumPy arrays with the dtypes EnvRunner + np.asarray produce
  (SURVEY.md §8a a1, §8d).  Same generator calls as derl_b200.make_rollout(device="cpu"),
  restated here so that the reference arm never imports derl_b200 (nor loads its .so)."""
  import math
  gen = torch.Generator()
  gen.manual_seed(seed)
  lead = (horizon,) if nenvs is None else (horizon, nenvs)
  tail = () if nenvs is None else (nenvs,)
  randn = lambda shape, dtype=torch.float32: torch.randn(shape, generator=gen, dtype=dtype)
  rand = lambda shape: torch.rand(shape, generator=gen)
  if kind == "atari":
    obs = torch.randint(0, 256, lead + (84, 84, 4), generator=gen, dtype=torch.uint8)
    latest = torch.randint(0, 256, tail + (84, 84, 4), generator=gen, dtype=torch.uint8)
    actions = torch.randint(0, nactions, lead, generator=gen, dtype=torch.int64)
    log_prob = -math.log(nactions) + 0.01 * randn(lead)
    rewards = (torch.sign(randn(lead)) * (rand(lead) < 0.1)).to(torch.float64)
    resets = rand(lead) < 0.01
  else:
    obs = randn(lead + (obs_dim,), torch.float64)
    latest = randn(tail + (obs_dim,), torch.float64)
    actions = randn(lead + (act_dim,))
    log_prob = (-0.5 * actions ** 2 - 0.5 * math.log(2 * math.pi)).sum(-1) + 0.01 * randn(lead)
    rewards = randn(lead, torch.float64)
    resets = rand(lead) < 0.001
  values = randn(lead + (1,))
  out = dict(observations=obs, actions=actions, log_prob=log_prob, values=values,
             rewards=rewards, resets=resets)
  out = {k: v.numpy() for k, v in out.items()}
  out["state"] = dict(latest_observations=latest.numpy())
  return out


def find_reference():
  """The live reference tree, looked up as SURVEY.md §8c says: $DERL_REF -> baseline/_ref ->
  /root/reference (none of them exists on the GPU box: the reference is pure Python, `pip
  install` of it drops its sub-packages — setup.py lists packages=["derl"] only — so there is
  no baseline/_ref to ship).  Returns the imported package or None."""
  for cand in (os.environ.get("DERL_REF"), os.path.join(REPO, "baseline", "_ref"),
               "/root/reference"):
    if cand and os.path.isfile(os.path.join(cand, "derl", "runners", "onpolicy.py")):
      for path in (os.path.join(REPO, "tests", "_stubs"), cand):   # gym / atari_py import stubs
        if path not in sys.path:
          sys.path.insert(0, path)
      try:
        import derl
        import derl.summary
        derl.summary.stop_recording()
        return derl
      except Exception as exc:   # noqa: BLE001 — any import problem means "not available here"
        sys.stderr.write(f"[bench] reference at {cand} not importable: {exc!r}\n")
  return None


CPU_CONFIGS = {
    "atari": dict(hp=dict(cliprange=HP["cliprange"], value_loss_coef=HP["value_loss_coef"],
                          entropy_coef=HP["entropy_coef"]), lr=HP["lr"]),
    # derl/factory/ppo.py:37-49 (MuJoCo defaults)
    "mujoco": dict(hp=dict(cliprange=0.2, value_loss_coef=0.25, entropy_coef=0.0), lr=3e-4),
}


class HostArrayRunner:
  """EnvRunner-shaped source yielding the same host rollout forever (for the live reference)."""

  def __init__(self, rollout, policy, nenvs, horizon):
    self.rollout, self.policy, self.horizon, self.nenvs = rollout, policy, horizon, nenvs
    self.env = type("E", (), {"nenvs": nenvs, "unwrapped": property(lambda s: s)})()
    self.nsteps, self.step_count = 10 ** 12, 0

  def is_exhausted(self):
    return False

  def run(self, obs=None):
    while True:
      self.step_count += self.horizon * (self.nenvs or 1)
      yield {k: (dict(v) if k == "state" else np.array(v)) for k, v in self.rollout.items()}


def cpu_update_seconds(kind, nenvs, horizon, epochs, minibatches, steps, warmup, nactions=4,
                       prefer_live=True):
  """Seconds per PPO update of the reference's CPU path on a host rollout of the given shape,
  with all the host threads this process may use.  The reference's own classes when its tree is
  present (kind "live": TransformInteractions[GAE, MergeTimeBatch] -> IterateWithMinibatches ->
  NormalizeAdvantages -> PPOLoss -> Trainer(Adam eps 1e-5, clip .5), derl/runners/onpolicy.py:
  65-75, derl/alg/common.py:66-78), else the oracle port of the same steps (kind "port").
  Returns (mean seconds per update, kind)."""
  torch.set_num_threads(max(1, len(os.sched_getaffinity(0))))  # torchrun exports OMP_NUM_THREADS=1
  cfg = CPU_CONFIGS[kind]
  rollout = host_rollout(kind, horizon, nenvs, seed=0, nactions=nactions)
  ref = find_reference() if prefer_live else None
  torch.manual_seed(0)
  np.random.seed(0)
  times = []
  if ref is not None:
    model = ref.NatureCNNModel([nactions, 1]) if kind == "atari" else ref.MuJoCoModel(17, [6, 1])
    model.to("cpu")
    policy = ref.ActorCriticPolicy(model)
    runner = ref.ppo_runner_wrap(HostArrayRunner(rollout, policy, nenvs, horizon),
                                 num_epochs=epochs, num_minibatches=minibatches)
    optimizer = torch.optim.Adam(model.parameters(), lr=cfg["lr"], eps=HP["eps"])
    alg = ref.PPO(runner, ref.Trainer(optimizer, max_grad_norm=HP["max_grad_norm"]), **cfg["hp"])
    it = runner.run()
    for i in range(warmup + steps):
      t0 = time.perf_counter()
      for _ in range(epochs * minibatches):
        alg.step(next(it))
      if i >= warmup:
        times.append(time.perf_counter() - t0)
    return sum(times) / len(times), "live"
  from oracle import derl_oracle as O
  model = O.NatureCNN(nactions) if kind == "atari" else O.MuJoCoMLP(17, 6)
  optimizer = torch.optim.Adam(model.parameters(), lr=cfg["lr"], eps=HP["eps"])
  latest = torch.from_numpy(rollout["state"]["latest_observations"])
  columns = {k: v for k, v in rollout.items() if k != "state"}
  for i in range(warmup + steps):
    t0 = time.perf_counter()
    with torch.no_grad():
      last_value = model(latest if nenvs is not None else latest[None])[-1].numpy()
    if nenvs is None:
      last_value = last_value[0]
    O.ppo_update(model, optimizer, columns, last_value, gamma=HP["gamma"], lambda_=HP["lambda_"],
                 num_epochs=epochs, num_minibatches=minibatches,
                 max_grad_norm=HP["max_grad_norm"], batched=nenvs is not None, **cfg["hp"])
    if i >= warmup:
      times.append(time.perf_counter() - t0)
  return sum(times) / len(times), "port"


def cpu_update_throughput(args, nenvs, steps, warmup):
  sec, kind = cpu_update_seconds("atari", nenvs, args.horizon, args.epochs, args.minibatches,
                                 steps, warmup, args.nactions)
  return args.horizon * nenvs * args.epochs / sec, sec, kind


# cpu_baseline.kind in the bench contract: "reference" = the reference's own code, "port" = oracle
CPU_KIND = {"live": "reference", "port": "port"}


def cpu_sample_text(args, sec, kind):
  what = {"live": "the reference's own classes (derl.ppo_runner_wrap + PPO/Trainer) on CPU",
          "port": "oracle port of the reference path (oracle/derl_oracle.py) on CPU"}[kind]
  return (f"{args.cpu_envs} envs x {args.horizon} steps (a 1/{args.envs_per_gpu // args.cpu_envs} "
          f"env-axis sample of the {args.envs_per_gpu}-env workload: per-sample cost of the CPU "
          f"path does not depend on the env count), {args.epochs} epochs x {args.minibatches} "
          f"minibatches, NatureCNN float32, full update, {sec:.2f} s per step; {what}")


def run_reference(args, rank):
  """The reference's CPU implementation of the path, on a bounded sample, rank 0 only.  Never
  imports derl_b200 (no CUDA library is loaded in this arm)."""
  if rank != 0:
    return
  value, sec, kind = cpu_update_throughput(args, args.cpu_envs, args.steps, args.warmup)
  cores = torch.get_num_threads()
  config = workload_config(args, args.gpus)
  config.update({
      "workload": (f"atari-shaped PPO update on the HOST CPU, bounded sample: {args.cpu_envs} envs "
                   f"x {args.horizon} steps of the {args.envs_per_gpu}-envs/GPU workload, 84x84x4 "
                   f"u8 obs, {args.epochs} epochs x {args.minibatches} minibatches, NatureCNN "
                   f"A={args.nactions}, float32 (BASELINE configs[2] shape per sample)"),
      "envs_total": args.cpu_envs, "micro_batch": None, "network": "fp32 (CPU)",
      "l2": "n/a (CPU)", "parallelism": f"{cores} host threads"})
  line = {
      "impl": "reference", "metric": "ppo_update_samples_per_sec", "value": value,
      "unit": "samples/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
      "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak",
      "vs_baseline": None, "dtype": "f32 (GAE f64 registers)", "data": "synthetic",
      "config": config,
      "cpu_baseline": {"value": value, "unit": "samples/s", "cores": cores,
                       "kind": CPU_KIND[kind], "source": kind,
                       "sample": cpu_sample_text(args, sec, kind)},
      "e2e": {"value": value, "unit": "samples/s", "h2d_bytes_per_step": 0,
              "d2h_bytes_per_step": 0},
  }
  print(json.dumps(line), flush=True)


def workload_config(args, world):
  return {"workload": (f"atari-shaped PPO update, {args.envs_per_gpu} envs/GPU x {args.horizon} "
                       f"steps, 84x84x4 u8 obs, {args.epochs} epochs x {args.minibatches} "
                       f"minibatches, NatureCNN A={args.nactions} (BASELINE configs[2]; "
                       f"{args.envs_per_gpu * world} envs total)"),
          "envs_total": args.envs_per_gpu * world, "horizon": args.horizon,
          "epochs": args.epochs, "minibatches": args.minibatches,
          "micro_batch": args.micro_batch, "network": args.net,
          "l2": "inputs >> L2 (14.8 GB rollout per GPU; every minibatch reads fresh rows)",
          "parallelism": f"env-axis dp{world}"}


# ------------------------------------------------------------------------------ our arm
def build_alg(args, d, source, world, device, graphed=False):
  model = source.policy.model
  sync = None
  group_norm = None
  if world > 1:
    from derl_b200 import parallel
    sync = parallel.GradientAllReduce(model)
    group_norm = torch.distributed.group.WORLD
  transforms = [d.GAE(source.policy, gamma=HP["gamma"], lambda_=HP["lambda_"], normalize=False),
                d.MergeTimeBatch()]
  runner = d.TransformInteractions(source, transforms)
  runner = d.IterateWithMinibatches(runner, args.epochs, args.minibatches)
  runner = d.TransformInteractions(runner, [d.NormalizeAdvantages(group=group_norm)])
  if graphed:   # CUDA-graph replay of the micro-batched step, frames fed by index (fused gather)
    runner.runner.fused_gather = True
    anneal_t = d.LinearAnneal(HP["lr"], 10e6 * 100, name="lr", device=device)
    optimizer = torch.optim.Adam(model.parameters(), lr=anneal_t.get_tensor(), eps=HP["eps"],
                                 capturable=True)
    trainer = d.GraphedTrainer(optimizer, anneals=[anneal_t], max_grad_norm=HP["max_grad_norm"],
                               grad_sync=sync, micro_batch=args.micro_batch)
    alg = d.PPO(runner, trainer, cliprange=HP["cliprange"],
                value_loss_coef=HP["value_loss_coef"], entropy_coef=HP["entropy_coef"])
    return alg, runner
  optimizer = torch.optim.Adam(model.parameters(), lr=HP["lr"], eps=HP["eps"], fused=True)
  anneal = d.LinearAnneal(HP["lr"], 10e6 * 100, name="lr")

  class GroupLR:  # LinearAnneal -> optimizer.param_groups (host float; closed-form step_to)
    name = "lr"

    def step_to(self, n):
      anneal.step_to(n)
      for group in optimizer.param_groups:
        group["lr"] = float(anneal.get_tensor())

    def summarize(self, n):
      pass

  trainer = d.Trainer(optimizer, anneals=[GroupLR()], max_grad_norm=HP["max_grad_norm"],
                      grad_sync=sync, micro_batch=args.micro_batch)
  alg = d.PPO(runner, trainer, cliprange=HP["cliprange"],
              value_loss_coef=HP["value_loss_coef"], entropy_coef=HP["entropy_coef"])
  return alg, runner


class HostRolloutSource:
  """EnvRunner-like source whose arrays live in pinned host memory (NumPy views)."""

  def __init__(self, policy, device_rollout, nenvs, horizon):
    self.policy, self.horizon, self.nenvs = policy, horizon, nenvs
    self.env = type("E", (), {"nenvs": nenvs, "unwrapped": property(lambda s: s)})()
    self.nsteps, self.step_count = None, 0
    self.host = {}
    self.bytes = 0
    for key, val in device_rollout.items():
      if key == "state":
        continue
      pinned = torch.empty(val.shape, dtype=val.dtype, pin_memory=True)
      pinned.copy_(val)
      self.host[key] = pinned.numpy()
      self.bytes += pinned.numel() * pinned.element_size()
    latest = device_rollout["state"]["latest_observations"]
    pinned = torch.empty(latest.shape, dtype=latest.dtype, pin_memory=True)
    pinned.copy_(latest)
    self.latest = pinned.numpy()
    self.bytes += pinned.numel() * pinned.element_size()

  def is_exhausted(self):
    return False

  def run(self, obs=None):
    while True:
      self.step_count += self.horizon * self.nenvs
      out = dict(self.host)
      out["state"] = dict(latest_observations=self.latest)
      yield out


def concurrent_h2d_gbps(host_array, device, world, nbytes=2 << 30):
  """GB/s of a pinned-host -> device copy issued by every rank at the same moment."""
  src = torch.from_numpy(host_array).reshape(-1).view(torch.uint8)[:nbytes]
  dst = torch.empty(src.numel(), dtype=torch.uint8, device=device)
  dst.copy_(src[:dst.numel()], non_blocking=True)   # warm-up
  torch.cuda.synchronize()
  if world > 1:
    torch.distributed.barrier()
  start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  start.record()
  dst.copy_(src, non_blocking=True)
  stop.record()
  torch.cuda.synchronize()
  gbps = src.numel() / start.elapsed_time(stop) / 1e6
  if world > 1:
    t = torch.tensor([gbps], device=device)
    torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MIN)
    gbps = float(t)
  return gbps


def one_update(alg, runner_iter, nbatches):
  losses = []
  for _ in range(nbatches):
    losses.append(alg.step(next(runner_iter)))
  return losses


def timed_updates(alg, runner, nbatches, steps, warmup, world, read_losses):
  """W untimed + K timed steps, barrier + synchronize on both sides, CUDA events, max over
  ranks.  Returns (seconds for K steps, losses of the last step)."""
  it = runner.run()
  losses = None
  for _ in range(warmup):
    losses = one_update(alg, it, nbatches)
  if world > 1:
    torch.distributed.barrier()
  torch.cuda.synchronize()
  start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  start.record()
  for _ in range(steps):
    losses = one_update(alg, it, nbatches)
    if read_losses:  # D2H read of the step's result inside the timed region
      losses = torch.stack([l.detach() for l in losses]).cpu()
  stop.record()
  torch.cuda.synchronize()
  if world > 1:
    torch.distributed.barrier()
  sec = start.elapsed_time(stop) / 1e3
  if world > 1:
    t = torch.tensor([sec], device="cuda")
    torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
    sec = float(t)
  return sec, losses


def kernel_times(profile):
  """name -> (launches, mean ms) from the (name, start, end) event triples."""
  out = {}
  for name, start, end in profile:
    out.setdefault(name, []).append(start.elapsed_time(end))
  return {k: (len(v), float(np.mean(v))) for k, v in out.items()}


def gae_sweep(d, hbm_peak):
  """BASELINE configs[4]: GAE kernel alone, T x N sweep, 17 B/element, L2 flushed between
  launches by a 256 MB memset (inputs of the large sizes exceed L2 anyway)."""
  K = torch.ops.derl_b200
  flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
  rows = []
  for nsteps in (128, 512, 2048):
    for nenvs in (8, 512, 4096, 32768, 65536):
      if nsteps * nenvs > (1 << 27):
        continue
      gen = torch.Generator(device="cuda").manual_seed(1)
      rewards = torch.randn(nsteps, nenvs, device="cuda", generator=gen)
      values = torch.randn(nsteps, nenvs, device="cuda", generator=gen)
      resets = torch.rand(nsteps, nenvs, device="cuda", generator=gen) < 0.01
      last = torch.randn(nenvs, device="cuda", generator=gen)
      for variant, vname in ((1, "direct"), (2, "tma")):
        if variant == 2 and (nenvs % 16 or nenvs < 32):
          continue
        times = []
        for it in range(8):
          flush.zero_()
          s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
          s.record()
          K.gae(rewards, values, resets, last, 0.99, 0.95, False, variant)
          e.record()
          torch.cuda.synchronize()
          if it >= 3:
            times.append(s.elapsed_time(e))
        ms = float(np.median(times))
        gbs = (17.0 * nsteps * nenvs + 4 * nenvs) / ms / 1e6
        rows.append({"T": nsteps, "N": nenvs, "variant": vname, "ms": ms, "GBps": gbs,
                     "frac": gbs / hbm_peak})
  return rows


# BASELINE configs[0] / configs[1]: the reference's own default shapes, through the public API
SMALL_CONFIGS = {
    "C1_atari_8x128": dict(kind="atari", nenvs=8, horizon=128, epochs=4, minibatches=4),
    "C2_mujoco_1x2048": dict(kind="mujoco", nenvs=None, horizon=2048, epochs=10, minibatches=32),
}


def small_config_gpu(d, cfg, mode, updates, warmup):
  """ms per PPO update of one small config.  mode: "eager" = ppo_runner_wrap + PPO + Trainer
  stepped minibatch by minibatch; "graphed" = the same with GraphedTrainer (CUDA-graph replay);
  "learn" = what a drop-in user runs, `PPO.learn()` with the defaults (which picks the fused
  whole-update kernel when the model qualifies)."""
  kind = cfg["kind"]
  hp, lr = CPU_CONFIGS[kind]["hp"], CPU_CONFIGS[kind]["lr"]
  torch.manual_seed(0)
  model = d.NatureCNNModel([4, 1]) if kind == "atari" else d.MuJoCoModel(17, [6, 1])
  policy = d.ActorCriticPolicy(model)
  per_update = cfg["epochs"] * cfg["minibatches"]
  rollout_steps = cfg["horizon"] * (cfg["nenvs"] or 1)
  source = d.SyntheticRolloutRunner(policy, kind, cfg["nenvs"], cfg["horizon"],
                                    nsteps=None, device="cuda", seed=1)
  runner = d.ppo_runner_wrap(source, num_epochs=cfg["epochs"], num_minibatches=cfg["minibatches"])
  if mode == "graphed":
    anneal = d.LinearAnneal(lr, 1e9, device="cuda", name="lr")
    opt = torch.optim.Adam(model.parameters(), lr=anneal.get_tensor(), eps=HP["eps"],
                           capturable=True)
    trainer = d.GraphedTrainer(opt, anneals=[anneal], max_grad_norm=HP["max_grad_norm"])
  else:
    opt = torch.optim.Adam(model.parameters(), lr=lr, eps=HP["eps"], fused=mode == "eager")
    trainer = d.Trainer(opt, max_grad_norm=HP["max_grad_norm"])
  alg = d.PPO(runner, trainer, **hp)
  np.random.seed(0)
  start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  if mode == "learn":
    def run(n):
      source.nsteps = source.step_count + n * rollout_steps
      alg.learn(progress=False)
    run(warmup)
    torch.cuda.synchronize()
    start.record()
    run(updates)
    stop.record()
    torch.cuda.synchronize()
    return start.elapsed_time(stop) / updates, None
  it = runner.run()
  for _ in range(warmup * per_update):
    alg.step(next(it))
  torch.cuda.synchronize()
  start.record()
  for _ in range(updates * per_update):
    loss = alg.step(next(it))
  stop.record()
  torch.cuda.synchronize()
  return start.elapsed_time(stop) / updates, float(loss)


def small_configs(d):
  out = {}
  for name, cfg in SMALL_CONFIGS.items():
    samples = cfg["horizon"] * (cfg["nenvs"] or 1) * cfg["epochs"]
    entry = {"optimizer_steps_per_update": cfg["epochs"] * cfg["minibatches"],
             "samples_per_update": samples, "unit": "samples/s"}
    for mode in ("eager", "graphed", "learn"):
      ms, loss = small_config_gpu(d, cfg, mode, updates=5, warmup=3)
      entry[mode] = {"ms_per_update": ms, "value": samples / ms * 1e3}
    sec, kind = cpu_update_seconds(cfg["kind"], cfg["nenvs"], cfg["horizon"], cfg["epochs"],
                                   cfg["minibatches"], steps=2, warmup=1)
    entry["cpu"] = {"ms_per_update": sec * 1e3, "value": samples / sec, "kind": CPU_KIND[kind],
                    "source": kind, "cores": torch.get_num_threads()}
    out[name] = entry
  return out


def gae_sweep_cpu():
  """The reference's GAE loop (NumPy, one core: derl/runners/trajectory_transforms.py:45-65,
  restated in oracle/derl_oracle.py:gae) on a few points of the same sweep, float64 rewards as
  its EnvRunner produces them (21 B per element).  Bounded: ~3 s in total."""
  from oracle import derl_oracle as O
  rows = []
  rng = np.random.RandomState(0)
  for nsteps, nenvs in ((128, 8), (128, 4096), (128, 65536), (2048, 4096)):
    rewards = rng.standard_normal((nsteps, nenvs))
    values = rng.standard_normal((nsteps, nenvs, 1)).astype(np.float32)
    resets = rng.rand(nsteps, nenvs) < 0.01
    last = rng.standard_normal((nenvs, 1)).astype(np.float32)
    best = float("inf")
    for _ in range(2):
      t0 = time.perf_counter()
      O.gae(rewards, values, resets, last, 0.99, 0.95, normalize=False)
      best = min(best, time.perf_counter() - t0)
    rows.append({"T": nsteps, "N": nenvs, "ms": best * 1e3,
                 "GBps": 21.0 * nsteps * nenvs / best / 1e9, "cores": 1})
  return rows


def run_ours(args, rank, world, local):
  import derl_b200 as d
  from derl_b200 import _lib, ops
  d.summary.stop_recording()
  device = torch.device("cuda", local)
  torch.cuda.set_device(device)
  torch.backends.cudnn.benchmark = True
  tf32 = args.net in ("tf32", "bf16")
  torch.backends.cudnn.allow_tf32 = tf32
  torch.backends.cuda.matmul.allow_tf32 = tf32

  if args.no_s2d_hidden:
    d.NatureCNNBase.space_to_depth_hidden = False
  torch.manual_seed(0)  # identical initial weights on every rank
  model = d.NatureCNNModel([args.nactions, 1])
  if args.net == "bf16":
    model.autocast_dtype = torch.bfloat16
  policy = d.ActorCriticPolicy(model)
  nenvs, horizon = args.envs_per_gpu, args.horizon
  source = d.SyntheticRolloutRunner(policy, "atari", nenvs, horizon, nsteps=None, device=device,
                                    seed=1000 + rank, nactions=args.nactions)
  np.random.seed(1234 + rank)  # local permutations per shard
  alg, runner = build_alg(args, d, source, world, device)
  nbatches = args.epochs * args.minibatches
  samples_per_step = horizon * nenvs * args.epochs * world
  hbm_peak, peak_kind = peaks()

  # ---- value: rollout resident in HBM
  sec0, _ = timed_updates(alg, runner, nbatches, 0, args.warmup, world, False)  # warm-up only
  ops.PROFILE = []
  launches0 = _lib.launch_count()
  with ClockSampler(local) as clocks:
    sec, losses = timed_updates(alg, runner, nbatches, args.steps, 0, world, False)
  launches = _lib.launch_count() - launches0
  ktimes = kernel_times(ops.PROFILE)
  ops.PROFILE = None
  value = samples_per_step * args.steps / sec
  last_loss = float(losses[-1])

  # ---- roofline of the dominant hand-written kernel (TMA row gather)
  mb_rows = horizon * nenvs // args.minibatches
  roofline, kernels = None, {}
  if "gather_rows" in ktimes:
    n, ms = ktimes["gather_rows"]
    bytes_per_launch = (8 + 2 * OBS_ROW_BYTES) * mb_rows
    achieved = bytes_per_launch / ms / 1e6
    traffic = ncu_traffic("gather_rows_tma_kernel") if (nenvs, horizon, args.minibatches) == \
        (4096, 128, 4) else None
    roofline = {"bound": "hbm", "kernel": "gather_rows_tma_kernel", "achieved": achieved,
                "peak": hbm_peak, "unit": "GB/s", "frac": achieved / hbm_peak, "traffic": traffic,
                "peak_source": peak_kind, "bytes_per_launch": bytes_per_launch,
                "launch_ms": ms, "launches_timed": n,
                "note": "dominant kernel of the north-star data path (GAE -> gather -> loss); by "
                        "time the largest derl_b200 kernels of the whole update are the network "
                        "stem's (SURVEY 8f rank 2): see kernels.stem_backward / stem_conv_relu "
                        "for their algorithmic GB/s and fraction of the same peak"}
  per_elem = {"gae": 17.0 * horizon * nenvs + 4 * nenvs,
              "frames_to_s2d": 5.0 * OBS_ROW_BYTES * min(args.micro_batch, mb_rows),
              "ppo_loss_categorical": (8 * args.nactions + 32) * min(args.micro_batch, mb_rows),
              "gather_columns": (8 + 2 * (8 + 4 + 4 + 4 + 4 + 4 + 1)) * mb_rows,
              "normalize": 8.0 * mb_rows}
  if args.net == "tf32":   # stem kernels per micro-batch chunk (algorithmic bytes)
    chunk = min(args.micro_batch, mb_rows)
    # K6t: frame in, fp32 activation + 1-bit ReLU mask out; K7t: frame + gradient + mask in
    per_elem["stem_conv_relu"] = (OBS_ROW_BYTES + 400 * 32 * 4.0 + 1600) * chunk
    per_elem["stem_backward"] = (OBS_ROW_BYTES + 400 * 32 * 4.0 + 1600) * chunk
    # K5 serves the 64x9x9 and 64x7x7 activations in turn: mean bytes of the two launches
    per_elem["relu_bwd_bias"] = 12.0 * chunk * 64 * (81 + 49) / 2
  for name, (n, ms) in ktimes.items():
    entry = {"launches": n, "mean_ms": ms}
    if name in per_elem:
      entry["GBps"] = per_elem[name] / ms / 1e6
      entry["frac"] = entry["GBps"] / hbm_peak
    kernels[name] = entry

  # the derl_b200 kernel that takes the most time in the step (a network-stem kernel, not one of
  # the three north-star kernels): its roofline object, same definition as `roofline`
  top = None
  timed_names = [n for n in kernels if "GBps" in kernels[n]]
  if timed_names:
    name = max(timed_names, key=lambda n: kernels[n]["launches"] * kernels[n]["mean_ms"])
    k = kernels[name]
    top = {"bound": "hbm", "kernel": name, "achieved": k["GBps"], "peak": hbm_peak, "unit": "GB/s",
           "frac": k["frac"], "traffic": None, "bytes_per_launch": per_elem[name],
           "launch_ms": k["mean_ms"], "launches_timed": k["launches"],
           "share_of_step": k["launches"] * k["mean_ms"] / (sec * 1e3)}

  # the north-star data path alone (K1 + K2 + K3 and their helpers, no network): aggregate
  # algorithmic bytes / summed CUDA-event time of those launches inside the timed region
  data_path = None
  names = [n for n in ("gae", "gather_rows", "gather_columns", "normalize", "ppo_loss_categorical")
           if n in ktimes]
  if names:
    bytes_of = dict(per_elem, gather_rows=float((8 + 2 * OBS_ROW_BYTES) * mb_rows))
    tot_ms = sum(ktimes[n][0] * ktimes[n][1] for n in names) / args.steps
    tot_bytes = sum(ktimes[n][0] * bytes_of[n] for n in names) / args.steps
    data_path = {"kernels": names, "ms_per_step": tot_ms, "bytes_per_step": tot_bytes,
                 "GBps": tot_bytes / tot_ms / 1e6, "frac": tot_bytes / tot_ms / 1e6 / hbm_peak,
                 "share_of_step": tot_ms / (sec / args.steps * 1e3)}

  if args.torch_profile and rank == 0:   # diagnostic only; never part of a reported number
    from torch.profiler import ProfilerActivity, profile
    it = runner.run()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
      one_update(alg, it, nbatches)
      torch.cuda.synchronize()
    with open(args.torch_profile, "w") as f:
      f.write(prof.key_averages().table(sort_by="cuda_time_total", row_limit=45,
                                        max_name_column_width=90))

  alt_steps = max(1, min(args.steps, args.alt_steps))
  # ---- the same update at the REFERENCE's arithmetic: float32 cuDNN/cuBLAS (TF32 off, which
  # also routes the stem through K4 + cuDNN instead of the INT8 kernels K6/K7).  The
  # precision-matched companion of `value` (derl/models.py:102-124 is float32 end to end).
  alt_fp32 = None
  if args.net != "fp32" and not args.no_alt:
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    # cuDNN's float32 engines are faster at 32768-row passes (0.47 vs 0.37 M samples/s at 65536)
    micro_fp32 = min(args.micro_batch, 32768)
    alg.trainer.micro_batch = micro_fp32
    sec_p, losses_p = timed_updates(alg, runner, nbatches, alt_steps, 2, world, False)
    alg.trainer.micro_batch = args.micro_batch
    torch.backends.cudnn.allow_tf32 = tf32
    torch.backends.cuda.matmul.allow_tf32 = tf32
    alt_fp32 = {"network": "fp32 (allow_tf32 off: cuDNN/cuBLAS float32, no INT8 stem kernels)",
                "dtype": "f32; GAE f64 registers; u8 gather",
                "value": samples_per_step * alt_steps / sec_p, "unit": "samples/s",
                "ms_per_step": sec_p / alt_steps * 1e3, "steps": alt_steps, "warmup": 2,
                "micro_batch": micro_fp32, "last_loss": float(losses_p[-1])}

  # ---- informational: the same update with the network under bf16 autocast
  alt = None
  if args.net != "bf16" and not args.no_alt:
    model.autocast_dtype = torch.bfloat16
    sec_a, _ = timed_updates(alg, runner, nbatches, alt_steps, 2, world, False)
    model.autocast_dtype = None
    alt = {"network": "bf16 autocast (parameters fp32)",
           "value": samples_per_step * alt_steps / sec_a, "unit": "samples/s",
           "ms_per_step": sec_a / alt_steps * 1e3, "steps": alt_steps}

  # ---- informational: minibatch gather fused into the stem kernels: the
  # observations of a minibatch are never materialised, K6/K7 read the rollout rows in place.
  # Not the headline: the headline keeps the reference's materialised minibatch (and K2, the
  # kernel the roofline object is about, inside the timed region).
  fused = None
  if not args.no_alt:
    runner.runner.fused_gather = True
    sec_f, _ = timed_updates(alg, runner, nbatches, alt_steps, 2, world, False)
    runner.runner.fused_gather = False
    fused = {"what": "IterateWithMinibatches(fused_gather=True): stem kernels gather rows in place",
             "value": samples_per_step * alt_steps / sec_f, "unit": "samples/s",
             "ms_per_step": sec_f / alt_steps * 1e3, "steps": alt_steps}

  # ---- informational: the same update as CUDA graphs (GraphedTrainer with micro-batches: one
  # graph per micro-batch chunk + one update graph; frames reach the stem kernels by index)
  graphed = None
  if not args.no_alt and args.net == "tf32":
    alg_g, runner_g = build_alg(args, d, source, world, device, graphed=True)
    sec_g, _ = timed_updates(alg_g, runner_g, nbatches, alt_steps, 4, world, False)
    graphed = {"what": "GraphedTrainer(micro_batch) + fused_gather: CUDA-graph replay of the step",
               "value": samples_per_step * alt_steps / sec_g, "unit": "samples/s",
               "ms_per_step": sec_g / alt_steps * 1e3, "steps": alt_steps,
               "graph_replays": alg_g.trainer.replays}
    del alg_g, runner_g
    torch.cuda.empty_cache()

  # ---- e2e: rollout in pinned host memory, uploaded inside the timed region
  e2e = None
  import psutil
  host_need = world * (horizon + 1) * nenvs * (OBS_ROW_BYTES + 32)
  if not args.no_e2e and psutil.virtual_memory().available < 2 * host_need:
    e2e = {"value": None, "unit": "samples/s", "skipped": "not enough free host RAM for the "
           f"pinned rollouts ({host_need >> 30} GiB needed)"}
  elif not args.no_e2e:
    host_source = HostRolloutSource(policy, source.rollout(), nenvs, horizon)
    source._cached = None
    torch.cuda.empty_cache()
    alg_h, runner_h = build_alg(args, d, host_source, world, device)
    sec_h, _ = timed_updates(alg_h, runner_h, nbatches, args.steps, 1, world, True)
    perm_bytes = args.epochs * horizon * nenvs * 8
    # what the host can feed: every rank copies 2 GiB pinned -> device at the same time (all
    # GPUs of a box pull through the same host memory system), min over ranks
    h2d_gbps = concurrent_h2d_gbps(host_source.host["observations"], device, world)
    h2d_total = world * (host_source.bytes + perm_bytes)
    e2e = {"value": samples_per_step * args.steps / sec_h, "unit": "samples/s",
           "h2d_bytes_per_step": h2d_total,
           "d2h_bytes_per_step": world * (nenvs * 4 + nbatches * 4),
           "ms_per_step": sec_h / args.steps * 1e3,
           "h2d_GBps_per_gpu_concurrent": h2d_gbps,
           "h2d_floor_ms_per_step": h2d_total / world / h2d_gbps / 1e6,
           "note": "the upload is overlapped with the first epoch (HostColumn); e2e - value is "
                   "bounded below by (h2d floor - one epoch of compute)"}
    del host_source, alg_h, runner_h

  # free the big rollouts before the side measurements
  del alg, runner
  source._cached = None
  torch.cuda.empty_cache()
  # BASELINE configs[4] (the "GAE HBM GB/s (frac of peak)" half of the metric) and configs[0..1]
  sweep = gae_sweep(d, hbm_peak) if (rank == 0 and world == 1 and not args.no_gae_sweep) else None
  small = small_configs(d) if (rank == 0 and world == 1 and not args.no_small_configs) else None

  cpu = None
  if rank == 0 and world == 1 and not args.no_cpu_baseline:
    v, sec_c, kind = cpu_update_throughput(args, args.cpu_envs, 1, 1)
    cpu = {"value": v, "unit": "samples/s", "cores": torch.get_num_threads(),
           "kind": CPU_KIND[kind], "source": kind,
           "sample": cpu_sample_text(args, sec_c, kind) + ", 1 warm-up + 1 timed"}

  if rank == 0:
    line = {
        "metric": "ppo_update_samples_per_sec", "value": value, "unit": "samples/s",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": sec / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None,
        "dtype": {"tf32": "f32 params, tf32 tensor-core conv/GEMM; GAE f64 registers; u8 gather",
                  "fp32": "f32; GAE f64 registers; u8 gather",
                  "bf16": "f32 params, bf16 autocast network; GAE f64 registers; u8 gather"}[args.net],
        "data": "synthetic", "config": workload_config(args, world),
        "clocks": clocks.summary(), "e2e": e2e, "gpu_launches": launches,
        "roofline": roofline, "roofline_top_by_time": top, "data_path": data_path,
        "kernels": kernels, "cpu_baseline": cpu,
        "alt_fp32": alt_fp32,
        "alt_network": alt, "alt_fused_gather": fused, "alt_graphed": graphed,
        "last_loss": last_loss,
    }
    if sweep is not None:
      line["gae_sweep"] = sweep
      line["gae_sweep_cpu"] = gae_sweep_cpu()
    if small is not None:
      line["small_configs"] = small
    print(json.dumps(line), flush=True)


def main():
  args = parse_args()
  rank = int(os.environ.get("RANK", "0"))
  world = int(os.environ.get("WORLD_SIZE", "1"))
  local = int(os.environ.get("LOCAL_RANK", "0"))
  if args.impl == "reference":
    run_reference(args, rank)
    return
  if world > 1:
    from derl_b200 import parallel
    # NCCL announces its version on stdout when the first communicator comes up; stdout is
    # reserved for the one JSON line, so point fd 1 at stderr until that has happened
    sys.stdout.flush()
    saved = os.dup(1)
    os.dup2(2, 1)
    try:
      parallel.init_from_env("nccl")
      torch.distributed.all_reduce(torch.zeros(1, device=f"cuda:{local}"))
      torch.cuda.synchronize()
    finally:
      sys.stdout.flush()
      os.dup2(saved, 1)
      os.close(saved)
  run_ours(args, rank, world, local)
  if world > 1:
    torch.distributed.destroy_process_group()


if __name__ == "__main__":
  main()
