"""Small default configs of the reference (BASELINE.json configs[0] and configs[1]) through the
same public API as bench.py, eager `Trainer` vs CUDA-graphed `GraphedTrainer`, next to the
oracle's CPU port of the reference path on the host cores.  One JSON line per config.

  configs[0]  Atari-shaped 8 envs x 128 steps, 4 epochs x 4 minibatches (256), NatureCNN
  configs[1]  MuJoCo-shaped 1 env x 2048 steps (unbatched), 10 epochs x 32 minibatches (64), MLP 64x64
"""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import derl_b200 as d  # noqa: E402
from oracle import derl_oracle as O  # noqa: E402

CONFIGS = {
    "atari_8x128": dict(kind="atari", nenvs=8, horizon=128, epochs=4, minibatches=4,
                        hp=dict(cliprange=.1, value_loss_coef=.25, entropy_coef=.01), lr=2.5e-4),
    "mujoco_1x2048": dict(kind="mujoco", nenvs=None, horizon=2048, epochs=10, minibatches=32,
                          hp=dict(cliprange=.2, value_loss_coef=.25, entropy_coef=0.), lr=3e-4),
}


def make_model(kind):
  torch.manual_seed(0)
  return d.NatureCNNModel([4, 1]) if kind == "atari" else d.MuJoCoModel(17, [6, 1])


def gpu_updates_per_sec(cfg, graphed, updates=6, warmup=3):
  model = make_model(cfg["kind"])
  policy = d.ActorCriticPolicy(model)
  source = d.SyntheticRolloutRunner(policy, cfg["kind"], cfg["nenvs"], cfg["horizon"],
                                    nsteps=None, device="cuda", seed=1)
  runner = d.ppo_runner_wrap(source, num_epochs=cfg["epochs"], num_minibatches=cfg["minibatches"])
  if graphed:
    lr = d.LinearAnneal(cfg["lr"], 1e9, device="cuda", name="lr")
    opt = torch.optim.Adam(model.parameters(), lr=lr.get_tensor(), eps=1e-5, capturable=True)
    trainer = d.GraphedTrainer(opt, anneals=[lr], max_grad_norm=.5)
  else:
    opt = torch.optim.Adam(model.parameters(), lr=cfg["lr"], eps=1e-5, fused=True)
    trainer = d.Trainer(opt, max_grad_norm=.5)
  alg = d.PPO(runner, trainer, **cfg["hp"])
  it = runner.run()
  per_update = cfg["epochs"] * cfg["minibatches"]
  np.random.seed(0)
  for _ in range(warmup * per_update):
    alg.step(next(it))
  torch.cuda.synchronize()
  start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  start.record()
  for _ in range(updates * per_update):
    loss = alg.step(next(it))
  stop.record()
  torch.cuda.synchronize()
  sec = start.elapsed_time(stop) / 1e3 / updates
  return sec, float(loss), getattr(trainer, "replays", 0)


def cpu_update_sec(cfg, updates=2):
  torch.manual_seed(0)
  model = O.NatureCNN(4) if cfg["kind"] == "atari" else O.MuJoCoMLP(17, 6)
  opt = torch.optim.Adam(model.parameters(), lr=cfg["lr"], eps=1e-5)
  rollout = d.make_rollout(cfg["kind"], cfg["horizon"], cfg["nenvs"], device="cpu", seed=1)
  latest = torch.from_numpy(rollout["state"]["latest_observations"])
  cols = {k: v for k, v in rollout.items() if k != "state"}
  np.random.seed(0)
  times = []
  for i in range(updates + 1):
    t0 = time.perf_counter()
    with torch.no_grad():
      last = model(latest if cfg["kind"] == "atari" else latest[None])[-1].numpy()
    if cfg["kind"] == "mujoco":
      last = last[0]
    O.ppo_update(model, opt, cols, last, num_epochs=cfg["epochs"],
                 num_minibatches=cfg["minibatches"], batched=cfg["kind"] == "atari", **cfg["hp"])
    if i:
      times.append(time.perf_counter() - t0)
  return float(np.mean(times))


def main():
  d.summary.stop_recording()
  torch.backends.cudnn.benchmark = True
  for name, cfg in CONFIGS.items():
    samples = cfg["horizon"] * (cfg["nenvs"] or 1) * cfg["epochs"]
    eager, loss_e, _ = gpu_updates_per_sec(cfg, graphed=False)
    graphed, loss_g, replays = gpu_updates_per_sec(cfg, graphed=True)
    cpu = cpu_update_sec(cfg)
    print(json.dumps({
        "config": name, "metric": "ppo_update_samples_per_sec",
        "eager": {"value": samples / eager, "ms_per_update": eager * 1e3, "last_loss": loss_e},
        "graphed": {"value": samples / graphed, "ms_per_update": graphed * 1e3,
                    "last_loss": loss_g, "graph_replays": replays},
        "cpu_port": {"value": samples / cpu, "ms_per_update": cpu * 1e3,
                     "cores": torch.get_num_threads()},
        "optimizer_steps_per_update": cfg["epochs"] * cfg["minibatches"]}), flush=True)


if __name__ == "__main__":
  main()
