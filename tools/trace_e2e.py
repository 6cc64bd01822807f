"""Timeline of one `e2e` bench step (rollout in pinned host memory): where the milliseconds between
the resident `value` and `e2e` go.  CUDA events are recorded on the side stream around every
upload+gather launch of `HostColumn` and on the main stream around every `alg.step`; all times are
printed relative to the start of the traced step, with the host-side time of the same moments.

    python tools/trace_e2e.py [--steps 2] [--upload-ctas 8] [bench.py workload flags]
"""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402


def main():
  steps, upload_ctas = 2, None
  if "--steps" in sys.argv:
    i = sys.argv.index("--steps")
    steps = int(sys.argv[i + 1])
    del sys.argv[i:i + 2]
  if "--upload-ctas" in sys.argv:   # CTAs of the upload+gather kernel (HostColumn default: 8)
    i = sys.argv.index("--upload-ctas")
    upload_ctas = int(sys.argv[i + 1])
    del sys.argv[i:i + 2]
  sys.argv = sys.argv[:1] + ["--steps", "1", "--warmup", "1"] + sys.argv[1:]
  args = bench.parse_args()
  import derl_b200 as d
  from derl_b200.runners import host_column
  d.summary.stop_recording()
  device = torch.device("cuda", 0)
  torch.cuda.set_device(device)
  torch.backends.cudnn.benchmark = True
  torch.backends.cudnn.allow_tf32 = True          # the bench's default arithmetic (--net tf32)
  torch.backends.cuda.matmul.allow_tf32 = True
  torch.manual_seed(0)
  model = d.NatureCNNModel([args.nactions, 1])
  policy = d.ActorCriticPolicy(model)
  nenvs, horizon = args.envs_per_gpu, args.horizon
  source = d.SyntheticRolloutRunner(policy, "atari", nenvs, horizon, nsteps=None, device=device,
                                    seed=1000, nactions=args.nactions)
  np.random.seed(1234)
  host_source = bench.HostRolloutSource(policy, source.rollout(), nenvs, horizon)
  source._cached = None
  torch.cuda.empty_cache()
  alg, runner = bench.build_alg(args, d, host_source, 1, device)
  nbatches = args.epochs * args.minibatches

  trace = []          # (label, start event, end event, host t at enqueue)
  issue = host_column.HostColumn._issue

  def traced_issue(self, perm, start, count):
    if upload_ctas is not None:
      self.max_ctas = upload_ctas
    with torch.cuda.stream(self._stream):
      s = torch.cuda.Event(enable_timing=True)
      s.record()
    t = time.perf_counter()
    out = issue(self, perm, start, count)
    with torch.cuda.stream(self._stream):
      e = torch.cuda.Event(enable_timing=True)
      e.record()
    trace.append((f"upload rows [{start}, {start + count})", s, e, t))
    return out

  host_column.HostColumn._issue = traced_issue
  it = runner.run()
  for step in range(1 + steps):
    trace.clear()
    torch.cuda.synchronize()
    origin = torch.cuda.Event(enable_timing=True)
    origin.record()
    t0 = time.perf_counter()
    losses = []
    for j in range(nbatches):
      th = time.perf_counter()
      batch = next(it)
      s = torch.cuda.Event(enable_timing=True)
      s.record()
      tn = time.perf_counter()
      losses.append(alg.step(batch))
      e = torch.cuda.Event(enable_timing=True)
      e.record()
      trace.append((f"step {j:2d} (next() took {1e3 * (tn - th):6.1f} ms on the host)", s, e, th))
    out = torch.stack([l.detach() for l in losses]).cpu()
    end = torch.cuda.Event(enable_timing=True)
    end.record()
    torch.cuda.synchronize()
    if step == 0:
      continue
    print(f"--- traced step {step}: {origin.elapsed_time(end):.1f} ms on the device, "
          f"{1e3 * (time.perf_counter() - t0):.1f} ms on the host, last loss {float(out[-1]):.5f}")
    rows = sorted(((origin.elapsed_time(s), origin.elapsed_time(e), label, 1e3 * (t - t0))
                   for label, s, e, t in trace), key=lambda r: r[0])
    for a, b, label, th in rows:
      print(f"  {a:8.1f} -> {b:8.1f} ms  ({b - a:6.1f})  host@{th:7.1f}  {label}")


if __name__ == "__main__":
  main()
