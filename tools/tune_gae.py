"""Times every GAE TMA tile configuration (DERL_GAE_TMA_CFG) against the direct variant and
checks that all of them produce identical bits.  L2 is flushed between launches."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import derl_b200  # noqa: E402,F401

K = torch.ops.derl_b200
flush = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")


def bench(fn, reps=7):
  times = []
  for it in range(reps):
    flush.zero_()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    out = fn()
    e.record()
    torch.cuda.synchronize()
    if it >= 2:
      times.append(s.elapsed_time(e))
  return float(np.median(times)), out


shapes = [(2048, 65536), (512, 65536), (128, 65536), (2048, 32768), (128, 32768), (128, 4096),
          (2048, 4096)]
cfgs = [int(c) for c in os.environ.get("CFGS", "0,1,2,5").split(",")]
for nsteps, nenvs in shapes:
  gen = torch.Generator(device="cuda").manual_seed(1)
  r = torch.randn(nsteps, nenvs, device="cuda", generator=gen)
  v = torch.randn(nsteps, nenvs, device="cuda", generator=gen)
  z = torch.rand(nsteps, nenvs, device="cuda", generator=gen) < 0.01
  lv = torch.randn(nenvs, device="cuda", generator=gen)
  nbytes = 17.0 * nsteps * nenvs + 4 * nenvs
  os.environ.pop("DERL_GAE_TMA_CFG", None)
  ms, ref = bench(lambda: K.gae(r, v, z, lv, .99, .95, False, 1))
  line = [f"T{nsteps} N{nenvs}: direct {nbytes / ms / 1e6:6.0f}"]
  for cfg in cfgs:
    os.environ["DERL_GAE_TMA_CFG"] = str(cfg)
    ms, out = bench(lambda: K.gae(r, v, z, lv, .99, .95, False, 2))
    same = torch.equal(out[0], ref[0]) and torch.equal(out[1], ref[1])
    line.append(f"c{cfg} {nbytes / ms / 1e6:6.0f}{'' if same else ' MISMATCH'}")
  print(" | ".join(line), flush=True)
