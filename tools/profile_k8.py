"""One K8 launch (MuJoCo-shaped whole update, 2048 samples x 10 epochs x 32 minibatches) for ncu."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import derl_b200  # noqa: E402,F401

K = torch.ops.derl_b200
gen = torch.Generator(device="cuda").manual_seed(0)
size, odim, adim = 2048, 17, 6
shapes = [(64, odim), (64,), (64, 64), (64,), (adim, 64), (adim,),
          (64, odim), (64,), (64, 64), (64,), (1, 64), (1,), (adim,)]
params = [torch.randn(sh, device="cuda", generator=gen) * .1 for sh in shapes]
m1 = [torch.zeros_like(p) for p in params]
m2 = [torch.zeros_like(p) for p in params]
obs = torch.randn(size, odim, device="cuda", generator=gen, dtype=torch.float64)
act = torch.randn(size, adim, device="cuda", generator=gen)
col = lambda: torch.randn(size, device="cuda", generator=gen)
lp, adv, vt, val = col() * .1 - 8, col(), col(), col()
perm = torch.cat([torch.randperm(size, device="cuda", generator=gen) for _ in range(10)])
for _ in range(2):
  s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  s.record()
  K.ppo_mlp_update(params, m1, m2, obs, act, lp, adv, vt, val, perm, 10, 64, True, 1e-8, .2, .25, 0.,
                   .5, 3e-4, .9, .999, 1e-5, 0)
  e.record()
  torch.cuda.synchronize()
  print(f"K8: {s.elapsed_time(e):.3f} ms per update (320 steps)")
