"""One NatureCNN forward + fused PPO loss + backward on a 32768-frame micro-batch, for an ncu
capture of the network's library kernels (tensor-pipe utilisation):

  python tools/profile_network.py                                  # plain run first
  ncu --set full --clock-control none --profile-from-start off -o gpurun_out/network \
      python tools/profile_network.py

The profiled region (cudaProfilerStart/Stop) is the 4th iteration, after cuDNN autotuning.
NET=bf16 switches the network to bf16 autocast.
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import derl_b200 as d  # noqa: E402


def main():
  d.summary.stop_recording()
  torch.backends.cudnn.benchmark = True
  torch.backends.cudnn.allow_tf32 = True
  torch.backends.cuda.matmul.allow_tf32 = True
  nb = int(os.environ.get("ROWS", 32768))
  torch.manual_seed(0)
  model = d.NatureCNNModel([4, 1])
  if os.environ.get("NET") == "bf16":
    model.autocast_dtype = torch.bfloat16
  loss_fn = d.PPOLoss(d.ActorCriticPolicy(model), cliprange=0.1)
  gen = torch.Generator(device="cuda").manual_seed(0)
  batch = dict(
      observations=torch.randint(0, 256, (nb, 84, 84, 4), device="cuda", dtype=torch.uint8,
                                 generator=gen),
      actions=torch.randint(0, 4, (nb,), device="cuda", generator=gen),
      log_prob=torch.randn(nb, device="cuda", generator=gen) * .05 - 1.39,
      advantages=torch.randn(nb, device="cuda", generator=gen),
      value_targets=torch.randn(nb, 1, device="cuda", generator=gen),
      values=torch.randn(nb, 1, device="cuda", generator=gen))
  for it in range(4):
    if it == 3:
      torch.cuda.synchronize()
      torch.cuda.cudart().cudaProfilerStart()
    model.zero_grad(set_to_none=True)
    loss_fn(batch).backward()
    if it == 3:
      torch.cuda.synchronize()
      torch.cuda.cudart().cudaProfilerStop()
  print("ok", float(loss_fn.last_stats[0]))


if __name__ == "__main__":
  main()
