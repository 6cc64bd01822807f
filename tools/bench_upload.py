"""PCIe microbenchmark for the first-epoch upload path: derl_b200::gather_rows_upload (TMA bulk
reads straight from pinned host memory, rows mirrored into the resident copy) against a plain
pinned cudaMemcpy of the same bytes, for several grid sizes."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import derl_b200  # noqa: E402,F401

K = torch.ops.derl_b200
rows = int(os.environ.get("ROWS", 131072))
host = torch.empty((rows, 84, 84, 4), dtype=torch.uint8, pin_memory=True)
host.random_(0, 256)
resident = torch.empty(host.shape, dtype=torch.uint8, device="cuda")
perm = torch.from_numpy(np.random.RandomState(0).permutation(rows)).cuda()
nbytes = host.numel()


def timed(fn, reps=3):
  best = 1e9
  for _ in range(reps):
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    fn()
    e.record()
    torch.cuda.synchronize()
    best = min(best, s.elapsed_time(e))
  return best


ms = timed(lambda: resident.copy_(host, non_blocking=True))
print(f"cudaMemcpy pinned H2D          {ms:8.2f} ms  {nbytes / ms / 1e6:7.1f} GB/s", flush=True)
for ctas in (2, 4, 8, 16, 32, 148):
  ms = timed(lambda: K.gather_rows_upload(host.data_ptr(), perm, 0, rows, resident, ctas))
  print(f"gather_rows_upload max_ctas={ctas:3d} {ms:8.2f} ms  {nbytes / ms / 1e6:7.1f} GB/s", flush=True)
out = K.gather_rows_upload(host.data_ptr(), perm, 0, rows, resident, 8)
torch.cuda.synchronize()
assert torch.equal(resident.cpu(), host) and torch.equal(out.cpu(), host[perm.cpu()])
print("upload gather verified bit-exact")
