"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: time per kernel name
over the LAST step of the run (everything after the second-to-last gae kernel launch).
usage: python tools/ncu_launches.py gpurun_out/launches.csv [n_last_gae=1]"""
import csv
import re
import sys
from collections import defaultdict

path = sys.argv[1]
rows = []
with open(path, newline="") as f:
  lines = [l for l in f if not l.startswith("==")]
reader = csv.reader(lines)
hdr = next(reader)
idx = {h: i for i, h in enumerate(hdr)}
for r in reader:
  if len(r) < len(hdr) or r[idx["Metric Name"]] != "gpu__time_duration.sum":
    continue
  val = float(r[idx["Metric Value"]].replace(",", ""))
  unit = r[idx["Metric Unit"]]
  scale = {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(unit, 1e-6)
  rows.append((r[idx["Kernel Name"]], val * scale))
gae = [i for i, (n, _) in enumerate(rows) if "gae_" in n]
start = gae[-1] if gae else 0
step = rows[start:]
tot = sum(t for _, t in step)
agg = defaultdict(lambda: [0, 0.0])
for n, t in step:
  short = n.replace("(anonymous namespace)::", "").replace("<unnamed>::", "").replace("void ", "")
  short = re.sub(r"<.*", "", short)
  short = re.sub(r"\(.*", "", short)[:70]
  agg[short][0] += 1
  agg[short][1] += t
print(f"launches in file {len(rows)}; last step: {len(step)} launches, {tot:.2f} ms (serialised, cold)")
for n, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:40]:
  print(f"{t:9.3f} ms {100 * t / tot:5.1f}% {c:5d}x  {n}")
