"""Pivot an `ncu --metrics ... --csv` log into one row per kernel launch.
usage: python tools/ncu_metrics_table.py file.csv"""
import csv
import re
import sys
from collections import OrderedDict

lines = [l for l in open(sys.argv[1], newline="") if not l.startswith("==")]
reader = csv.reader(lines)
hdr = next(reader)
idx = {h: i for i, h in enumerate(hdr)}
rows = OrderedDict()
for r in reader:
  if len(r) < len(hdr):
    continue
  key = r[idx["ID"]]
  name = re.sub(r"\(.*", "", r[idx["Kernel Name"]])
  name = re.sub(r"^void ", "", name).replace("derl::<unnamed>::", "derl::")
  entry = rows.setdefault(key, {"name": name[:78]})
  entry[r[idx["Metric Name"]]] = (r[idx["Metric Value"]], r[idx["Metric Unit"]])
short = {"gpu__time_duration.sum": "time", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active": "tensor%",
         "sm__throughput.avg.pct_of_peak_sustained_elapsed": "sm%", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed": "dram%",
         "dram__bytes_read.sum": "rd", "dram__bytes_write.sum": "wr", "sm__inst_executed_pipe_tensor.sum": "tensor_inst"}
total = 0.0
for e in rows.values():
  v, u = e.get("gpu__time_duration.sum", ("0", "ns"))
  total += float(v.replace(",", "")) * {"ns": 1e-6, "us": 1e-3, "ms": 1.0}.get(u, 1e-6)
print(f"# {len(rows)} launches, {total:.3f} ms total (serialised)")
for e in rows.values():
  cells = []
  for m, s in short.items():
    if m in e:
      v, u = e[m]
      cells.append(f"{s}={v}{u if u not in ('%', '') else ''}")
  print(f"{e['name']:78s} " + " ".join(cells))
