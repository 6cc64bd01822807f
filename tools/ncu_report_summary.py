"""Summarise an `ncu --set full` report (read here, no GPU needed) into a table and a traffic JSON:

  python tools/ncu_report_summary.py gpurun_out/kernels.ncu-rep|raw.csv profiles/r02_kernels_ncu.txt \\
         profiles/r02_kernel_traffic.json

One row per (kernel, grid): the LAST captured launch of each — duration, DRAM bytes read / written,
DRAM and tensor-pipe utilisation, registers, shared memory, the top warp-stall reasons.
"""
import csv
import io
import json
import re
import subprocess
import sys

METRICS = {
    "gpu__time_duration.sum": "duration",
    "dram__bytes_read.sum": "rd",
    "dram__bytes_write.sum": "wr",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed": "dram%",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active": "tensor%",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed": "sm%",
    "sm__warps_active.avg.pct_of_peak_sustained_active": "warps%",
    "launch__registers_per_thread": "regs",
    "launch__shared_mem_per_block_dynamic": "smem_dyn",
    "launch__grid_size": "grid",
    "launch__block_size": "block",
}
SCALE = {"ns": 1e-6, "us": 1e-3, "usecond": 1e-3, "ms": 1.0, "msecond": 1.0, "nsecond": 1e-6,
         "byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}


def main():
  rep, table_path, json_path = sys.argv[1:4]
  if rep.endswith(".csv"):   # already exported on the GPU box: ncu -i x.ncu-rep --page raw --csv
    raw = open(rep).read()
  else:
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True,
                         text=True, check=True).stdout
  rows = list(csv.reader(io.StringIO(raw)))
  hdr, units = rows[0], rows[1]
  col = {h: i for i, h in enumerate(hdr)}
  stall_cols = [(h, i) for i, h in enumerate(hdr)
                if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("per_issue_active.ratio")]
  entries = {}
  for r in rows[2:]:
    name = re.sub(r"\(.*", "", r[col["Kernel Name"]]).replace("void ", "")
    name = name.replace("derl::<unnamed>::", "derl::").replace("<unnamed>::", "derl::")
    e = {"name": name}
    for metric, short in METRICS.items():
      if metric in col:
        val = float(r[col[metric]].replace(",", "") or 0)
        unit = units[col[metric]].split("/")[0]
        e[short] = val * SCALE.get(unit, 1.0)
    stalls = sorted(((float(r[i].replace(",", "") or 0),
                      h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", ""))
                     for h, i in stall_cols), reverse=True)[:3]
    e["stalls"] = ", ".join(f"{n} {v:.2f}" for v, n in stalls)
    entries[f"{name}|grid={int(e.get('grid', 0))}"] = e
  with open(table_path, "w") as f:
    f.write(f"# from {rep}: ncu --set full --clock-control none, last captured launch per (kernel, grid); "
            "times are cold-cache and serialised\n")
    for key, e in entries.items():
      f.write(f"{e['name'][:70]:70s} grid {int(e.get('grid', 0)):5d} x {int(e.get('block', 0)):4d}  "
              f"{e.get('duration', 0):8.4f} ms  rd {e.get('rd', 0) / 1e6:9.2f} MB  wr {e.get('wr', 0) / 1e6:9.2f} MB  "
              f"dram {e.get('dram%', 0):5.1f}%  tensor {e.get('tensor%', 0):5.1f}%  sm {e.get('sm%', 0):5.1f}%  "
              f"regs {int(e.get('regs', 0)):3d}  smem {e.get('smem_dyn', 0) / 1024:6.1f} KB  stalls: {e['stalls']}\n")
  traffic = {k: {"dram_read_bytes": e.get("rd", 0), "dram_write_bytes": e.get("wr", 0),
                 "duration_ms": e.get("duration", 0)} for k, e in entries.items()}
  with open(json_path, "w") as f:
    json.dump(traffic, f, indent=1)
  print(open(table_path).read())


if __name__ == "__main__":
  main()
