"""Summarise `ncu --page source --csv` output: stall-reason totals and the hottest SASS lines.
usage: ncu -i rep.ncu-rep --page source --csv --kernel-name regex:NAME --launch-count 1 | python tools/ncu_stalls.py"""
import csv
import sys

rows = list(csv.reader(sys.stdin))
start = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr, data = rows[start], [r for r in rows[start + 1:] if len(r) >= len(rows[start]) and r[0].startswith("0x")]
idx = {h: i for i, h in enumerate(hdr)}
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
tot = dict.fromkeys(stalls, 0)
total, items = 0, []
for r in data:
  n = int(r[idx["# Samples"]])
  total += n
  per = {s: int(r[idx[s]]) for s in stalls}
  for s, v in per.items():
    tot[s] += v
  items.append((n, r[idx["Source"]].strip(), {s: v for s, v in per.items() if v}))
print("total samples", total, " instructions", len(data))
for s, v in sorted(tot.items(), key=lambda kv: -kv[1])[:8]:
  print(f"  {s:26s} {v:8d} {100 * v / max(total, 1):5.1f}%")
top = int(sys.argv[1]) if len(sys.argv) > 1 else 25
for n, src, st in sorted(items, key=lambda x: -x[0])[:top]:
  print(f"{n:7d} {100 * n / max(total, 1):5.1f}%  {src[:72]:72s} "
        f"{dict(sorted(st.items(), key=lambda kv: -kv[1])[:3])}")
