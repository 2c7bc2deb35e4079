"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: time and count per kernel.
    python tools/launch_summary.py gpurun_out/launches.csv [n_epochs] > profiles/xyz.summary.txt"""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
n_ep = int(sys.argv[2]) if len(sys.argv) > 2 else 1
start = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
hdr = rows[start]
ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
agg = collections.defaultdict(lambda: [0, 0.0])
for r in rows[start + 1:]:
    if len(r) <= vi:
        continue
    try:
        v = float(r[vi].replace(",", ""))
    except ValueError:
        continue
    v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(r[ui], 1e-3)
    name = r[ki].split("(")[0][:90]
    agg[name][0] += 1
    agg[name][1] += v
tot = sum(v[1] for v in agg.values())
print(f"total {tot:.1f} us over {sum(v[0] for v in agg.values())} launches ({n_ep} epoch(s): {tot / n_ep:.1f} us of kernels per epoch)")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{v[1]:10.1f} us {100 * v[1] / tot:5.1f}%  n={v[0]:4d}  avg {v[1] / v[0]:8.1f} us  {k}")
