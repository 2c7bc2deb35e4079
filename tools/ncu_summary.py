"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: time share per kernel."""
import collections
import csv
import re
import sys


def main(path, top=30):
    lines = [l for l in open(path) if not l.startswith("==")]
    agg = collections.defaultdict(lambda: [0, 0.0])
    tot = 0.0
    for row in csv.DictReader(lines):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(row["Metric Value"].replace(",", ""))
        v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(row["Metric Unit"], 1.0)
        name = row["Kernel Name"]
        short = re.sub(r"\(.*", "", name)
        short = re.sub(r"^void ", "", short)[:110]
        agg[short][0] += 1
        agg[short][1] += v
        tot += v
    print(f"total {tot:.1f} us over {sum(c for c, _ in agg.values())} launches")
    for k, (c, t) in sorted(agg.items(), key=lambda x: -x[1][1])[:top]:
        print(f"{t:10.1f} us {100 * t / tot:5.1f}%  n={c:4d}  avg {t / c:8.1f} us  {k}")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 30)
