"""Collect measured bench lines (profiles/*bench*.json) into profiles/r2_secondary.json: the compact list `bench.py`
attaches to its default line as `"secondary"` (north-star configs at the GPU counts they were measured at).

    python tools/make_secondary.py profiles/r2h_bench_*.json
"""
import json
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
out = []
for f in sorted(sys.argv[1:]):
    txt = Path(f).read_text().strip().splitlines()
    if not txt:
        continue
    try:
        d = json.loads(txt[-1])
    except json.JSONDecodeError:
        continue
    if "value" not in d:
        continue
    r = d.get("roofline") or {}
    out.append({
        "workload": d["config"]["workload"].split(":")[0], "n_gpus": d["n_gpus"], "dtype": d["dtype"],
        "partition_method": d["config"].get("partition_method"), "use_pp": d["config"].get("use_pp"),
        "scale_down": d["config"].get("scale_down", 1),
        "n_nodes": d["config"].get("n_nodes"), "n_edges": d["config"].get("n_edges"),
        "epochs_per_s": round(d["value"], 3), "ms_per_step": round(d["ms_per_step"], 3),
        "exposed_comm_s_per_epoch": d.get("exposed_comm_s_per_epoch"), "exposed_comm_frac": d.get("exposed_comm_frac"),
        "aggregate_roofline_frac": r.get("frac"), "aggregate_avg_launch_ms": r.get("avg_launch_ms"),
        "e2e_epochs_per_s": (d.get("e2e") or {}).get("value"),
        "parity_ok": (d.get("parity") or {}).get("ok"), "source": str(Path(f).relative_to(ROOT)) if Path(f).is_absolute() else f,
    })
(ROOT / "profiles" / "r2_secondary.json").write_text(json.dumps(out, indent=1) + "\n")
print(f"{len(out)} lines -> profiles/r2_secondary.json")
