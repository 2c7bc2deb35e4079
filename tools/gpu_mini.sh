#!/bin/bash
TAG=${1:-mini}; OUT=gpurun_out/$TAG; mkdir -p $OUT
export CUDA_DEVICE_MAX_CONNECTIONS=32
timeout 200 python -m pytest tests -m gpu -q -x 2>&1 | tail -4
timeout 200 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-e2e > $OUT/bench.json 2> $OUT/bench.err; tail -2 $OUT/bench.err
python - <<PY
import json
d=json.load(open("$OUT/bench.json"))
print("e/s", round(d["value"],1), "ms/step", round(d["ms_per_step"],2), "eager", round(d["eager_ms_per_step"],2), "agg avg ms", round(d["roofline"]["avg_launch_ms"],3), "frac", round(d["roofline"]["frac"],4), "graph", d["cuda_graph"])
PY
