#!/bin/bash
# quick A/B: tests + bench variants
TAG=${1:-q}
OUT=gpurun_out/$TAG
mkdir -p $OUT
export CUDA_DEVICE_MAX_CONNECTIONS=32
echo "== pytest gpu"; timeout 900 python -m pytest tests -m gpu -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -12 $OUT/pytest_gpu.log
for U in 8; do
echo "== bench U=$U"; PG_AGG_UNROLL=$U timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > $OUT/bench_u$U.json 2> $OUT/bench_u$U.err; tail -2 $OUT/bench_u$U.err; python - <<PY
import json
d=json.load(open("$OUT/bench_u$U.json"))
print("U=$U", "ms/step", round(d["ms_per_step"],2), "agg avg ms", round(d["roofline"]["avg_launch_ms"],3), "frac", round(d["roofline"]["frac"],4), "share", round(d["roofline"]["share_of_step"],3))
PY
done
for SL in 512 1024 2048; do
echo "== bench seg_len=$SL"; PG_SEG_LEN=$SL timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > $OUT/bench_sl$SL.json 2> $OUT/bench_sl$SL.err; tail -2 $OUT/bench_sl$SL.err; python - <<PY
import json
d=json.load(open("$OUT/bench_sl$SL.json"))
print("seg_len=$SL", "ms/step", round(d["ms_per_step"],2), "agg avg ms", round(d["roofline"]["avg_launch_ms"],3))
PY
done
echo "== ncu launches"; timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file $OUT/launches.csv python bench.py --steps 2 --warmup 2 --no-cpu-baseline --no-e2e --ncu-region > $OUT/ncu_launches.log 2>&1; echo "rc=$?"
python tools/ncu_summary.py $OUT/launches.csv 14
