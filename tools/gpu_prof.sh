#!/bin/bash
TAG=${1:-prof}; OUT=gpurun_out/$TAG; mkdir -p $OUT
export CUDA_DEVICE_MAX_CONNECTIONS=32
echo "== ncu launches"; timeout 300 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file $OUT/launches.csv python bench.py --steps 2 --warmup 2 --no-cpu-baseline --no-e2e --no-graph --ncu-region > $OUT/ncu_launches.log 2>&1; echo "rc=$?"
echo "== ncu full"; timeout 300 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:"agg_kernel|linear_tcgen05" -c 4 -o $OUT/prof_agg -f python bench.py --steps 1 --warmup 2 --no-cpu-baseline --no-e2e --no-graph --ncu-region > $OUT/ncu_full.log 2>&1; echo "rc=$?"
echo "== bench"; timeout 300 python bench.py --steps 20 --warmup 3 > $OUT/bench.json 2> $OUT/bench.err; echo "rc=$?"; tail -2 $OUT/bench.err; cut -c1-400 $OUT/bench.json
