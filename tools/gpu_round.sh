#!/bin/bash
# One gpurun call: smoke, GPU parity tests, bench, ncu launch list + full capture of the aggregate kernel.
# usage: gpurun --timeout 1500 -- 'bash tools/gpu_round.sh [tag]'
TAG=${1:-r1}
OUT=gpurun_out/$TAG
mkdir -p $OUT
export CUDA_DEVICE_MAX_CONNECTIONS=32
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/gpu.txt 2>&1
echo "== smoke"; timeout 300 python __graft_entry__.py smoke > $OUT/smoke.log 2>&1; echo "smoke rc=$?"; tail -3 $OUT/smoke.log
echo "== pytest gpu"; timeout 900 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -15 $OUT/pytest_gpu.log
echo "== reference arm"; timeout 300 python bench.py --impl reference --steps 3 --warmup 2 > $OUT/bench_reference.json 2> $OUT/bench_reference.err; echo "rc=$?"; cut -c1-600 $OUT/bench_reference.json
echo "== bench"; timeout 600 python bench.py --steps 20 --warmup 3 > $OUT/bench.json 2> $OUT/bench.err; echo "rc=$?"; tail -5 $OUT/bench.err; cat $OUT/bench.json
if [ "${SKIP_NCU:-0}" != "1" ]; then
echo "== ncu launches"; timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file $OUT/launches.csv python bench.py --steps 2 --warmup 2 --no-cpu-baseline --no-e2e --ncu-region > $OUT/ncu_launches.log 2>&1; echo "rc=$?"
echo "== ncu full agg"; timeout 600 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:"agg_kernel|linear_tcgen05" -c 8 -o $OUT/prof_agg -f python bench.py --steps 1 --warmup 2 --no-cpu-baseline --no-e2e --ncu-region > $OUT/ncu_full.log 2>&1; echo "rc=$?"; tail -3 $OUT/ncu_full.log
fi
ls -la $OUT
