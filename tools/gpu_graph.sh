#!/bin/bash
TAG=${1:-g1}; OUT=gpurun_out/$TAG; mkdir -p $OUT
export CUDA_DEVICE_MAX_CONNECTIONS=32
echo "== pytest graph"; timeout 600 python -m pytest tests/test_engine_gpu.py -q -k "graph" > $OUT/pytest_graph.log 2>&1; echo "rc=$?"; tail -25 $OUT/pytest_graph.log
echo "== bench graph"; timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > $OUT/bench.json 2> $OUT/bench.err; echo "rc=$?"; tail -5 $OUT/bench.err; cat $OUT/bench.json | cut -c1-2500
echo "== bench small graph"; timeout 600 python bench.py --workload small --steps 20 --warmup 3 --no-cpu-baseline > $OUT/bench_small.json 2> $OUT/bench_small.err; echo "rc=$?"; tail -3 $OUT/bench_small.err; cat $OUT/bench_small.json | cut -c1-400
