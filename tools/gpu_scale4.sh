#!/bin/bash
N=${1:-4}; OUT=gpurun_out/s$N; mkdir -p $OUT
export CUDA_DEVICE_MAX_CONNECTIONS=32
echo "== bench rmat-1m x$N"; timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus $N --steps 20 --warmup 3 > $OUT/bench_rmat.json 2> $OUT/bench_rmat.err; echo "rc=$?"; tail -1 $OUT/bench_rmat.json | cut -c1-300
echo "== bench reddit-shaped x$N"; timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus $N --workload reddit-shaped --steps 10 --warmup 3 --no-e2e > $OUT/bench_reddit.json 2> $OUT/bench_reddit.err; echo "rc=$?"; tail -2 $OUT/bench_reddit.err | cut -c1-300; tail -1 $OUT/bench_reddit.json | cut -c1-300
