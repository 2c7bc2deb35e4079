#!/bin/bash
# multi-GPU round: IPC/NVLink parity against the oracle, then bench at this GPU count
N=${1:-2}; TAG=${2:-m$N}
OUT=gpurun_out/$TAG; mkdir -p $OUT
export CUDA_DEVICE_MAX_CONNECTIONS=32
nvidia-smi topo -m > $OUT/topo.txt 2>&1
echo "== dist parity tiny"; timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tools/dist_parity.py tiny > $OUT/parity_tiny.log 2>&1; echo "rc=$?"; grep dist_parity $OUT/parity_tiny.log | tail -6; tail -5 $OUT/parity_tiny.log | grep -v dist_parity
if [ "${QUICK:-0}" != "1" ]; then
echo "== dist parity small"; timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 tools/dist_parity.py small > $OUT/parity_small.log 2>&1; echo "rc=$?"; grep dist_parity $OUT/parity_small.log | tail -6
fi
echo "== dist parity tiny (graph)"; PG_PARITY_GRAPH=1 timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29514 tools/dist_parity.py tiny > $OUT/parity_tiny_graph.log 2>&1; echo "rc=$?"; grep dist_parity $OUT/parity_tiny_graph.log | tail -6; grep -i "error" $OUT/parity_tiny_graph.log | head -5
echo "== bench x$N"; timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --steps 20 --warmup 3 > $OUT/bench.json 2> $OUT/bench.err; echo "rc=$?"; tail -3 $OUT/bench.err; tail -1 $OUT/bench.json; echo "== bench x$N (no graph)"; timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29515 bench.py --gpus $N --steps 20 --warmup 3 --no-graph --no-e2e > $OUT/bench_nograph.json 2> $OUT/bench_nograph.err; tail -1 $OUT/bench_nograph.json | cut -c1-200
