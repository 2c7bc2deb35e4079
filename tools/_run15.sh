set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_dist_gpu.py -m gpu -q --timeout 800 > gpurun_out/r2o_pytest_dist.log 2>&1; echo "dist rc=$?"; tail -3 gpurun_out/r2o_pytest_dist.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29571 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2o_bench_n2.json 2> gpurun_out/r2o_bench_n2.err; echo "bench2 rc=$?"; grep -v "^\*\|OMP" gpurun_out/r2o_bench_n2.err | tail -3
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2o_bench_n2.json').read().strip().splitlines()[-1])
print('value',d['value'],'e2e',d['e2e']['value'],'parity',d['parity']['ok'],d['parity']['worst_rel'],'exposed',d['exposed_comm_s_per_epoch'])
PY
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29572 bench.py --gpus 2 --impl reference --steps 2 --warmup 1 2>/dev/null | cut -c1-200
