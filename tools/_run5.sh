set -x
mkdir -p gpurun_out
nvidia-smi -L
timeout 900 python -m pytest tests/test_dist_gpu.py -m gpu -q --timeout 800 > gpurun_out/r2e_pytest_dist.log 2>&1; echo "dist rc=$?"; tail -15 gpurun_out/r2e_pytest_dist.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2e_bench_n2.json 2> gpurun_out/r2e_bench_n2.err; echo "bench2 rc=$?"; tail -c 2500 gpurun_out/r2e_bench_n2.json; tail -8 gpurun_out/r2e_bench_n2.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus 2 --steps 10 --warmup 3 --workload products-shaped --no-parity > gpurun_out/r2e_bench_products_n2.json 2> gpurun_out/r2e_bench_products_n2.err; echo "products rc=$?"; tail -c 1500 gpurun_out/r2e_bench_products_n2.json; tail -8 gpurun_out/r2e_bench_products_n2.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29543 bench.py --gpus 2 --steps 5 --warmup 3 --workload papers100m-shaped --scale-down 16 --no-parity --no-e2e > gpurun_out/r2e_bench_papers_div16_n2.json 2> gpurun_out/r2e_bench_papers_div16_n2.err; echo "papers rc=$?"; tail -c 1500 gpurun_out/r2e_bench_papers_div16_n2.json; tail -8 gpurun_out/r2e_bench_papers_div16_n2.err
du -sh gpurun_out
