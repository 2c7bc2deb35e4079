set -x
mkdir -p gpurun_out
timeout 300 python tools/tf32_trunc_probe.py 2>&1 | tail -4
timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_engine_gpu.py tests/test_baseline_configs_gpu.py -m gpu -q --timeout 600 -x > gpurun_out/r2g_pytest.log 2>&1; echo "pytest rc=$?"; grep -n "^FAILED\|passed\|failed\|^E  .*Error" gpurun_out/r2g_pytest.log | tail -12
timeout 600 python bench.py --workload products-shaped --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r2g_bench_products_n1.json 2> gpurun_out/r2g_bench_products_n1.err; echo "rc=$?"; cut -c1-170 gpurun_out/r2g_bench_products_n1.json; tail -3 gpurun_out/r2g_bench_products_n1.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29543 bench.py --gpus 2 --steps 5 --warmup 3 --workload papers100m-shaped --scale-down 16 --no-parity --no-e2e > gpurun_out/r2g_bench_papers_div16_n2.json 2> gpurun_out/r2g_bench_papers_div16_n2.err; echo "papers rc=$?"; tail -c 1800 gpurun_out/r2g_bench_papers_div16_n2.json; grep -v "^\*\|OMP" gpurun_out/r2g_bench_papers_div16_n2.err | tail -8
