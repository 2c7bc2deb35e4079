set -x
mkdir -p gpurun_out
nvidia-smi -L | head -2
timeout 1500 python -m pytest tests -m gpu -q -x --timeout 600 > gpurun_out/r2a_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/r2a_pytest.log
tail -15 gpurun_out/r2a_pytest.log
timeout 200 tools/bin/gather_micro > gpurun_out/r2a_gather_micro.jsonl 2>&1; echo "micro rc=$?"; cat gpurun_out/r2a_gather_micro.jsonl
timeout 300 python tools/agg_bench.py rmat-1m 1 bf16 256 > gpurun_out/r2a_agg_bench_p1.jsonl 2>&1; cat gpurun_out/r2a_agg_bench_p1.jsonl
timeout 300 python tools/agg_bench.py rmat-1m 8 bf16 256 > gpurun_out/r2a_agg_bench_p8.jsonl 2>&1; cat gpurun_out/r2a_agg_bench_p8.jsonl
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2a_bench_n1.json 2> gpurun_out/r2a_bench_n1.err; echo "bench rc=$?"; tail -c 3000 gpurun_out/r2a_bench_n1.json; tail -5 gpurun_out/r2a_bench_n1.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:agg -o gpurun_out/r2a_agg_prof python tools/agg_bench.py rmat-1m 1 bf16 256 --once > gpurun_out/r2a_ncu.log 2>&1; echo "ncu rc=$?"; tail -3 gpurun_out/r2a_ncu.log
ls -la gpurun_out | tail -12
