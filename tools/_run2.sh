set -x
mkdir -p gpurun_out
rm -f gpurun_out/*.ncu-rep
timeout 900 python -m pytest tests/test_kernels_gpu.py -m gpu -q --timeout 300 > gpurun_out/r2b_pytest_kernels.log 2>&1; echo "kernels rc=$?"; tail -25 gpurun_out/r2b_pytest_kernels.log
timeout 300 python tools/agg_bench.py rmat-1m 1 bf16 256 > gpurun_out/r2b_agg_bench_p1.jsonl 2>&1; cat gpurun_out/r2b_agg_bench_p1.jsonl
timeout 300 python tools/agg_bench.py rmat-1m 8 bf16 256 > gpurun_out/r2b_agg_bench_p8.jsonl 2>&1; cat gpurun_out/r2b_agg_bench_p8.jsonl
timeout 300 python tools/agg_bench.py reddit-shaped 8 fp32 256 > gpurun_out/r2b_agg_bench_reddit_p8.jsonl 2>&1; cat gpurun_out/r2b_agg_bench_reddit_p8.jsonl
timeout 1500 python -m pytest tests -m gpu -q --timeout 600 --maxfail=8 --deselect tests/test_kernels_gpu.py > gpurun_out/r2b_pytest_rest.log 2>&1; echo "rest rc=$?"; tail -60 gpurun_out/r2b_pytest_rest.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2b_bench_n1.json 2> gpurun_out/r2b_bench_n1.err; echo "bench rc=$?"; tail -c 2500 gpurun_out/r2b_bench_n1.json; tail -5 gpurun_out/r2b_bench_n1.err
timeout 600 ncu --set full --clock-control none -k regex:agg -c 4 -o gpurun_out/r2b_agg_prof python tools/agg_bench.py rmat-1m 1 bf16 256 --once --variants 1,2h4 > gpurun_out/r2b_ncu.log 2>&1; echo "ncu rc=$?"; tail -3 gpurun_out/r2b_ncu.log
ls -la gpurun_out | tail -12; du -sh gpurun_out
