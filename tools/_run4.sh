set -x
mkdir -p gpurun_out
rm -f gpurun_out/*.ncu-rep
timeout 900 python -m pytest tests/test_kernels_gpu.py -m gpu -q --timeout 300 > gpurun_out/r2d_pytest_kernels.log 2>&1; echo "kernels rc=$?"; grep -n "^FAILED\|passed\|failed" gpurun_out/r2d_pytest_kernels.log | tail -25
timeout 300 python tools/agg_bench.py rmat-1m 1 bf16 256 2>/dev/null > gpurun_out/r2d_agg_bench_p1.jsonl; cut -c1-40,330-900 gpurun_out/r2d_agg_bench_p1.jsonl
timeout 300 python tools/agg_bench.py rmat-1m 8 bf16 256 2>/dev/null > gpurun_out/r2d_agg_bench_p8.jsonl; cut -c1-40,330-900 gpurun_out/r2d_agg_bench_p8.jsonl
timeout 300 python tools/agg_bench.py rmat-1m 1 bf16 64 2>/dev/null > gpurun_out/r2d_agg_bench_p1_d64.jsonl; cut -c1-40,330-900 gpurun_out/r2d_agg_bench_p1_d64.jsonl
timeout 300 python tools/agg_bench.py reddit-shaped 8 fp32 256 2>/dev/null > gpurun_out/r2d_agg_bench_reddit_p8.jsonl; cut -c1-40,330-900 gpurun_out/r2d_agg_bench_reddit_p8.jsonl
timeout 1800 python -m pytest tests -m gpu -q --timeout 600 --deselect tests/test_kernels_gpu.py > gpurun_out/r2d_pytest_rest.log 2>&1; echo "rest rc=$?"; grep -n "^FAILED\|passed\|failed\|^E  .*Error" gpurun_out/r2d_pytest_rest.log | tail -40
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2d_bench_n1.json 2> gpurun_out/r2d_bench_n1.err; echo "bench rc=$?"; tail -c 1800 gpurun_out/r2d_bench_n1.json; tail -5 gpurun_out/r2d_bench_n1.err
PG_NARROW=0 timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-e2e > gpurun_out/r2d_bench_n1_nonarrow.json 2>/dev/null; cut -c1-200 gpurun_out/r2d_bench_n1_nonarrow.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2d_launches.csv python bench.py --steps 2 --warmup 3 --no-graph --no-cpu-baseline --no-e2e > gpurun_out/r2d_launches.log 2>&1; echo "launches rc=$?"
du -sh gpurun_out
