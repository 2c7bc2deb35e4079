set -x
mkdir -p gpurun_out
rm -f gpurun_out/*.ncu-rep
timeout 900 python -m pytest tests/test_kernels_gpu.py -m gpu -q --timeout 300 > gpurun_out/r2c_pytest_kernels.log 2>&1; echo "kernels rc=$?"; grep -n "^FAILED\|passed\|failed" gpurun_out/r2c_pytest_kernels.log | tail -25
timeout 300 python tools/agg_bench.py rmat-1m 1 bf16 256 2>/dev/null > gpurun_out/r2c_agg_bench_p1.jsonl; cut -c1-40,330-900 gpurun_out/r2c_agg_bench_p1.jsonl
timeout 300 python tools/agg_bench.py rmat-1m 8 bf16 256 2>/dev/null > gpurun_out/r2c_agg_bench_p8.jsonl; cut -c1-40,330-900 gpurun_out/r2c_agg_bench_p8.jsonl
timeout 300 python tools/agg_bench.py reddit-shaped 8 fp32 256 2>/dev/null > gpurun_out/r2c_agg_bench_reddit_p8.jsonl; cut -c1-40,330-900 gpurun_out/r2c_agg_bench_reddit_p8.jsonl
timeout 300 python tools/agg_bench.py reddit-shaped 1 fp32 602 2>/dev/null > gpurun_out/r2c_agg_bench_reddit_p1_602.jsonl; cut -c1-40,330-900 gpurun_out/r2c_agg_bench_reddit_p1_602.jsonl
timeout 1800 python -m pytest tests -m gpu -q --timeout 600 --deselect tests/test_kernels_gpu.py > gpurun_out/r2c_pytest_rest.log 2>&1; echo "rest rc=$?"; grep -n "^FAILED\|passed\|failed\|^E  .*Error" gpurun_out/r2c_pytest_rest.log | tail -40
timeout 600 ncu --set full --clock-control none -k regex:agg -c 6 -o gpurun_out/r2c_agg_prof python tools/agg_bench.py rmat-1m 1 bf16 256 --once --variants 2h4,3h4 > gpurun_out/r2c_ncu.log 2>&1; echo "ncu rc=$?"; tail -3 gpurun_out/r2c_ncu.log
du -sh gpurun_out
