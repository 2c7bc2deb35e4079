set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 600 > gpurun_out/r2i_pytest.log 2>&1; echo "pytest rc=$?"; grep -n "^FAILED\|passed\|failed\|^E  " gpurun_out/r2i_pytest.log | tail -25
timeout 300 python tools/agg_bench.py rmat-1m 1 bf16 64 --variants 1,2n4w,2n4 2>/dev/null > gpurun_out/r2i_agg_bench_p1_d64.jsonl; cut -c1-40,330-900 gpurun_out/r2i_agg_bench_p1_d64.jsonl
timeout 300 python tools/agg_bench.py reddit-shaped 8 fp32 41 --variants 2n4w,2n4 2>/dev/null > gpurun_out/r2i_agg_bench_reddit_p8_d41.jsonl; cut -c1-40,330-900 gpurun_out/r2i_agg_bench_reddit_p8_d41.jsonl
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2i_bench_n1.json 2> gpurun_out/r2i_bench_n1.err; echo "bench rc=$?"; cut -c1-200 gpurun_out/r2i_bench_n1.json; tail -3 gpurun_out/r2i_bench_n1.err
PG_FUSED_DROPOUT=0 timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-e2e > gpurun_out/r2i_bench_n1_unfused.json 2>/dev/null; cut -c1-200 gpurun_out/r2i_bench_n1_unfused.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r2i_launches_rmat1m_n1.csv python bench.py --steps 2 --warmup 3 --no-graph --no-cpu-baseline --no-e2e --ncu-region > gpurun_out/r2i_l1.log 2>&1; python tools/launch_summary.py gpurun_out/r2i_launches_rmat1m_n1.csv 2 > gpurun_out/r2i_launches_rmat1m_n1.summary.txt; head -18 gpurun_out/r2i_launches_rmat1m_n1.summary.txt
