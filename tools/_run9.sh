set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 600 -x > gpurun_out/r2i_pytest.log 2>&1; echo "pytest rc=$?"; grep -n "^FAILED\|passed\|failed\|^E  " gpurun_out/r2i_pytest.log | tail -15
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2i_bench_n1.json 2> gpurun_out/r2i_bench_n1.err; echo "bench rc=$?"; cut -c1-200 gpurun_out/r2i_bench_n1.json; tail -3 gpurun_out/r2i_bench_n1.err
PG_FUSED_DROPOUT=0 timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-e2e > gpurun_out/r2i_bench_n1_unfused.json 2>/dev/null; cut -c1-200 gpurun_out/r2i_bench_n1_unfused.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r2i_launches_rmat1m_n1.csv python bench.py --steps 2 --warmup 3 --no-graph --no-cpu-baseline --no-e2e --ncu-region > gpurun_out/r2i_l1.log 2>&1; python tools/launch_summary.py gpurun_out/r2i_launches_rmat1m_n1.csv 2 > gpurun_out/r2i_launches_rmat1m_n1.summary.txt; head -16 gpurun_out/r2i_launches_rmat1m_n1.summary.txt
