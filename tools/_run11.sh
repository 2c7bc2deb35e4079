set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 600 > gpurun_out/r2k_pytest.log 2>&1; echo "pytest rc=$?"; grep -n "^FAILED\|passed\|failed\|^E  " gpurun_out/r2k_pytest.log | tail -25 | cut -c1-300
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2k_bench_n1.json 2> gpurun_out/r2k_bench_n1.err; echo "bench rc=$?"; cut -c1-200 gpurun_out/r2k_bench_n1.json; tail -3 gpurun_out/r2k_bench_n1.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r2k_launches_localworld_p8.csv python tools/local_profile.py 8 > gpurun_out/r2k_lp.log 2>&1; python tools/launch_summary.py gpurun_out/r2k_launches_localworld_p8.csv 1 > gpurun_out/r2k_launches_localworld_p8.summary.txt; head -32 gpurun_out/r2k_launches_localworld_p8.summary.txt
