set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 600 > gpurun_out/r2j_pytest.log 2>&1; echo "pytest rc=$?"; grep -n "^FAILED\|passed\|failed\|^E  " gpurun_out/r2j_pytest.log | tail -25
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2j_bench_n1.json 2> gpurun_out/r2j_bench_n1.err; echo "bench rc=$?"; cut -c1-200 gpurun_out/r2j_bench_n1.json; tail -3 gpurun_out/r2j_bench_n1.err
timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r2j_bench_ref.json 2> gpurun_out/r2j_bench_ref.err; echo "ref rc=$?"; cut -c1-300 gpurun_out/r2j_bench_ref.json; tail -3 gpurun_out/r2j_bench_ref.err
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
