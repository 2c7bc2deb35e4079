set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 600 > gpurun_out/r2n_pytest.log 2>&1; echo "pytest rc=$?"; grep -n "^FAILED\|passed\|failed\|^E  " gpurun_out/r2n_pytest.log | tail -12 | cut -c1-300
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2n_bench_n1.json 2> gpurun_out/r2n_bench_n1.err; echo "bench rc=$?"; tail -3 gpurun_out/r2n_bench_n1.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2n_bench_n1.json').read().strip().splitlines()[-1])
print('value',d['value'],'e2e',d['e2e']['value'],'frac',d['roofline']['frac'],'launches',d['gpu_launches'],'secondary',len(d['secondary'] or []), 'cpu', d['cpu_baseline']['value'])
PY
