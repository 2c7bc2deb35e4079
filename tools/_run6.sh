set -x
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r2f_launches_rmat1m_n1.csv python bench.py --steps 2 --warmup 3 --no-graph --no-cpu-baseline --no-e2e --ncu-region > gpurun_out/r2f_l1.log 2>&1; echo "rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r2f_launches_products_n1.csv python bench.py --workload products-shaped --steps 2 --warmup 3 --no-graph --no-cpu-baseline --no-e2e --ncu-region > gpurun_out/r2f_l2.log 2>&1; echo "rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r2f_launches_reddit_pp_n1.csv python bench.py --workload reddit-shaped --use-pp --steps 2 --warmup 3 --no-graph --no-cpu-baseline --no-e2e --ncu-region > gpurun_out/r2f_l3.log 2>&1; echo "rc=$?"
for w in products-shaped reddit-shaped; do timeout 600 python bench.py --workload $w --steps 10 --warmup 3 --no-cpu-baseline $( [ $w = reddit-shaped ] && echo --use-pp ) > gpurun_out/r2f_bench_${w}_n1.json 2> gpurun_out/r2f_bench_${w}_n1.err; echo "rc=$?"; cut -c1-160 gpurun_out/r2f_bench_${w}_n1.json; done
for f in rmat1m products reddit_pp; do python tools/launch_summary.py gpurun_out/r2f_launches_${f}_n1.csv 2 > gpurun_out/r2f_launches_${f}_n1.summary.txt; head -24 gpurun_out/r2f_launches_${f}_n1.summary.txt; done
du -sh gpurun_out
