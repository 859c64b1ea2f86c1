set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
T=r2z
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/${T}_bench_n1.json 2> gpurun_out/${T}_bench_n1.err; cat gpurun_out/${T}_bench_n1.json | cut -c1-200; tail -3 gpurun_out/${T}_bench_n1.err
timeout 300 python bench.py --steps 3 --warmup 3 --no-graph --no-cpu-baseline --no-train > gpurun_out/${T}_bench_nograph.json 2> gpurun_out/${T}_bench_nograph.err && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/${T}_launches.csv python bench.py --steps 3 --warmup 3 --no-graph --no-cpu-baseline --no-train > gpurun_out/${T}_ncu.log 2>&1; tail -1 gpurun_out/${T}_ncu.log | cut -c1-200
