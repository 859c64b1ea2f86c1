set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
T=r2x
timeout 900 python -m pytest tests -q -m gpu -x > gpurun_out/${T}_pytest.log 2>&1; tail -3 gpurun_out/${T}_pytest.log
timeout 300 python bench.py --steps 3 --warmup 3 --no-graph --no-cpu-baseline --no-train > gpurun_out/${T}_bench_nograph.json 2> gpurun_out/${T}_bench_nograph.err && \
timeout 900 ncu --set full --clock-control none --import-source on -k "regex:tc_fwd_kernel|tc_bwd_ds_kernel" --launch-skip 8 --launch-count 2 -f -o gpurun_out/${T}_full python bench.py --steps 3 --warmup 3 --no-graph --no-cpu-baseline --no-train > gpurun_out/${T}_ncu.log 2>&1; tail -2 gpurun_out/${T}_ncu.log | cut -c1-200
