set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q --timeout=1500 > gpurun_out/r2j_pytest.log 2>&1
tail -6 gpurun_out/r2j_pytest.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2j_bench_n1.json 2> gpurun_out/r2j_bench.err
tail -3 gpurun_out/r2j_bench.err
python -c "
import json;j=json.load(open('gpurun_out/r2j_bench_n1.json'))
print({k:j[k] for k in ('ms_per_step','value','gpu_launches')}, j['e2e']['ms_per_step'], j['dropin']['ms_per_step'], j['e2e_dropin']['ms_per_step'], j['roofline']['kernel_ms'], j['roofline']['step_frac_of_sfu_peak'], j['train'])
"
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r2j_bench_reference.json 2>> gpurun_out/r2j_bench.err
cut -c1-400 gpurun_out/r2j_bench_reference.json
# launch list of the eager C-ABI step (same kernels as the graph) and a full capture of the two sweeps
timeout 300 python bench.py --steps 3 --warmup 3 --no-graph --no-cpu-baseline --no-train > gpurun_out/r2j_plain.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2j_launches.csv python bench.py --steps 3 --warmup 3 --no-graph --no-cpu-baseline --no-train > gpurun_out/r2j_ncu_launch.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'tc_fwd_kernel|tc_bwd_ds_kernel' --launch-skip 8 --launch-count 2 -o gpurun_out/r2j_full python bench.py --steps 3 --warmup 3 --no-graph --no-cpu-baseline --no-train > gpurun_out/r2j_ncu_full.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/r2j_launches_b64.csv python tools/small_batch_bench.py --batch 64 --reps 5 > gpurun_out/r2j_ncu_b64.log 2>&1
ls -la gpurun_out | grep r2j
