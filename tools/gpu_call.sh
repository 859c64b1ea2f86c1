set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q --timeout=1500 > gpurun_out/r2o_pytest.log 2>&1
tail -4 gpurun_out/r2o_pytest.log
timeout 300 python tools/tune_bwd.py --variants=0 --reps 9 --rows 1024 --fwd-seg 0 > gpurun_out/r2o_tune_1024.log 2>&1; cat gpurun_out/r2o_tune_1024.log
timeout 300 python tools/tune_bwd.py --variants=0 --reps 7 --fwd-seg 0 > gpurun_out/r2o_tune_8192.log 2>&1; cat gpurun_out/r2o_tune_8192.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2o_bench_n1.json 2> gpurun_out/r2o_bench.err
tail -3 gpurun_out/r2o_bench.err
python -c "
import json;j=json.load(open('gpurun_out/r2o_bench_n1.json'))
print({k:j[k] for k in ('ms_per_step','value','gpu_launches')}, j['e2e']['ms_per_step'], j['dropin']['ms_per_step'], j['e2e_dropin']['ms_per_step'], j['roofline']['kernel_ms'], j['roofline']['step_frac_of_sfu_peak'], j['roofline']['traffic'], j['train']['value'])
"
timeout 300 python tools/small_batch_bench.py --batch 64 > gpurun_out/r2o_small64.json 2>/dev/null; cat gpurun_out/r2o_small64.json
