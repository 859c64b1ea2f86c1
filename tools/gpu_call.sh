set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q --timeout=1500 -x > gpurun_out/r2f_pytest.log 2>&1
tail -15 gpurun_out/r2f_pytest.log
timeout 300 python tools/small_batch_bench.py --batch 64 > gpurun_out/r2f_small64.json 2> gpurun_out/r2f_small.err; cat gpurun_out/r2f_small64.json; tail -3 gpurun_out/r2f_small.err
timeout 300 python tools/small_batch_bench.py --batch 3 > gpurun_out/r2f_small3.json 2>> gpurun_out/r2f_small.err; cat gpurun_out/r2f_small3.json
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r2f_bench.json 2> gpurun_out/r2f_bench.err
tail -3 gpurun_out/r2f_bench.err
python -c "
import json;j=json.load(open('gpurun_out/r2f_bench.json'))
print({k:j[k] for k in ('ms_per_step','value','gpu_launches')}, j['e2e'], j['dropin'], j['e2e_dropin'], j['roofline']['kernel_ms'], j['roofline']['step_frac_of_sfu_peak'], j['cpu_baseline'])
"
