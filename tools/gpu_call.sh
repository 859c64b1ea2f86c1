set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python tools/tune_bwd.py --variants=0,1,2,3,4,5,6,7,8,9,10,11,12 --reps 5 > gpurun_out/r2e_tune_8192.log 2>&1
timeout 300 python tools/tune_bwd.py --variants=0,4,6,9 --reps 7 --rows 1024 > gpurun_out/r2e_tune_1024.log 2>&1
timeout 300 python tools/tune_bwd.py --variants=0,4 --reps 5 --batch 4096 --zdim 512 > gpurun_out/r2e_tune_d512.log 2>&1
timeout 300 python tools/tune_bwd.py --variants=0 --reps 5 --batch 4096 --zdim 64 > gpurun_out/r2e_tune_d64.log 2>&1
timeout 300 python tools/tune_bwd.py --variants=0 --reps 5 --batch 4000 --zdim 20 > gpurun_out/r2e_tune_d20.log 2>&1
cat gpurun_out/r2e_tune_*.log
timeout 2400 python -m pytest tests -m gpu -q --timeout=1500 -k "cfg4 or peer or non_finite or seeded or golden" > gpurun_out/r2e_pytest.log 2>&1
tail -5 gpurun_out/r2e_pytest.log
