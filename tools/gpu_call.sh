set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python tools/tune_bwd.py --variants=-1,40,43,44,45,41,46,42 --reps 5 > gpurun_out/r2b_tune_8192.log 2>&1
timeout 300 python tools/tune_bwd.py --variants=-1,40,43,44 --reps 7 --rows 1024 > gpurun_out/r2b_tune_1024.log 2>&1
timeout 300 python tools/tune_bwd.py --variants=-1,40,43 --reps 5 --batch 4096 --zdim 512 > gpurun_out/r2b_tune_d512.log 2>&1
timeout 300 python tools/tune_bwd.py --variants=-1,40,43 --reps 5 --batch 4096 --zdim 64 > gpurun_out/r2b_tune_d64.log 2>&1
timeout 300 python tools/tune_bwd.py --variants=-1,40,43 --reps 5 --batch 4000 --zdim 20 > gpurun_out/r2b_tune_d20.log 2>&1
timeout 600 python -m pytest tests/test_reference_dropin_gpu.py -m gpu -q > gpurun_out/r2b_pytest.log 2>&1
tail -3 gpurun_out/r2b_pytest.log
cat gpurun_out/r2b_tune_*.log
