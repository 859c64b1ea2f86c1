set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
T=r2u
timeout 900 python -m pytest tests -q -m gpu -x > gpurun_out/${T}_pytest.log 2>&1; tail -5 gpurun_out/${T}_pytest.log
timeout 300 python tools/tune_bwd.py --variants=0,1,2,3,4,5,6 --reps 7 > gpurun_out/${T}_tune.log 2>&1; cat gpurun_out/${T}_tune.log
timeout 300 python tools/tune_bwd.py --variants=0,1,2,5,6 --reps 7 --rows 1024 > gpurun_out/${T}_tune_r1024.log 2>&1; cat gpurun_out/${T}_tune_r1024.log
