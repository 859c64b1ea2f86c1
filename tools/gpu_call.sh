set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
T=r2r
timeout 900 python -m pytest tests -q -m gpu -x > gpurun_out/${T}_pytest.log 2>&1; tail -5 gpurun_out/${T}_pytest.log
timeout 300 python tools/tune_bwd.py --variants=0 --reps 7 --fwd-map 1,0 --fwd-wave 3,4,0 --fwd-seg 16,21,32 > gpurun_out/${T}_tune_128.log 2>&1; cat gpurun_out/${T}_tune_128.log
timeout 300 python tools/tune_bwd.py --variants=0 --reps 7 --rows 1024 --fwd-map 1,0 --fwd-wave 3,4 > gpurun_out/${T}_tune_128_r1024.log 2>&1; cat gpurun_out/${T}_tune_128_r1024.log
timeout 300 python tools/tune_bwd.py --variants=0 --reps 7 --batch 4096 --zdim 64 --fwd-map 1,0 > gpurun_out/${T}_tune_64.log 2>&1; cat gpurun_out/${T}_tune_64.log
timeout 300 python tools/tune_bwd.py --variants=0 --reps 7 --batch 4000 --zdim 20 --fwd-map 1,0 > gpurun_out/${T}_tune_20.log 2>&1; cat gpurun_out/${T}_tune_20.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/${T}_bench_n1.json 2> gpurun_out/${T}_bench_n1.err; cat gpurun_out/${T}_bench_n1.json; tail -3 gpurun_out/${T}_bench_n1.err
