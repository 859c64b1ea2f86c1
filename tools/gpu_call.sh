set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 300 python tools/tune_bwd.py --variants=0 --reps 2 --rows 1024 > gpurun_out/r2n_plain.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'tc_fwd_kernel|tc_bwd_ds_kernel|prep_kernel|finalize' --launch-skip 7 --launch-count 7 -o gpurun_out/r2n_rows1024 python tools/tune_bwd.py --variants=0 --reps 2 --rows 1024 > gpurun_out/r2n_ncu.log 2>&1
tail -3 gpurun_out/r2n_ncu.log
