#!/usr/bin/env bash
# ncu evidence of the bench step on one B200 (ONE ncu pass per call, and only after the same command exited 0 without ncu):
#   bash tools/gpu_profile.sh launches <tag>   -> gpurun_out/<tag>_launches.csv      (copy to profiles/r2_launches.csv)
#   bash tools/gpu_profile.sh full <tag>       -> gpurun_out/<tag>_full.ncu-rep      (export here: ncu -i ... --page raw --csv > profiles/r2_ncu_full_raw.csv)
# then `python tools/make_profile_md.py > profiles/r2_summary.md` and `python tools/sass_excerpt.py > profiles/r2_sass_hotloops.txt`.
set -x
cd "${GRAFT_REPO_ROOT:-.}"
MODE=${1:-launches}; T=${2:-prof}
mkdir -p gpurun_out
CMD="python bench.py --steps 3 --warmup 3 --no-graph --no-cpu-baseline --no-train"
timeout 300 $CMD > gpurun_out/${T}_bench_nograph.json 2> gpurun_out/${T}_bench_nograph.err || exit 1
if [ "$MODE" = full ]; then
  timeout 900 ncu --set full --clock-control none --import-source on -k "regex:tc_fwd_kernel|tc_bwd_ds_kernel" --launch-skip 8 --launch-count 2 \
      -f -o gpurun_out/${T}_full $CMD > gpurun_out/${T}_ncu.log 2>&1
else
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/${T}_launches.csv $CMD > gpurun_out/${T}_ncu.log 2>&1
fi
tail -2 gpurun_out/${T}_ncu.log | cut -c1-200
