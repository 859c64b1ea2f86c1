set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
TOOL=${1:-memcheck}
timeout 300 python tools/sanitize_case.py > gpurun_out/r2_san_plain.log 2>&1 && \
timeout 1500 compute-sanitizer --tool $TOOL --log-file gpurun_out/r2_san_$TOOL.log python tools/sanitize_case.py > gpurun_out/r2_san_${TOOL}_stdout.log 2>&1
echo "exit $?"
tail -5 gpurun_out/r2_san_$TOOL.log
tail -5 gpurun_out/r2_san_${TOOL}_stdout.log
