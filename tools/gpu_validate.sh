#!/usr/bin/env bash
# End-of-round validation on one B200 (run from the repo root on the GPU box, e.g. `gpurun -- 'bash tools/gpu_validate.sh r2'`):
# the -m gpu parity suite, the smoke entry point, the bench line of both arms.  Outputs land in gpurun_out/<tag>_*.
set -x
cd "${GRAFT_REPO_ROOT:-.}"
T=${1:-val}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu > gpurun_out/${T}_pytest.log 2>&1; tail -3 gpurun_out/${T}_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/${T}_smoke.log 2>&1; tail -2 gpurun_out/${T}_smoke.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/${T}_bench_n1.json 2> gpurun_out/${T}_bench_n1.err; cut -c1-300 gpurun_out/${T}_bench_n1.json; tail -2 gpurun_out/${T}_bench_n1.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${T}_bench_reference.json 2> gpurun_out/${T}_bench_reference.err; cut -c1-300 gpurun_out/${T}_bench_reference.json
