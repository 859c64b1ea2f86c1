set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/check_sharded.py > gpurun_out/r2k_check_sharded.log 2>&1
echo "rc=$?"; grep -c " OK" gpurun_out/r2k_check_sharded.log; grep -c MISMATCH gpurun_out/r2k_check_sharded.log; tail -3 gpurun_out/r2k_check_sharded.log | cut -c1-300
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 5 --no-train > gpurun_out/r2k_bench_n2.json 2> gpurun_out/r2k_bench_n2.err
echo "rc=$?"; tail -3 gpurun_out/r2k_bench_n2.err
python -c "
import json;j=json.loads([l for l in open('gpurun_out/r2k_bench_n2.json') if l.startswith('{')][-1])
print({k:j[k] for k in ('ms_per_step','value','gpu_launches')}, j['e2e']['ms_per_step'], j['dropin']['ms_per_step'], j['e2e_dropin']['ms_per_step'], j['roofline']['kernel_ms'], j['config']['exchange'])
"
TCELBO_PEER_SYNC=host timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 20 --warmup 5 --no-train > gpurun_out/r2k_bench_n2_hostsync.json 2> gpurun_out/r2k_bench_n2_hostsync.err
python -c "
import json;j=json.loads([l for l in open('gpurun_out/r2k_bench_n2_hostsync.json') if l.startswith('{')][-1])
print('hostsync', {k:j[k] for k in ('ms_per_step','value','gpu_launches')}, j['e2e']['ms_per_step'], j['dropin']['ms_per_step'])
"
