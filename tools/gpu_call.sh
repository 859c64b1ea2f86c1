set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
for N in 8 4; do
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$N bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r2h_bench_n$N.json 2> gpurun_out/r2h_bench_n$N.err
tail -2 gpurun_out/r2h_bench_n$N.err
python -c "
import json;j=json.loads([l for l in open('gpurun_out/r2h_bench_n$N.json') if l.startswith('{')][-1])
print($N, {k:j[k] for k in ('ms_per_step','value','gpu_launches')}, j['e2e']['ms_per_step'], j['dropin']['ms_per_step'], j['e2e_dropin']['ms_per_step'], j['roofline']['kernel_ms'], j['config']['exchange'], j['train'])
"
done
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29520 bench.py --gpus 8 --steps 20 --warmup 5 --exchange nccl --no-train > gpurun_out/r2h_bench_n8_nccl.json 2> gpurun_out/r2h_bench_n8_nccl.err
python -c "
import json;j=json.loads([l for l in open('gpurun_out/r2h_bench_n8_nccl.json') if l.startswith('{')][-1])
print('nccl', {k:j[k] for k in ('ms_per_step','value')}, j['e2e']['ms_per_step'], j['dropin']['ms_per_step'], j['e2e_dropin']['ms_per_step'], j['config']['exchange'])
"
