set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
for N in 8 4 2; do
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$N bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r2p_bench_n$N.json 2> gpurun_out/r2p_bench_n$N.err
tail -2 gpurun_out/r2p_bench_n$N.err
python -c "
import json;j=json.loads([l for l in open('gpurun_out/r2p_bench_n$N.json') if l.startswith('{')][-1])
print($N, {k:j[k] for k in ('ms_per_step','value','gpu_launches')}, j['e2e']['ms_per_step'], j['dropin']['ms_per_step'], j['e2e_dropin']['ms_per_step'], j['roofline']['kernel_ms'], (j['train'] or {}).get('value'))
"
done
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 10 --warmup 3 --batch 32768 --zdim 512 --no-train > gpurun_out/r2p_bench_cfg4_n8.json 2> gpurun_out/r2p_bench_cfg4_n8.err
python -c "
import json;j=json.loads([l for l in open('gpurun_out/r2p_bench_cfg4_n8.json') if l.startswith('{')][-1])
print('cfg4', {k:j[k] for k in ('ms_per_step','value')}, j['e2e']['ms_per_step'], j['roofline']['kernel_ms'])
"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29522 tools/train_bench.py --image 64 --zdim 128 --batch 64 --peer > gpurun_out/r2p_train_n8_64.json 2> gpurun_out/r2p_train_n8_64.err; cat gpurun_out/r2p_train_n8_64.json | cut -c1-300
