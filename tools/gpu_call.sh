set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
T=r3c
for N in 4 2; do
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$N bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/${T}_bench_n$N.json 2> gpurun_out/${T}_bench_n$N.err
tail -2 gpurun_out/${T}_bench_n$N.err
python -c "
import json;j=json.loads([l for l in open('gpurun_out/${T}_bench_n$N.json') if l.startswith('{')][-1])
print($N, {k:j[k] for k in ('ms_per_step','value','gpu_launches')}, j['e2e']['ms_per_step'], j['e2e']['serial_ms_per_step'], j['dropin']['ms_per_step'], j['e2e_dropin']['ms_per_step'], j['roofline']['kernel_ms'], (j['train'] or {}).get('value'))
"
done
