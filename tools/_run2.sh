set -x
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/check_sharded.py > gpurun_out/check_sharded.log 2>&1; echo "check rc=$?" >> gpurun_out/check_sharded.log
for ex in nccl peer; do
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 3 --exchange $ex --no-cpu-baseline > gpurun_out/bench_n2_$ex.log 2>&1; echo "rc=$?" >> gpurun_out/bench_n2_$ex.log
done
