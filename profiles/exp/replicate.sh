# N independent replicas of the single-part database (the reference's -replicate mode): bench at N GPUs
mkdir -p gpurun_out
N=${1:-2}
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29536 bench.py --gpus $N --replicate --steps 5 --warmup 3 > gpurun_out/bench_replicas_n$N.log 2>&1
tail -1 gpurun_out/bench_replicas_n$N.log | cut -c1-1800
