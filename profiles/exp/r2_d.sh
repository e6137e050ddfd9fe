# round 2, call D (2 GPUs): NCCL feature-sharded parity test, bench at N=2 (feature- and target-sharded)
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/gpus_d.txt
( timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -x -q 2>&1 | tail -30 ) > gpurun_out/pytest_multi_d.log
tail -8 gpurun_out/pytest_multi_d.log
( timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 2>gpurun_out/bench_n2_d.err | tail -1 ) > gpurun_out/bench_n2_d.log
cut -c1-3500 gpurun_out/bench_n2_d.log; tail -15 gpurun_out/bench_n2_d.err
