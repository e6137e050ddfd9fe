# round 2, call C: feature-space sharding on one GPU (shards = threads), then the whole GPU suite
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_shard.py -m gpu -x -q --durations=5 2>&1 | tail -40 ) > gpurun_out/pytest_gpu_c.log
tail -25 gpurun_out/pytest_gpu_c.log
( timeout 1500 python -m pytest tests -m gpu -q --durations=6 2>&1 | tail -40 ) > gpurun_out/pytest_gpu_all_c.log
tail -12 gpurun_out/pytest_gpu_all_c.log
