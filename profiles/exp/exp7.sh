mkdir -p gpurun_out
( timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 ) > gpurun_out/pytest_gpu.log
tail -8 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --workload C3 --steps 3 > gpurun_out/bench_c3.log 2>&1
tail -1 gpurun_out/bench_c3.log | cut -c1-3000
