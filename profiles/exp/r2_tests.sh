mkdir -p gpurun_out
( timeout 1500 python -m pytest tests -m gpu -q --durations=6 2>&1 | tail -40 ) > gpurun_out/pytest_gpu.log
tail -12 gpurun_out/pytest_gpu.log
