mkdir -p gpurun_out
( timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 ) > gpurun_out/pytest_gpu.log
tail -8 gpurun_out/pytest_gpu.log
timeout 900 python profiles/exp/file_e2e.py 10000000 > gpurun_out/file_e2e.log 2>&1
tail -3 gpurun_out/file_e2e.log | cut -c1-2500
