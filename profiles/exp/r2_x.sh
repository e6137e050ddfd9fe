# round 2, call X (1 GPU): full GPU suite with the final library + shim (merged load through the C++ shim)
mkdir -p gpurun_out
( timeout 1800 python -m pytest tests -m gpu -q --durations=6 2>&1 | tail -16 ) > gpurun_out/pytest_gpu.log
tail -6 gpurun_out/pytest_gpu.log
python __graft_entry__.py smoke 2>&1 | tail -2
