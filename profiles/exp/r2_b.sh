# round 2, call B: packed host path: new parity tests first (fast fail), bench both arms
mkdir -p gpurun_out
nproc > gpurun_out/nproc.txt; free -g | head -2 >> gpurun_out/nproc.txt; lscpu | grep -E "Model name|^CPU\(s\)" >> gpurun_out/nproc.txt
( timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q --durations=5 2>&1 | tail -30 ) > gpurun_out/pytest_gpu_b.log
tail -5 gpurun_out/pytest_gpu_b.log
( timeout 900 python bench.py --impl reference 2>gpurun_out/bench_ref_b.err | tail -1 ) > gpurun_out/bench_ref_b.log
cut -c1-1500 gpurun_out/bench_ref_b.log; tail -5 gpurun_out/bench_ref_b.err
( timeout 900 python bench.py 2>gpurun_out/bench_b.err | tail -1 ) > gpurun_out/bench_b.log
cut -c1-4000 gpurun_out/bench_b.log; tail -5 gpurun_out/bench_b.err
