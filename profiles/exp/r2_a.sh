# round 2, call A: GPU parity tests + default bench (baseline of the round on today's box)
mkdir -p gpurun_out
nproc > gpurun_out/nproc.txt; free -g | head -2 >> gpurun_out/nproc.txt
( timeout 1500 python -m pytest tests -m gpu -x -q --durations=8 2>&1 | tail -30 ) > gpurun_out/pytest_gpu_a.log
tail -5 gpurun_out/pytest_gpu_a.log
( timeout 900 python bench.py 2>gpurun_out/bench_a.err | tail -1 ) > gpurun_out/bench_a.log
cut -c1-3000 gpurun_out/bench_a.log; tail -5 gpurun_out/bench_a.err
