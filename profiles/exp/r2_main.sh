# Round-2 evidence run (one gpurun call, 1 GPU): GPU parity tests, both bench arms, C3, the ncu launch list and
# one `--set full` capture each of the two dominant kernels.  Outputs land in gpurun_out/ and are
# turned into profiles/*_r2.* by `python profiles/collect.py r2`.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
nproc > gpurun_out/nproc.txt
( timeout 1800 python -m pytest tests -m gpu -q --durations=8 2>&1 | tail -25 ) > gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
( timeout 600 python bench.py --impl reference 2>&1 | tail -1 ) > gpurun_out/bench_ref.log
( timeout 900 python bench.py 2>gpurun_out/bench_full.err | tail -1 ) > gpurun_out/bench_full.log
( timeout 600 python bench.py --workload C3 2>gpurun_out/bench_c3.err | tail -1 ) > gpurun_out/bench_c3.log
cut -c1-400 gpurun_out/bench_ref.log; cut -c1-1800 gpurun_out/bench_full.log; cut -c1-300 gpurun_out/bench_c3.log
K='regex:^(encode|count_windows|fill_windows|sketch|sketch_fast|query_fast|query_warp|query_heavy|query_cta_hash|merge_candidates|count_hits|table_insert|classify)_kernel|DeviceScan'
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k "$K" -c 400 --csv \
    --log-file gpurun_out/launches_r2.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/launches_bench.log 2>&1
tail -c 200 gpurun_out/launches_bench.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:query_fast_kernel -s 6 -c 1 -f -o gpurun_out/prof_query_r2 \
    python bench.py --reads 1000000 --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/prof_query.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:sketch_fast_kernel -s 8 -c 1 -f -o gpurun_out/prof_sketch_r2 \
    python bench.py --reads 1000000 --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/prof_sketch.log 2>&1
ls -la gpurun_out | tail -30
