# ncu launch list of the library's kernels during bench.py (cold-cache, serialised: compare SHARES)
mkdir -p gpurun_out
K='regex:^(encode|count_windows|fill_windows|sketch|query_fast|query_warp|query_heavy|merge_candidates|count_hits|table_insert)_kernel|DeviceScan'
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k "$K" -c 400 --csv \
    --log-file gpurun_out/launches_r1.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/launches_bench.log 2>&1
tail -c 300 gpurun_out/launches_bench.log
