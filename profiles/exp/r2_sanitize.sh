# round 2: compute-sanitizer over the small GPU parity tests AND the new paths (feature shards as threads, merged
# tables, CTA-wide distinct tier, single-hit filter): memcheck, then racecheck on the warp-synchronous kernels
mkdir -p gpurun_out
K='kat_sketches or g1_matches or tophits_without or heavy_paths or random_reads or fast_kernel or wide_location or sketch_geometries or lowest_rank or classify_on_device or host_packed'
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 99 --print-limit 20 \
    python -m pytest tests/test_gpu_parity.py tests/test_gpu_shard.py -x -q -k "$K or feature_shards or single_shard or all_parts_merged or routing" > gpurun_out/sanitize_memcheck_r2.log 2>&1
echo "memcheck exit $?"; grep -E "ERROR SUMMARY|passed|failed|Invalid|out of bounds" gpurun_out/sanitize_memcheck_r2.log | tail -8
timeout 1200 compute-sanitizer --tool racecheck --error-exitcode 99 --print-limit 20 \
    python -m pytest tests/test_gpu_parity.py tests/test_gpu_shard.py -x -q -k "g1_tophits_without or fast_kernel_tophits or all_parts_merged or single_shard" > gpurun_out/sanitize_racecheck_r2.log 2>&1
echo "racecheck exit $?"; grep -E "RACECHECK SUMMARY|passed|failed|hazard" gpurun_out/sanitize_racecheck_r2.log | tail -8
