# compute-sanitizer over the small GPU parity tests: memcheck (out-of-bounds / misaligned accesses, leaks of
# device errors) and racecheck (shared-memory hazards in the warp-synchronous kernels)
mkdir -p gpurun_out
T='tests/test_gpu_parity.py'
K='kat_sketches or g1_matches or tophits_without or heavy_paths or random_reads or fast_kernel or wide_location or sketch_geometries or lowest_rank or classify_on_device'
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 99 --print-limit 20 \
    python -m pytest $T -x -q -k "$K" > gpurun_out/sanitize_memcheck.log 2>&1
echo "memcheck exit $?"; grep -E "ERROR SUMMARY|passed|failed|Invalid|out of bounds" gpurun_out/sanitize_memcheck.log | tail -8
timeout 1200 compute-sanitizer --tool racecheck --error-exitcode 99 --print-limit 20 \
    python -m pytest $T -x -q -k "g1_tophits_without or fast_kernel_tophits or sketch_geometries" > gpurun_out/sanitize_racecheck.log 2>&1
echo "racecheck exit $?"; grep -E "RACECHECK SUMMARY|passed|failed|hazard" gpurun_out/sanitize_racecheck.log | tail -8
