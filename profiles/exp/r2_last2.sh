# round 2, very last call (1 GPU, ~6 GPU-minutes left): the extended C1 test (abundance tables from the device's
# classifications) and what bounds the e2e step now that the packer is faster: channels, worker count, slots
mkdir -p gpurun_out
( timeout 120 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "c1_bundled" 2>&1 | tail -5 ) > gpurun_out/pytest_c1.log
run () {  # name, env, flags
  ( env $2 timeout 70 python bench.py --no-cpu-baseline $3 2>/dev/null | tail -1 ) > gpurun_out/e2e_$1.log
}
run base   "X=1" ""
run conn32 "CUDA_DEVICE_MAX_CONNECTIONS=32" ""
run thr14  "X=1" "--e2e-threads 14"
run thr12  "X=1" "--e2e-threads 12"
run spw4   "X=1" "--e2e-slots-per-worker 4"
run spw2   "X=1" "--e2e-slots-per-worker 2"
run c32s4  "CUDA_DEVICE_MAX_CONNECTIONS=32" "--e2e-slots-per-worker 4"
run s100k  "X=1" "--slot-reads 100000"
run c32t14 "CUDA_DEVICE_MAX_CONNECTIONS=32" "--e2e-threads 14 --e2e-slots-per-worker 4"
python - <<PY
import json, glob
for p in sorted(glob.glob("gpurun_out/e2e_*.log")):
    try:
        x = json.load(open(p)); e = x["e2e"]
        print(p[15:-4], round(e["value"]/1e6,1), "M e2e", e["ms_per_step"], "ms", e["host_ms_per_thread_per_step"], e["host_threads"], "thr", e["slots_per_worker"], "spw", e["reads_per_slot"], "| prefilled", round(e["prefilled"]["value"]/1e6,1), "| value", round(x["value"]/1e6,1))
    except Exception as ex:
        print(p, "failed", ex)
PY
tail -3 gpurun_out/pytest_c1.log
