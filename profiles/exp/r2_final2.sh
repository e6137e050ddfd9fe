# round 2, final evidence run (1 GPU): default bench, GPU tests, reference arm, C3 - the library with the faster
# packer and 32 stream channels by default
mkdir -p gpurun_out
( timeout 200 python bench.py 2>gpurun_out/bench_full.err | tail -1 ) > gpurun_out/bench_full.log
( timeout 200 python -m pytest tests -m gpu -x -q --durations=5 2>&1 | tail -14 ) > gpurun_out/pytest_gpu.log
( timeout 150 python bench.py --impl reference 2>&1 | tail -1 ) > gpurun_out/bench_ref.log
( timeout 150 python bench.py --workload C3 2>gpurun_out/bench_c3.err | tail -1 ) > gpurun_out/bench_c3.log
python - <<PY
import json
def ld(p):
    try: return json.load(open(p))
    except Exception as ex: return None
d=ld("gpurun_out/bench_full.log"); r=ld("gpurun_out/bench_ref.log"); c=ld("gpurun_out/bench_c3.log")
if d: print("value", round(d["value"]/1e6,1), "e2e", round(d["e2e"]["value"]/1e6,1), d["e2e"]["ms_per_step"], d["e2e"]["host_ms_per_thread_per_step"], "same", d["e2e"]["results_equal_device_resident_path"], "prefilled", round(d["e2e"]["prefilled"]["value"]/1e6,1), "parity", d["parity"]["mismatches"], "/", d["parity"]["reads"], "cpu", round(d["cpu_baseline"]["value"]/1e6,3))
if r: print("ref", round(r["value"]/1e6,3), r["cpu_baseline"]["cores"])
if c: print("C3", round(c["value"]/1e6,2), round(c["ms_per_step"],1), "ms e2e", round(c["e2e"]["value"]/1e6,2), c["e2e"]["host_ms_per_thread_per_step"])
PY
tail -4 gpurun_out/pytest_gpu.log; tail -2 gpurun_out/bench_full.err gpurun_out/bench_c3.err
