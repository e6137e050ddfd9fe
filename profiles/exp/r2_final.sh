# round 2, final lines from the final library (1 GPU): reference arm, default bench, C3
mkdir -p gpurun_out
( timeout 600 python bench.py --impl reference 2>&1 | tail -1 ) > gpurun_out/bench_ref.log
( timeout 900 python bench.py 2>gpurun_out/bench_full.err | tail -1 ) > gpurun_out/bench_full.log
( timeout 600 python bench.py --workload C3 2>gpurun_out/bench_c3.err | tail -1 ) > gpurun_out/bench_c3.log
python - <<PY
import json
r=json.load(open("gpurun_out/bench_ref.log")); d=json.load(open("gpurun_out/bench_full.log")); c=json.load(open("gpurun_out/bench_c3.log"))
print("ref", round(r["value"]/1e6,3), "M on", r["cpu_baseline"]["cores"], "threads; value", round(d["value"]/1e6,1), "kernel", d["roofline"]["kernel_ms_per_launch"], "frac", d["roofline"]["frac"], "e2e", round(d["e2e"]["value"]/1e6,1), d["e2e"]["host_ms_per_thread_per_step"], "prefilled", round(d["e2e"]["prefilled"]["value"]/1e6,1), "parity", d["parity"]["mismatches"], "/", d["parity"]["reads"])
print("C3", round(c["value"]/1e6,2), "M reads/s", round(c["ms_per_step"],1), "ms e2e", round(c["e2e"]["value"]/1e6,2), c["roofline"]["stage_ms_per_step"])
PY
tail -2 gpurun_out/bench_full.err gpurun_out/bench_c3.err
