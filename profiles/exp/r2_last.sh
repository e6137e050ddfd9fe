# round 2, last call (1 GPU, ~9 GPU-minutes left): the new host packer + e2e chunk scheduling on a GPU box
mkdir -p gpurun_out
( nvidia-smi -L; grep -m1 "model name" /proc/cpuinfo; nproc ) > gpurun_out/box.txt 2>&1
( timeout 240 python bench.py 2>gpurun_out/bench_full.err | tail -1 ) > gpurun_out/bench_full.log
( timeout 200 python -m pytest tests -m gpu -x -q 2>&1 | tail -12 ) > gpurun_out/pytest_gpu.log
( timeout 150 python bench.py --impl reference 2>&1 | tail -1 ) > gpurun_out/bench_ref.log
( timeout 100 python bench.py --no-cpu-baseline --slot-reads 250000 2>/dev/null | tail -1 ) > gpurun_out/bench_slot250k.log
( timeout 100 python bench.py --no-cpu-baseline --slot-reads 62500 2>/dev/null | tail -1 ) > gpurun_out/bench_slot62k.log
python - <<PY
import json
def ld(p):
    try: return json.load(open(p))
    except Exception as ex: return None
d=ld("gpurun_out/bench_full.log"); r=ld("gpurun_out/bench_ref.log")
if d: print("value", round(d["value"]/1e6,1), "e2e", round(d["e2e"]["value"]/1e6,1), d["e2e"]["host_ms_per_thread_per_step"], d["e2e"]["host_threads"], "thr", d["e2e"].get("packer"), "same", d["e2e"]["results_equal_device_resident_path"], "prefilled", round(d["e2e"]["prefilled"]["value"]/1e6,1), "parity", d["parity"], d.get("host"))
if r: print("ref", round(r["value"]/1e6,3), r["cpu_baseline"]["cores"])
for n in ("slot250k","slot62k"):
    x=ld("gpurun_out/bench_%s.log"%n)
    if x: print(n, round(x["e2e"]["value"]/1e6,1), x["e2e"]["host_ms_per_thread_per_step"])
PY
cat gpurun_out/box.txt; tail -3 gpurun_out/pytest_gpu.log; tail -2 gpurun_out/bench_full.err
