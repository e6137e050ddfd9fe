# round 2, call Q (1 GPU): both bench arms after the AVX-512 packer (e2e), pack test on the box's CPU
mkdir -p gpurun_out
lscpu | grep -E "Model name|^CPU\(s\)|Thread|Core|Socket" > gpurun_out/cpu_q.txt; grep -o "avx512bw\|avx512f \|avx2" /proc/cpuinfo | sort | uniq -c >> gpurun_out/cpu_q.txt
cat gpurun_out/cpu_q.txt
python -m pytest tests/test_pack.py -q 2>&1 | tail -2
( timeout 600 python bench.py --impl reference 2>&1 | tail -1 ) > gpurun_out/bench_ref.log
( timeout 900 python bench.py 2>gpurun_out/bench_full.err | tail -1 ) > gpurun_out/bench_full.log
python - <<PY
import json
r=json.load(open("gpurun_out/bench_ref.log")); d=json.load(open("gpurun_out/bench_full.log"))
print("ref", round(r["value"]/1e6,3), "M on", r["cpu_baseline"]["cores"], "cores; value", round(d["value"]/1e6,1), "e2e", round(d["e2e"]["value"]/1e6,1), d["e2e"]["host_ms_per_thread_per_step"], "prefilled", round(d["e2e"]["prefilled"]["value"]/1e6,1), "ratio e2e", round(d["e2e"]["value"]/r["value"],1), "parity", d["parity"]["mismatches"], d["parity"]["reads"])
PY
tail -3 gpurun_out/bench_full.err
