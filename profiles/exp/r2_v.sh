# round 2, call V (1 GPU): the kernel on the 8-part and 4-part merged database (stage 768, load factor 0.3)
mkdir -p gpurun_out
for P in 8 4; do
( timeout 900 python bench.py --merged-parts $P --no-e2e --steps 5 2>gpurun_out/bench_v_$P.err | tail -1 ) > gpurun_out/bench_v_$P.log
python - <<PY
import json
d=json.load(open("gpurun_out/bench_v_$P.log"))
r=d["roofline"]
print("parts $P:", round(d["value"]/1e6,1), "M reads/s/GPU", round(d["ms_per_step"],2), "ms", r["stage_ms_per_step"], r["per_read"], d["config"]["db"])
PY
tail -2 gpurun_out/bench_v_$P.err | cut -c1-300
done
( timeout 600 python -m pytest tests/test_gpu_shard.py tests/test_gpu_parity.py -m gpu -q 2>&1 | tail -3 ) > gpurun_out/pytest_v.log; tail -2 gpurun_out/pytest_v.log
