# round 2, call M (1 GPU): parity with two-size CTA hash tiers + single-hit filter (lists mode), C3 number
mkdir -p gpurun_out
( timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_gpu_scale.py tests/test_gpu_shard.py -m gpu -q --durations=5 2>&1 | tail -25 ) > gpurun_out/pytest_gpu_m.log
tail -8 gpurun_out/pytest_gpu_m.log
( timeout 600 python bench.py --workload C3 --no-e2e --steps 5 2>gpurun_out/bench_m_c3.err | tail -1 ) > gpurun_out/bench_m_c3.log
python - <<PY
import json
d=json.load(open("gpurun_out/bench_m_c3.log"))
r=d["roofline"]
print("C3:", round(d["value"]/1e6,2), "M reads/s", round(d["ms_per_step"],2), "ms", r["stage_ms_per_step"], r["queries_fused_warp"], r["queries_cta_smem"], r["queries_cta_global"])
PY
tail -3 gpurun_out/bench_m_c3.err
