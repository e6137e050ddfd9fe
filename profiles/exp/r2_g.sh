# round 2, call G (1 GPU): drop-in CLI test, shard tests, default bench (e2e with host timing)
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_dropin.py tests/test_gpu_shard.py -m gpu -x -q --durations=5 2>&1 | tail -30 ) > gpurun_out/pytest_gpu_g.log
tail -12 gpurun_out/pytest_gpu_g.log
( timeout 900 python bench.py 2>gpurun_out/bench_g.err | tail -1 ) > gpurun_out/bench_g.log
python - <<PY
import json
d=json.load(open("gpurun_out/bench_g.log"))
print(round(d["value"]/1e6,1), "M reads/s", round(d["ms_per_step"],2), "ms; e2e", json.dumps(d["e2e"])[:900])
PY
tail -3 gpurun_out/bench_g.err
