# round 2, call R (4 GPUs): multi-GPU tests with the final kernels; feature-sharded bench at N=4 after the single-hit filter
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_multi.py tests/test_gpu_parity.py -m gpu -q 2>&1 | tail -6 ) > gpurun_out/pytest_multi_r.log
tail -3 gpurun_out/pytest_multi_r.log
( timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 4 --steps 5 --warmup 3 --shard-by feature 2>gpurun_out/bench_n4_feature_r.err | grep '^{"metric"' | tail -1 ) > gpurun_out/bench_n4_feature_r.log
python - <<PY
import json
d=json.load(open("gpurun_out/bench_n4_feature_r.log"))
print("N4 feature:", round(d["value"]/1e6,1), "M reads/s", round(d["ms_per_step"],2), "ms e2e", round(d["e2e"]["value"]/1e6,1), d["roofline"]["phase_ms_per_step"])
PY
tail -2 gpurun_out/bench_n4_feature_r.err | cut -c1-300
