# round 2, 8 GPUs: NCCL parity tests at 2/4/8 ranks, default bench (target-sharded), feature-sharded, C5 as specified
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -q 2>&1 | tail -6 ) > gpurun_out/pytest_multi_n8.log
tail -3 gpurun_out/pytest_multi_n8.log
run () {  # name extra-args
  ( timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 5 --warmup 3 $2 2>gpurun_out/bench_$1.err | grep '^{"metric"' | tail -1 ) > gpurun_out/bench_$1.log
  python - <<PY
import json
d=json.load(open("gpurun_out/bench_$1.log"))
r=d["roofline"]
print("$1", round(d["value"]/1e6,1), "M reads/s", round(d["ms_per_step"],2), "ms e2e", round(d["e2e"]["value"]/1e6,1), r.get("phase_ms_per_step"), r.get("per_read"), d["config"]["db"])
PY
  tail -2 gpurun_out/bench_$1.err | cut -c1-300
}
run n8_target_r2 ""
run n8_feature_r2 "--shard-by feature"
run n8_c5_r2 "--targets 105000 --reads 12500000"
