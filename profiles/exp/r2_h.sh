# round 2, call H (4 GPUs): NCCL parity tests (2 and 4 ranks), feature- vs target-sharded at N=4, feature-sharded at N=2
mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -x -q 2>&1 | tail -8 ) > gpurun_out/pytest_multi_h.log
tail -4 gpurun_out/pytest_multi_h.log
run () {  # name nproc extra-args
  ( timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $2 --steps 5 --warmup 3 $3 2>gpurun_out/bench_$1.err | grep '^{"metric"' | tail -1 ) > gpurun_out/bench_$1.log
  python - <<PY
import json
d=json.load(open("gpurun_out/bench_$1.log"))
r=d["roofline"]
print("$1", round(d["value"]/1e6,1), "M reads/s", round(d["ms_per_step"],2), "ms e2e", round(d["e2e"]["value"]/1e6,1), r.get("phase_ms_per_step"), r.get("per_read"))
PY
  tail -2 gpurun_out/bench_$1.err | cut -c1-300
}
run n4_feature_h 4 "--shard-by feature"
run n4_target_h 4 "--shard-by target"
run n2_feature_h 2 "--shard-by feature"
