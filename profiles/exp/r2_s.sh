# round 2, call S (4 GPUs): merged-table test, --replicate-merged at N=2 and N=4
mkdir -p gpurun_out
( CUDA_VISIBLE_DEVICES=0 timeout 600 python -m pytest tests/test_gpu_shard.py -m gpu -q 2>&1 | tail -6 ) > gpurun_out/pytest_shard_s.log
tail -3 gpurun_out/pytest_shard_s.log
run () {  # name nproc extra
  ( timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $2 --steps 5 --warmup 3 --replicate-merged --no-e2e $3 2>gpurun_out/bench_$1.err | grep '^{"metric"' | tail -1 ) > gpurun_out/bench_$1.log
  python - <<PY
import json
d=json.load(open("gpurun_out/bench_$1.log"))
r=d["roofline"]
print("$1", round(d["value"]/1e6,1), "M reads/s", round(d["ms_per_step"],2), "ms", r["stage_ms_per_step"], r["per_read"], d["config"]["db"], r["queries_fused_warp"], r["queries_cta_smem"])
PY
  tail -2 gpurun_out/bench_$1.err | cut -c1-300
}
run n2_merged_s 2 ""
run n4_merged_s 4 ""
run n4_merged_t1024_s 4 "--table-slots 1024"
