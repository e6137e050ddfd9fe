# round 2, call T (4 GPUs): parity (filter instantiation for merged tables), --replicate-merged at N=2 and N=4, C2 N=1 sanity
mkdir -p gpurun_out
( CUDA_VISIBLE_DEVICES=0 timeout 900 python -m pytest tests/test_gpu_shard.py tests/test_gpu_parity.py tests/test_gpu_scale.py -m gpu -q 2>&1 | tail -6 ) > gpurun_out/pytest_t.log
tail -3 gpurun_out/pytest_t.log
( CUDA_VISIBLE_DEVICES=0 timeout 600 python bench.py --no-e2e --steps 5 2>gpurun_out/bench_t.err | tail -1 ) > gpurun_out/bench_t.log
python - <<PY
import json
d=json.load(open("gpurun_out/bench_t.log"))
r=d["roofline"]
print("N1:", round(d["value"]/1e6,1), "M reads/s", round(d["ms_per_step"],2), "ms; kernel", r["kernel_ms_per_launch"])
PY
run () {  # name nproc extra
  ( timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $2 --steps 5 --warmup 3 --replicate-merged $3 2>gpurun_out/bench_$1.err | grep '^{"metric"' | tail -1 ) > gpurun_out/bench_$1.log
  python - <<PY
import json
d=json.load(open("gpurun_out/bench_$1.log"))
r=d["roofline"]
print("$1", round(d["value"]/1e6,1), "M reads/s", round(d["ms_per_step"],2), "ms e2e", d["e2e"] and round(d["e2e"]["value"]/1e6,1), r["stage_ms_per_step"], r["per_read"], r["queries_fused_warp"], r["queries_cta_smem"])
PY
  tail -2 gpurun_out/bench_$1.err | cut -c1-300
}
run n2_merged_t 2 "--no-e2e"
run n4_merged_t 4 ""
