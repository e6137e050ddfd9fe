# round 2, call Z (2 GPUs): default bench at N=2 with the target-sharded step measured beside the merged mode
mkdir -p gpurun_out
( timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 2>gpurun_out/bench_n2_z.err | grep '^{"metric"' | tail -1 ) > gpurun_out/bench_n2_z.log
python - <<PY
import json
d=json.load(open("gpurun_out/bench_n2_z.log"))
print("N2 default:", round(d["value"]/1e6,1), "M reads/s", round(d["ms_per_step"],2), "ms e2e", d["e2e"] and round(d["e2e"]["value"]/1e6,1), d["config"]["parallelism"], "| target-sharded same run:", d.get("target_sharded_same_run"))
PY
grep -v "OMP_NUM\|^\*\*\*" gpurun_out/bench_n2_z.err | tail -3 | cut -c1-300
