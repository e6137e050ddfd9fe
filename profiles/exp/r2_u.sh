# round 2, call U (8 GPUs): default bench at N=8 (auto -> all parts merged on every GPU)
mkdir -p gpurun_out
( timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 5 --warmup 3 2>gpurun_out/bench_n8_u.err | grep '^{"metric"' | tail -1 ) > gpurun_out/bench_n8_u.log
python - <<PY
import json
d=json.load(open("gpurun_out/bench_n8_u.log"))
r=d["roofline"]
print("N8 default:", round(d["value"]/1e6,1), "M reads/s", round(d["ms_per_step"],2), "ms e2e", d["e2e"] and round(d["e2e"]["value"]/1e6,1), r["stage_ms_per_step"], r["per_read"], d["config"]["db"], d["config"]["parallelism"])
PY
tail -4 gpurun_out/bench_n8_u.err | cut -c1-400
nvidia-smi --query-gpu=memory.used --format=csv | head -3
