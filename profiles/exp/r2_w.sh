# round 2, call W (2 GPUs): default bench at N=2 (auto -> merged), final code
mkdir -p gpurun_out
( timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 2>gpurun_out/bench_n2_w.err | grep '^{"metric"' | tail -1 ) > gpurun_out/bench_n2_w.log
python - <<PY
import json
d=json.load(open("gpurun_out/bench_n2_w.log"))
r=d["roofline"]
print("N2 default:", round(d["value"]/1e6,1), "M reads/s", round(d["ms_per_step"],2), "ms e2e", d["e2e"] and round(d["e2e"]["value"]/1e6,1), r["stage_ms_per_step"], d["config"]["parallelism"])
PY
tail -3 gpurun_out/bench_n2_w.err | cut -c1-300
( timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --impl reference --gpus 2 --steps 3 --warmup 1 2>&1 | grep '^{"impl"' | tail -1 | cut -c1-300 )
