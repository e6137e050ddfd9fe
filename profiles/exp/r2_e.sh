# round 2, call E (2 GPUs): feature-sharded bench, two chunk sizes, no e2e noise needed
mkdir -p gpurun_out
for cr in 1250000 2500000; do
( timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 --chunk-reads $cr 2>gpurun_out/bench_n2_e_$cr.err | tail -1 ) > gpurun_out/bench_n2_e_$cr.log
python - <<PY
import json
d=json.load(open("gpurun_out/bench_n2_e_$cr.log"))
print($cr, round(d["value"]/1e6,1), "M reads/s", round(d["ms_per_step"],2), "ms", "e2e", round(d["e2e"]["value"]/1e6,1), d["roofline"]["phase_ms_per_step"], d["roofline"]["stage_ms_per_step"])
PY
tail -3 gpurun_out/bench_n2_e_$cr.err
done
