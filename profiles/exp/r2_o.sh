# round 2, call O (2 GPUs): feature-sharded bench after the single-hit filter, pipelined and serial
mkdir -p gpurun_out
for mode in 3 1; do
( timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 --shard-by feature --shard-streams $mode --chunk-reads 2500000 2>gpurun_out/bench_n2_o_$mode.err | grep '^{"metric"' | tail -1 ) > gpurun_out/bench_n2_o_$mode.log
python - <<PY
import json
d=json.load(open("gpurun_out/bench_n2_o_$mode.log"))
print("streams $mode:", round(d["value"]/1e6,1), "M reads/s", round(d["ms_per_step"],2), "ms", d["roofline"]["phase_ms_per_step"])
PY
tail -2 gpurun_out/bench_n2_o_$mode.err | cut -c1-200
done
