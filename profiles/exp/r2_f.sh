# round 2, call F (2 GPUs): feature-sharded bench with ONE stream (true per-operation times)
mkdir -p gpurun_out
( NCCL_DEBUG=INFO NCCL_DEBUG_SUBSYS=INIT,P2P NCCL_DEBUG_FILE=gpurun_out/nccl_f.%p.log timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 --shard-streams 1 --chunk-reads 2500000 2>gpurun_out/bench_n2_f.err | grep '^{"metric"' | tail -1 ) > gpurun_out/bench_n2_f.log
python - <<PY
import json
d=json.load(open("gpurun_out/bench_n2_f.log"))
print(round(d["value"]/1e6,1), "M reads/s", round(d["ms_per_step"],2), "ms", d["roofline"]["phase_ms_per_step"], d["roofline"]["stage_ms_per_step"])
PY
cat gpurun_out/nccl_f.*.log | grep -iE "P2P|NVLS|via|Connected" | head -8
