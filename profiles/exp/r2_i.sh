# round 2, call I (2 GPUs): all GPU tests (kernel change A+C, drop-in, multi-device store, NCCL modes), bench N=1 and N=2 target-sharded
mkdir -p gpurun_out
( timeout 1800 python -m pytest tests -m gpu -q --durations=8 2>&1 | tail -40 ) > gpurun_out/pytest_gpu_i.log
tail -30 gpurun_out/pytest_gpu_i.log
( CUDA_VISIBLE_DEVICES=0 timeout 900 python bench.py 2>gpurun_out/bench_i.err | tail -1 ) > gpurun_out/bench_i.log
python - <<PY
import json
d=json.load(open("gpurun_out/bench_i.log"))
r=d["roofline"]
print("N1", round(d["value"]/1e6,1), "M reads/s", round(d["ms_per_step"],2), "ms; kernel", r["kernel_ms_per_launch"], "frac", r["frac"], "parity", d["parity"], "e2e", json.dumps(d["e2e"])[:700])
PY
tail -2 gpurun_out/bench_i.err | cut -c1-300
( timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 2>gpurun_out/bench_n2_i.err | grep '^{"metric"' | tail -1 ) > gpurun_out/bench_n2_i.log
python - <<PY
import json
d=json.load(open("gpurun_out/bench_n2_i.log"))
print("N2", round(d["value"]/1e6,1), "M reads/s", round(d["ms_per_step"],2), "ms e2e", round(d["e2e"]["value"]/1e6,1))
PY
tail -2 gpurun_out/bench_n2_i.err | cut -c1-300
