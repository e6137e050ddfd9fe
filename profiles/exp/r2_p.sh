# round 2, call P (2 GPUs): parity after the table-mode single-hit filter + pinned loader + NVTX; N=1 and N=2 (target) numbers
mkdir -p gpurun_out
( CUDA_VISIBLE_DEVICES=0 timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_gpu_scale.py tests/test_gpu_dropin.py -m gpu -q --durations=4 2>&1 | tail -12 ) > gpurun_out/pytest_gpu_p.log
tail -6 gpurun_out/pytest_gpu_p.log
( CUDA_VISIBLE_DEVICES=0 timeout 600 python bench.py --no-e2e --steps 5 2>gpurun_out/bench_p.err | tail -1 ) > gpurun_out/bench_p.log
python - <<PY
import json
d=json.load(open("gpurun_out/bench_p.log"))
r=d["roofline"]
print("N1:", round(d["value"]/1e6,1), "M reads/s", round(d["ms_per_step"],2), "ms; kernel", r["kernel_ms_per_launch"], "frac", r["frac"])
PY
( timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 --no-e2e 2>gpurun_out/bench_n2_p.err | grep '^{"metric"' | tail -1 ) > gpurun_out/bench_n2_p.log
python - <<PY
import json
d=json.load(open("gpurun_out/bench_n2_p.log"))
print("N2 target:", round(d["value"]/1e6,1), "M reads/s", round(d["ms_per_step"],2), "ms", d["roofline"]["stage_ms_per_step"])
PY
tail -2 gpurun_out/bench_n2_p.err | cut -c1-300
