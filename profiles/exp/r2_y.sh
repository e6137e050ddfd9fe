# round 2, call Y (1 GPU): warps per CTA of the fused kernel: 8 (default) vs 4 vs 16, on C2 and on the 8-part merged table
mkdir -p gpurun_out
cp metacache_b200/libmcb200.so metacache_b200/libmcb200_q8.so.variant
for q in 8 4 16; do
cp metacache_b200/libmcb200_q$q.so.variant metacache_b200/libmcb200.so
( timeout 600 python bench.py --no-e2e --steps 5 2>gpurun_out/bench_y_$q.err | tail -1 ) > gpurun_out/bench_y_$q.log
( timeout 600 python bench.py --no-e2e --steps 5 --merged-parts 8 2>gpurun_out/bench_y8_$q.err | tail -1 ) > gpurun_out/bench_y8_$q.log
python - <<PY
import json
d=json.load(open("gpurun_out/bench_y_$q.log")); e=json.load(open("gpurun_out/bench_y8_$q.log"))
print("warps/CTA $q: C2 kernel", d["roofline"]["kernel_ms_per_launch"], "ms step", round(d["ms_per_step"],2), "| 8-part merged kernel", e["roofline"]["kernel_ms_per_launch"], "ms step", round(e["ms_per_step"],2))
PY
done
cp metacache_b200/libmcb200_q8.so.variant metacache_b200/libmcb200.so
