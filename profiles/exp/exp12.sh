mkdir -p gpurun_out
( timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 ) > gpurun_out/pytest_gpu.log
tail -8 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']
print('C2 value %.1fM e2e %.1fM ms %.2f | frac %.3f ra %.3f | stage %s | fused %d cta %d' % (d['value']/1e6, d['e2e']['value']/1e6, d['ms_per_step'], r['frac'], r['random_access']['frac_of_floor'], r['stage_ms_per_step'], r['queries_fused_warp'], r['queries_cta_smem']))"
