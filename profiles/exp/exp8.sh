mkdir -p gpurun_out
( timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 ) > gpurun_out/pytest_gpu.log
tail -8 gpurun_out/pytest_gpu.log
run() { echo "== $*"; env $1 timeout 300 python bench.py --no-cpu-baseline --steps 3 ${@:2} 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']
print('value %.1fM e2e %.1fM ms %.2f | frac %.3f | stage %s | ra %s' % (d['value']/1e6, d['e2e']['value']/1e6, d['ms_per_step'], r['frac'], r['stage_ms_per_step'], r['random_access']))"; }
run MCB200_SKETCH_STAGES=1
run MCB200_SKETCH_STAGES=2
run MCB200_SKETCH_STAGES=1 --slot-reads 500000
