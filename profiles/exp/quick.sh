# the check every kernel change of the round went through: GPU parity tests, then one bench summary line
# per configuration.  usage: bash profiles/exp/quick.sh [C2|C3|both] [extra bench.py arguments]
mkdir -p gpurun_out
WHAT=${1:-C2}; shift
( timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 ) > gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
summary() { python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']
print('%s value %.2fM e2e %.2fM ms %.2f | frac %.3f of floor %.3f | stage %s | fused %d cta %d' % (sys.argv[1], d['value']/1e6, d['e2e']['value']/1e6, d['ms_per_step'], r['frac'], r['random_access']['frac_of_floor'], r['stage_ms_per_step'], r['queries_fused_warp'], r['queries_cta_smem']))" $1; }
if [ "$WHAT" != "C3" ]; then timeout 600 python bench.py --steps 3 --no-cpu-baseline "$@" 2>&1 | tail -1 | summary C2; fi
if [ "$WHAT" != "C2" ]; then timeout 600 python bench.py --workload C3 --steps 3 --no-cpu-baseline "$@" 2>&1 | tail -1 | summary C3; fi
