run() { echo "== $*"; env $1 $2 timeout 300 python bench.py --no-cpu-baseline --steps 3 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']
print('value %.1fM (%.2f ms) e2e %.1fM (%.2f ms) | %s' % (d['value']/1e6, d['ms_per_step'], d['e2e']['value']/1e6, d['e2e']['ms_per_step'], r['stage_ms_per_step']))"; }
run MCB200_SKETCH_CTAS=8 MCB200_QUERY_CTAS=0
run MCB200_SKETCH_CTAS=4 MCB200_QUERY_CTAS=4
run MCB200_SKETCH_CTAS=3 MCB200_QUERY_CTAS=4
run MCB200_SKETCH_CTAS=2 MCB200_QUERY_CTAS=5
run MCB200_SKETCH_CTAS=4 MCB200_QUERY_CTAS=0
