mkdir -p gpurun_out
run() { echo "== $*"; env $1 timeout 300 python bench.py --no-cpu-baseline --steps 3 ${@:2} 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']
print('value %.1fM e2e %.1fM ms %.2f | stage %s | sectors/read %.1f fused %d cta %d' % (d['value']/1e6, d['e2e']['value']/1e6, d['ms_per_step'], r['stage_ms_per_step'], r['per_read']['table_sectors_32B'], r['queries_fused_warp'], r['queries_cta_smem']))"; }
run MCB200_L2_FETCH=64
run MCB200_L2_FETCH=32
run MCB200_L2_FETCH=128
run MCB200_L2_FETCH=32 --table-slots 128
run MCB200_L2_FETCH=32 --table-slots 512
run MCB200_L2_FETCH=32 --load-factor 0.25
run MCB200_L2_FETCH=32 --load-factor 0.8
