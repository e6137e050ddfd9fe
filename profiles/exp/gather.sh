# random-gather micro-benchmark: lookups/s per granule size, then DRAM bytes per lookup under ncu
mkdir -p gpurun_out
./profiles/exp/gather_bench > gpurun_out/gather_bench.log 2>&1
cat gpurun_out/gather_bench.log
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,lts__t_sectors_srcunit_tex_op_read.sum,lts__t_requests_srcunit_tex_op_read.sum \
   --clock-control none -k regex:pat_ --csv --log-file gpurun_out/gather_ncu.csv ./profiles/exp/gather_bench > /dev/null 2>&1
python - <<'PY'
import csv, collections
rows = list(csv.reader(open('gpurun_out/gather_ncu.csv')))
hi = [i for i, r in enumerate(rows) if r and r[0] == 'ID'][0]
idx = {h: i for i, h in enumerate(rows[hi])}
d = collections.OrderedDict()
for r in rows[hi + 1:]:
    if len(r) < len(rows[hi]): continue
    d.setdefault(r[idx['ID']], {'name': r[idx['Kernel Name']]})[r[idx['Metric Name']]] = (float(r[idx['Metric Value']].replace(',', '')), r[idx['Metric Unit']])
for k, v in d.items():
    print(k, v['name'][:40], {m: x for m, x in v.items() if m != 'name'})
PY
