# C3: launch durations of the query kernels + one full capture of the small-tier CTA kernel
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,launch__registers_per_thread,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active \
   --clock-control none -k regex:query_ -s 10 -c 6 --csv --log-file gpurun_out/c3_launches.csv \
   python bench.py --workload C3 --reads 500000 --slot-reads 500000 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/c3_l.log 2>&1
grep -v "^==" gpurun_out/c3_launches.csv | python -c "
import csv,sys
rows=list(csv.reader(sys.stdin)); hi=[i for i,r in enumerate(rows) if r and r[0]=='ID'][0]; idx={h:i for i,h in enumerate(rows[hi])}
for r in rows[hi+1:]:
    if len(r)>=len(rows[hi]): print(r[idx['ID']], r[idx['Kernel Name']][:40], r[idx['Grid Size']], r[idx['Metric Name']], r[idx['Metric Value']])
"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:query_heavy_kernel -s 6 -c 1 -f -o gpurun_out/prof_heavy_c3 \
    python bench.py --workload C3 --reads 500000 --slot-reads 500000 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/prof_heavy.log 2>&1
tail -c 200 gpurun_out/prof_heavy.log
