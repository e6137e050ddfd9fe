# one `ncu --set full` capture of the fused probe kernel on a 1 M-read launch (tag = $1)
mkdir -p gpurun_out
TAG=${1:-x}
timeout 600 ncu --set full --clock-control none --import-source on -k regex:query_fast_kernel -s 3 -c 1 -f -o gpurun_out/prof_query_$TAG \
    python bench.py --reads 1000000 --slot-reads 1000000 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/prof_query.log 2>&1
tail -c 300 gpurun_out/prof_query.log
