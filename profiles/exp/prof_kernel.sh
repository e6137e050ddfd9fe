# one `ncu --set full` capture of one kernel on a 1 M-read launch: prof_kernel.sh <regex> <launch-skip> <out tag>
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:$1 -s $2 -c 1 -f -o gpurun_out/prof_$3 \
    python bench.py --reads 1000000 --slot-reads 1000000 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/prof_$3.log 2>&1
tail -c 200 gpurun_out/prof_$3.log
