mkdir -p gpurun_out
timeout 900 python profiles/exp/file_e2e.py 10000000 > gpurun_out/file_e2e.log 2>&1
tail -1 gpurun_out/file_e2e.log | cut -c1-3000
