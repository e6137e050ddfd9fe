# round 2: file in -> candidates out (query_file with reader threads) and the CLIs (reference, drop-in) on the same file
mkdir -p gpurun_out
( timeout 1500 python tests/file_e2e.py 2>gpurun_out/file_e2e.err | tail -1 ) > gpurun_out/file_e2e_r2.json
cut -c1-2500 gpurun_out/file_e2e_r2.json; tail -3 gpurun_out/file_e2e.err
