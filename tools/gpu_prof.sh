#!/bin/bash
# one ncu --set full capture of the env-step kernel (source-level stall samples) on the kernel micro-benchmark;
# the library that was profiled is kept next to the report (tools/ncu_lines.py / ncu_ops.py need the matching SASS)
tag=${1:-prof}
mkdir -p gpurun_out/$tag
cp apex_b200/libapex_b200.so gpurun_out/$tag/lib.so
timeout 240 ncu --set full --clock-control none --import-source on -k regex:k_env_step -s 5 -c 1 -f -o gpurun_out/$tag/envstep \
  python tools/quick_bench.py 4096 4 > gpurun_out/$tag/ncu.log 2>&1
tail -n 2 gpurun_out/$tag/ncu.log
