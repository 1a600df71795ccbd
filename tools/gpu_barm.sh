#!/bin/bash
# sweep of the CTA synchronisation modes of the step kernel (apex_cassie_set_barrier_mask) on the kernel micro-benchmark
tag=${1:-barm}; shift
mkdir -p gpurun_out/$tag
: > gpurun_out/$tag/quick.txt
for m in "$@"; do
  echo "BARM=$m WPB=${WPB:-default}" >> gpurun_out/$tag/quick.txt
  BARM=$m timeout 300 python tools/quick_bench.py 4096 20 >> gpurun_out/$tag/quick.txt 2>&1
done
cat gpurun_out/$tag/quick.txt
