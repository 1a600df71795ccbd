#!/bin/bash
tag=${1:-wpb}; shift
mkdir -p gpurun_out/$tag
: > gpurun_out/$tag/quick.txt
for w in "$@"; do
  echo "WPB=$w" >> gpurun_out/$tag/quick.txt
  WPB=$w timeout 300 python tools/quick_bench.py 4096 20 >> gpurun_out/$tag/quick.txt 2>&1
done
echo "BALANCE=0" >> gpurun_out/$tag/quick.txt
BALANCE=0 timeout 300 python tools/quick_bench.py 4096 20 >> gpurun_out/$tag/quick.txt 2>&1
cat gpurun_out/$tag/quick.txt
