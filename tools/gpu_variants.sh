#!/bin/bash
# kernel micro-benchmark of the default library and of the named build/variants (no parity tests: experiments)
tag=${1:-var}; shift
mkdir -p gpurun_out/$tag
echo "default" > gpurun_out/$tag/quick.txt
timeout 300 python tools/quick_bench.py 4096 20 >> gpurun_out/$tag/quick.txt 2>&1
for v in "$@"; do
  echo "variant $v" >> gpurun_out/$tag/quick.txt
  APEX_B200_LIB=$PWD/build/variants/lib_$v.so timeout 300 python tools/quick_bench.py 4096 20 >> gpurun_out/$tag/quick.txt 2>&1
done
cat gpurun_out/$tag/quick.txt
