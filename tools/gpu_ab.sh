#!/bin/bash
# A/B of kernel builds: default library + every build/variants/lib_*.so given as arguments (names)
tag=${1:-ab}; shift
mkdir -p gpurun_out/$tag
timeout 600 python -m pytest tests/test_env_gpu.py -x -q > gpurun_out/$tag/pytest_env.txt 2>&1
tail -n 3 gpurun_out/$tag/pytest_env.txt
echo "default" > gpurun_out/$tag/quick.txt
timeout 300 python tools/quick_bench.py 4096 20 >> gpurun_out/$tag/quick.txt 2>&1
for v in "$@"; do
  echo "variant $v" >> gpurun_out/$tag/quick.txt
  APEX_B200_LIB=$PWD/build/variants/lib_$v.so timeout 300 python tools/quick_bench.py 4096 20 >> gpurun_out/$tag/quick.txt 2>&1
done
cat gpurun_out/$tag/quick.txt
