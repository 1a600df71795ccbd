#!/bin/bash
# env-kernel iteration loop on the GPU box: parity tests of the env kernel, then the kernel micro-benchmark
tag=${1:-q}
mkdir -p gpurun_out/$tag
timeout 600 python -m pytest tests/test_env_gpu.py -x -q > gpurun_out/$tag/pytest_env.txt 2>&1
tail -n 4 gpurun_out/$tag/pytest_env.txt
timeout 300 python tools/quick_bench.py 4096 20 > gpurun_out/$tag/quick.txt 2>&1
timeout 300 python tools/quick_bench.py 8192 10 >> gpurun_out/$tag/quick.txt 2>&1
cat gpurun_out/$tag/quick.txt
