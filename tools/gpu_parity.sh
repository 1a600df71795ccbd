#!/bin/bash
tag=${1:-par}
mkdir -p gpurun_out/$tag
timeout 1200 python -m pytest tests/test_env_gpu.py -x -q -s > gpurun_out/$tag/pytest_env.txt 2>&1
tail -n 12 gpurun_out/$tag/pytest_env.txt
timeout 300 python tools/quick_bench.py 4096 20 2>&1 | tee gpurun_out/$tag/quick.txt
