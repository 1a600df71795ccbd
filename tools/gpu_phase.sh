#!/bin/bash
mkdir -p gpurun_out/phase
timeout 300 python -m pytest tests/test_env_gpu.py -x -q --timeout 120 > gpurun_out/phase/pytest_env.txt 2>&1
tail -n 15 gpurun_out/phase/pytest_env.txt
timeout 200 python tools/quick_bench.py 4096 20 2>&1 | tail -3
