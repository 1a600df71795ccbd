#!/bin/bash
mkdir -p gpurun_out/tests
timeout 2400 python -m pytest tests -m gpu -x -q > gpurun_out/tests/pytest_gpu.txt 2>&1
tail -n 25 gpurun_out/tests/pytest_gpu.txt
