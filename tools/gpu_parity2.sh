#!/bin/bash
mkdir -p gpurun_out/par2
APEX_B200_LIB=$PWD/build/variants/lib_plain.so timeout 600 python -m pytest tests/test_env_gpu.py -x -q -s -k f32_parity_report > gpurun_out/par2/plain.txt 2>&1
cp gpurun_out/parity_f32.json gpurun_out/par2/parity_f32_plain.json
timeout 1200 python -m pytest tests/test_env_gpu.py -x -q -s > gpurun_out/par2/pytest_env.txt 2>&1
tail -n 5 gpurun_out/par2/pytest_env.txt; grep "full-size" gpurun_out/par2/pytest_env.txt
