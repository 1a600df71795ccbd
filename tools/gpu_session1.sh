#!/bin/bash
# round-2 GPU session 1: first hardware numbers for BASELINE configs 3/4/5 + baseline of the env kernel
mkdir -p gpurun_out/s1
cd /root/repo
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/s1/smi.txt 2>&1
timeout 300 python tools/quick_bench.py 4096 20 > gpurun_out/s1/quick_4096.txt 2>&1
WPB=15 timeout 300 python tools/quick_bench.py 4096 20 >> gpurun_out/s1/quick_4096.txt 2>&1
timeout 300 python tools/quick_bench.py 8192 10 >> gpurun_out/s1/quick_4096.txt 2>&1
timeout 400 python tools/algo_bench.py td3 200 > gpurun_out/s1/td3.json 2> gpurun_out/s1/td3.err
timeout 400 python tools/algo_bench.py ars 2 > gpurun_out/s1/ars.json 2> gpurun_out/s1/ars.err
timeout 400 python bench.py --env CassieTraj-v0 --envs 8192 --precision bf16 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/s1/cfg4_1gpu.json 2> gpurun_out/s1/cfg4.err
timeout 300 python tools/phase_times.py 64 > gpurun_out/s1/phase_times.txt 2>&1
tail -n 3 gpurun_out/s1/*.txt gpurun_out/s1/*.json gpurun_out/s1/*.err
