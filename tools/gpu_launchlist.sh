#!/bin/bash
tag=${1:-r02d}
mkdir -p gpurun_out/$tag
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 3600 -c 1800 --csv --log-file gpurun_out/$tag/launches.csv \
  python bench.py --steps 1 --warmup 2 --horizon 32 --no-cpu-baseline --no-extras > gpurun_out/$tag/bench_under_ncu.log 2>&1
tail -c 300 gpurun_out/$tag/bench_under_ncu.log
wc -l gpurun_out/$tag/launches.csv
