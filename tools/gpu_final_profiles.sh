#!/bin/bash
# round-2 evidence: ncu --set full of the env-step kernel, FP-op counters, launch list of a short bench step, TD3 replay gather
tag=${1:-r02}
mkdir -p gpurun_out/$tag
cp apex_b200/libapex_b200.so gpurun_out/$tag/lib.so
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_env_step -s 5 -c 1 -f -o gpurun_out/$tag/envstep \
  python tools/quick_bench.py 4096 4 > gpurun_out/$tag/ncu_envstep.log 2>&1
timeout 600 ncu --clock-control none -k regex:k_env_step -s 5 -c 1 --csv --metrics \
smsp__sass_thread_inst_executed_op_fadd_pred_on.sum,smsp__sass_thread_inst_executed_op_fmul_pred_on.sum,smsp__sass_thread_inst_executed_op_ffma_pred_on.sum,smsp__sass_thread_inst_executed_op_dadd_pred_on.sum,smsp__sass_thread_inst_executed_op_dmul_pred_on.sum,smsp__sass_thread_inst_executed_op_dfma_pred_on.sum,gpu__time_duration.sum,smsp__inst_executed.sum \
  python tools/quick_bench.py 4096 4 > gpurun_out/$tag/fpops.csv 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 1400 -c 1400 --csv --log-file gpurun_out/$tag/launches.csv \
  python bench.py --steps 1 --warmup 1 --horizon 32 --no-cpu-baseline --no-extras > gpurun_out/$tag/bench_under_ncu.log 2>&1
timeout 600 ncu --set full --clock-control none -k regex:k_replay_gather -s 8 -c 1 -f -o gpurun_out/$tag/replay_gather \
  python tools/algo_bench.py td3 20 > gpurun_out/$tag/ncu_td3.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 600 -c 130 --csv --log-file gpurun_out/$tag/td3_launches.csv \
  python tools/algo_bench.py td3 20 > gpurun_out/$tag/td3_under_ncu.log 2>&1
ls -la gpurun_out/$tag
