#!/bin/bash
# f32 parity report (tests/test_env_gpu.py::test_env_f32_parity_report) for the default library and the named variants
mkdir -p gpurun_out/pv
for v in default "$@"; do
  if [ "$v" = default ]; then unset APEX_B200_LIB; else export APEX_B200_LIB=$PWD/build/variants/lib_$v.so; fi
  timeout 300 python -m pytest tests/test_env_gpu.py -x -q -s -k f32_parity_report 2>&1 | grep "f32 parity report" | sed "s/^/$v /" | tee -a gpurun_out/pv/report.txt
done
