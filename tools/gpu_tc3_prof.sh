#!/bin/bash
# ncu --set full captures of the tensor-core learner kernels on the actor's update shape (65536 rows)
mkdir -p gpurun_out/tc3
for k in k_tc3_nt k_tc3_tn; do
  timeout 200 ncu --set full --clock-control none --import-source on -k regex:$k -s 2 -c 1 -f -o gpurun_out/tc3/$k \
    python tools/tc3_profile_target.py 3 > gpurun_out/tc3/ncu_$k.log 2>&1
  tail -n 1 gpurun_out/tc3/ncu_$k.log
done
