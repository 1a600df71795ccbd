#!/bin/bash
# round-2 (second half) evidence: full bench line, launch list of a short bench step, ncu --set full of the tensor-core learner kernels
tag=${1:-r02b}
mkdir -p gpurun_out/$tag
cp apex_b200/libapex_b200.so gpurun_out/$tag/lib.so
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/$tag/bench.json 2> gpurun_out/$tag/bench.err
tail -c 400 gpurun_out/$tag/bench.err
python - <<PY
import json
d=json.loads(open('gpurun_out/$tag/bench.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, d['e2e']['value'], d['roofline']['kernel_ms'], d['roofline']['kernel_share_of_step'], d['learner'])
for k in ('cfg3','cfg4','cfg5','cpu_baseline'): print(k, json.dumps(d.get(k))[:900])
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 1800 -c 1800 --csv --log-file gpurun_out/$tag/launches.csv \
  python bench.py --steps 1 --warmup 1 --horizon 32 --no-cpu-baseline --no-extras > gpurun_out/$tag/bench_under_ncu.log 2>&1
for k in k_tc3_nt k_tc3_tn; do
  timeout 200 ncu --set full --clock-control none --import-source on -k regex:$k -s 4 -c 1 -f -o gpurun_out/$tag/$k \
    python tools/tc3_profile_target.py 3 > gpurun_out/$tag/ncu_$k.log 2>&1
  tail -n 1 gpurun_out/$tag/ncu_$k.log
done
ls gpurun_out/$tag
