#!/bin/bash
mkdir -p gpurun_out/modes
for m in 3 0; do
  timeout 300 python bench.py --steps 3 --warmup 3 --no-extras --no-cpu-baseline --tc-mode $m > gpurun_out/modes/bench_tc$m.json 2> gpurun_out/modes/bench_tc$m.err
  tail -c 300 gpurun_out/modes/bench_tc$m.err
  python - <<PY
import json
d=json.loads(open('gpurun_out/modes/bench_tc$m.json').read().strip().splitlines()[-1])
print('tc_mode $m', {k:d[k] for k in ('value','ms_per_step','gpu_launches')}, 'e2e', d['e2e']['value'], 'learner', d['learner']['update_ms'], 'rollout', d['rollout_only']['ms'], 'kernel_ms', d['roofline']['kernel_ms'])
PY
done
