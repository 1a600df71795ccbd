#!/bin/bash
tag=${1:-bench}; steps=${2:-2}
mkdir -p gpurun_out/$tag
timeout 900 python bench.py --steps $steps --warmup 3 > gpurun_out/$tag/bench.json 2> gpurun_out/$tag/bench.err
tail -c 600 gpurun_out/$tag/bench.err; python - <<PY
import json
d=json.loads(open('gpurun_out/$tag/bench.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, d['e2e']['value'], d['roofline']['kernel_ms'], d['roofline']['kernel_share_of_step'])
for k in ('cfg3','cfg4','cfg5','cpu_baseline'): print(k, json.dumps(d.get(k))[:700])
PY
