#!/bin/bash
mkdir -p gpurun_out/graph
timeout 600 python -m pytest tests/test_ppo_gpu.py tests/test_train_gpu.py -x -q --timeout 200 > gpurun_out/graph/pytest.txt 2>&1
tail -n 12 gpurun_out/graph/pytest.txt
timeout 300 python bench.py --steps 3 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/graph/bench.json 2> gpurun_out/graph/bench.err
tail -c 400 gpurun_out/graph/bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/graph/bench.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, 'e2e', d['e2e']['value'], 'learner', d['learner']['update_ms'], 'rollout', d['rollout_only']['ms'], 'kernel_ms', d['roofline']['kernel_ms'])
PY
