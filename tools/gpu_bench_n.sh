#!/bin/bash
# bench.py under torchrun on N GPUs (the driver's launch line), short run
n=${1:-2}; steps=${2:-2}
mkdir -p gpurun_out/bn$n
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n --steps $steps --warmup 3 > gpurun_out/bn$n/bench.json 2> gpurun_out/bn$n/bench.err
tail -c 800 gpurun_out/bn$n/bench.err
python - <<PY
import json
d=json.loads(open('gpurun_out/bn$n/bench.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','n_gpus')}, d['e2e']['value'], d.get('nccl'))
for k in ('cfg3','cfg4','cfg5'): print(k, json.dumps(d.get(k))[:400])
PY
