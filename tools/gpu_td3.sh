#!/bin/bash
mkdir -p gpurun_out/td3
timeout 300 python -m pytest tests/test_td3_gpu.py -x -q --timeout 120 > gpurun_out/td3/pytest.txt 2>&1
tail -n 15 gpurun_out/td3/pytest.txt
timeout 300 python - <<'PY'
import json, torch, bench
r = bench.extra_cfg3_td3(torch.device("cuda:0"))
print(json.dumps(r["sweep"], indent=0))
json.dump(r, open("gpurun_out/td3/cfg3.json", "w"), indent=1)
PY
