#!/bin/bash
mkdir -p gpurun_out/tc3
timeout 150 python -m pytest tests/test_tc3_gpu.py -x -q --timeout 60 > gpurun_out/tc3/pytest.txt 2>&1
tail -n 12 gpurun_out/tc3/pytest.txt
timeout 120 python tools/tc3_bench.py gpurun_out/tc3/bench.json 2>&1 | tail -3
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 40 --csv --log-file gpurun_out/tc3/launches.csv python tools/tc3_profile_target.py 3 > /dev/null 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/tc3/launches.csv')) if len(r)>5 and r[0].isdigit()]
for r in rows[-12:]: print(r[4][:60], r[-1])
PY
timeout 900 python -m pytest tests -m gpu -x -q --timeout 300 > gpurun_out/tc3/pytest_all.txt 2>&1
tail -n 15 gpurun_out/tc3/pytest_all.txt
