#!/bin/bash
O=gpurun_out
T=r2y
MDX_BULK_OBS=2 timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_r2.py tests/test_gpu_stale.py -x -q -m gpu > $O/${T}_tests.log 2>&1; echo "tests rc=$?"; tail -3 $O/${T}_tests.log
for wl in c3 c5 c4 c3k100; do
  for bo in 0 1 2; do
    MDX_BULK_OBS=$bo timeout 300 python bench.py --workload $wl --steps 60 --warmup 5 --no-cpu --no-secondary --sustained-episodes 0 --no-feat-host > $O/${T}_${wl}_b$bo.json 2> $O/${T}_${wl}_b$bo.err
    python - <<PY
import json
try:
    d=json.loads(open("$O/${T}_${wl}_b$bo.json").read().strip().splitlines()[-1])
    print("$wl bulk=$bo", round(d["value"]/1e6,2), "frac", round(d["roofline"]["frac"],4), "single", round(d["roofline"]["single_launch"]["frac"],4), d["config"]["kernel_build"][:24])
except Exception as e:
    print("$wl bulk=$bo failed", e); print(open("$O/${T}_${wl}_b$bo.err").read()[-600:])
PY
  done
done
