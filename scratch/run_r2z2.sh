#!/bin/bash
O=gpurun_out
T=r2z2
MDX_BULK_OBS=4 timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_r2.py -x -q -m gpu > $O/${T}_tests.log 2>&1; echo "tests rc=$?"; tail -3 $O/${T}_tests.log
for wl in c5 c3 c4 c3k100; do
  for bo in 0 1 4 3; do
    MDX_BULK_OBS=$bo timeout 300 python bench.py --workload $wl --steps 60 --warmup 5 --no-cpu --no-secondary --sustained-episodes 0 --no-feat-host > $O/${T}_${wl}_b$bo.json 2> $O/${T}_${wl}_b$bo.err
    python - <<PY
import json
d=json.loads(open("$O/${T}_${wl}_b$bo.json").read().strip().splitlines()[-1])
print("$wl bulk=$bo", round(d["value"]/1e6,2), "ms", round(d["ms_per_step"],4), "frac", round(d["roofline"]["frac"],4), "single ms", round(d["roofline"]["single_launch"]["mean_launch_ms"],4))
PY
  done
done
