#!/bin/bash
O=gpurun_out
T=r2z3
for wl in c5 c4; do
  for bo in 0 5 0 5; do
    MDX_BULK_OBS=$bo timeout 300 python bench.py --workload $wl --steps 60 --warmup 5 --no-cpu --no-secondary --sustained-episodes 0 --no-feat-host > $O/${T}_${wl}_b$bo.json 2> $O/${T}_${wl}_b$bo.err
    python - <<PY
import json
d=json.loads(open("$O/${T}_${wl}_b$bo.json").read().strip().splitlines()[-1])
print("$wl rot=$bo", round(d["value"]/1e6,2), "ms", round(d["ms_per_step"],4), "frac", round(d["roofline"]["frac"],4), "single ms", round(d["roofline"]["single_launch"]["mean_launch_ms"],4))
PY
  done
done
