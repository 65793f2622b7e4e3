#!/bin/bash
O=gpurun_out
T=${TAG:-r2v}
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_r2.py tests/test_gpu_next.py -x -q -m gpu > $O/${T}_tests.log 2>&1; echo "tests rc=$?"; tail -3 $O/${T}_tests.log
for wl in c5 c4 c3k100; do
  for sp in "" "--no-specialize"; do
    timeout 300 python bench.py --workload $wl $sp --steps 60 --warmup 5 --no-cpu --no-secondary --sustained-episodes 0 --no-feat-host > $O/${T}_$wl$sp.json 2> $O/${T}_$wl$sp.err
    python - <<PY
import json
d=json.loads(open("$O/${T}_$wl$sp.json").read().strip().splitlines()[-1])
print("$wl $sp", round(d["value"]/1e6,2), "frac", round(d["roofline"]["frac"],4), "single", round(d["roofline"]["single_launch"]["frac"],4), d["config"]["kernel_build"][:30])
PY
  done
done
