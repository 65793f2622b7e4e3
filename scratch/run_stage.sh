#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out
for wl in c5 c4 c3k100; do for sb in 8192 4096 3000 2048; do
  MDX_STAGE_BYTES=$sb python bench.py --workload $wl --steps 600 --warmup 10 --no-cpu > $O/stage_${wl}_$sb.json 2> $O/stage_${wl}_$sb.err
  python - <<PY
import json
try:
    d=json.load(open("$O/stage_${wl}_$sb.json")); print("$wl stage=$sb", round(d["value"]/1e6,2), "M frac", round(d["roofline"]["frac"],4))
except Exception as ex: print("$wl $sb failed", ex)
PY
done; done
