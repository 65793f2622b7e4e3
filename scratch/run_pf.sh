#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out
for d in 0 1024 2048 4096 5328 8192 16384 0; do
MDX_PF_DIST=$d python bench.py --steps 1000 --warmup 20 --no-cpu > $O/pf_$d.json 2> $O/pf_$d.err
python - <<PY
import json
try:
    d=json.load(open("$O/pf_$d.json")); print("pf=$d", round(d["value"]/1e6,2), "M frac", round(d["roofline"]["frac"],4))
except Exception as ex: print("pf=$d failed", ex)
PY
done
