#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out
show() { python - <<PY
import json
try:
    d=json.load(open("$1")); g=d.get("cuda_graph",{}); print("$2", round(d["value"]/1e6,2), "M  frac", round(d["roofline"]["frac"],4), "launch_us", round(d["roofline"]["mean_launch_ms"]*1e3,2))
except Exception as ex: print("$2 failed", ex)
PY
}
timeout 200 python bench.py --steps 1000 --warmup 20 --no-cpu > $O/x8_c3.json 2> $O/x8_c3.err; show $O/x8_c3.json "c3 early"
MDX_NO_EARLY=1 timeout 200 python bench.py --steps 1000 --warmup 20 --no-cpu > $O/x8_c3_ne.json 2> $O/x8_c3_ne.err; show $O/x8_c3_ne.json "c3 no-early"
timeout 200 python bench.py --steps 1000 --warmup 20 --no-cpu > $O/x8_c3b.json 2> $O/x8_c3b.err; show $O/x8_c3b.json "c3 early (2)"
timeout 600 python -m pytest tests -m gpu -x -q --timeout 120 2>&1 | tail -4
for wl in c5 c4 c3k100; do
timeout 200 python bench.py --workload $wl --steps 300 --warmup 10 --no-cpu > $O/x8_$wl.json 2> $O/x8_$wl.err; show $O/x8_$wl.json "$wl early"
MDX_NO_EARLY=1 timeout 200 python bench.py --workload $wl --steps 300 --warmup 10 --no-cpu > $O/x8_${wl}_ne.json 2> $O/x8_${wl}_ne.err; show $O/x8_${wl}_ne.json "$wl no-early"
done
