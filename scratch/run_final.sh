#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 300 python bench.py > $O/final_bench.json 2> $O/final_bench.err; python - <<PY
import json
d=json.load(open("$O/final_bench.json")); print("c3", round(d["value"]/1e6,2), "M frac", round(d["roofline"]["frac"],4), "e2e", round(d["e2e"]["value"]/1e6,1), round(d["e2e"]["pipelined_value"]/1e6,1), "cpu", round(d["cpu_baseline"]["value"]), d["cpu_baseline_c"], d["clocks"], d["gpu_launches"])
PY
timeout 100 python bench.py --impl reference | cut -c1-300
