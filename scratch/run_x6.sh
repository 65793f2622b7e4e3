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
python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python bench.py --workload c3k100 --steps 600 --warmup 10 --no-cpu > $O/x6_k100.json 2> $O/x6_k100.err; show $O/x6_k100.json "k100 split"
MDX_NO_VEC_SPLIT=1 python bench.py --workload c3k100 --steps 600 --warmup 10 --no-cpu > $O/x6_k100_ns.json 2> $O/x6_k100_ns.err; show $O/x6_k100_ns.json "k100 whole rows"
python bench.py --workload greedy --steps 1000 --warmup 20 --no-cpu > $O/x6_greedy.json 2> $O/x6_greedy.err; show $O/x6_greedy.json "greedy"
python bench.py --workload c4 --steps 600 --warmup 10 --no-cpu > $O/x6_c4.json 2> $O/x6_c4.err; show $O/x6_c4.json "c4"
