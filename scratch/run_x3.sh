#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out
show() { python - <<PY
import json
try:
    d=json.load(open("$1")); print("$2", round(d["value"]/1e6,2), "M  frac", round(d["roofline"]["frac"],4), "launch_us", round(d["roofline"]["mean_launch_ms"]*1e3,2), "e2e", round(d["e2e"]["value"]/1e6,1), "pipe", round(d["e2e"].get("pipelined_value",0)/1e6,1))
except Exception as ex: print("$2 failed", ex)
PY
}
python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python bench.py --workload idq --steps 600 --warmup 10 --no-cpu > $O/x3_idq.json 2> $O/x3_idq.err; show $O/x3_idq.json "idq fused"
MDX_IDQ_UNFUSED=1 python bench.py --workload idq --steps 600 --warmup 10 --no-cpu > $O/x3_idq_unf.json 2> $O/x3_idq_unf.err; show $O/x3_idq_unf.json "idq 2 launches"
python bench.py --workload c5 --steps 300 --warmup 10 --no-cpu > $O/x3_c5.json 2> $O/x3_c5.err; show $O/x3_c5.json "c5"
python bench.py --workload c4 --steps 600 --warmup 10 --no-cpu > $O/x3_c4.json 2> $O/x3_c4.err; show $O/x3_c4.json c4
python bench.py --workload c3k100 --steps 600 --warmup 10 --no-cpu > $O/x3_k100.json 2> $O/x3_k100.err; show $O/x3_k100.json k100
python bench.py --steps 1000 --warmup 20 --no-cpu > $O/x3_c3.json 2> $O/x3_c3.err; show $O/x3_c3.json c3
python profiles/bench_gae.py 2>&1 | tail -1
