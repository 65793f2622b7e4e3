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
timeout 400 python -m pytest tests/test_gpu_parity.py tests/test_gpu_scale.py -q -k "edge_shapes or step_only_large" --timeout 60 2>&1 | tail -15
timeout 600 python -m pytest tests -m gpu -x -q --timeout 120 --deselect tests/test_gpu_parity.py::test_edge_shapes_against_oracle 2>&1 | tail -4
timeout 200 python bench.py --workload c3k100 --steps 600 --warmup 10 --no-cpu > $O/x6_k100.json 2> $O/x6_k100.err; show $O/x6_k100.json "k100 split"
MDX_NO_VEC_SPLIT=1 timeout 200 python bench.py --workload c3k100 --steps 600 --warmup 10 --no-cpu > $O/x6_k100_ns.json 2> $O/x6_k100_ns.err; show $O/x6_k100_ns.json "k100 whole rows"
timeout 200 python bench.py --workload greedy --steps 1000 --warmup 20 --no-cpu > $O/x6_greedy.json 2> $O/x6_greedy.err; show $O/x6_greedy.json "greedy"
timeout 200 python bench.py --workload c4 --steps 600 --warmup 10 --no-cpu > $O/x6_c4.json 2> $O/x6_c4.err; show $O/x6_c4.json "c4"
