#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out
for args in "--workload greedy --steps 100" "--workload idq --steps 100" "--workload c2 --steps 200 --graph" "--envs-per-gpu 1001 --steps 30" "--split 1 --steps 30" "--workload c4 --steps 30 --split 3"; do
  tag=$(echo $args | tr -d ' -' | cut -c1-24)
  timeout 300 python bench.py $args --warmup 5 --no-cpu --no-secondary --sustained-episodes 1 --no-feat-host > $O/r2r_$tag.json 2> $O/r2r_$tag.err
  python - <<PY
import json
try:
    d=json.loads([l for l in open("$O/r2r_$tag.json").read().splitlines() if l.startswith("{")][-1])
    print("$args ->", "%.1f M" % (d["value"]/1e6), "frac %.3f" % d["roofline"]["frac"], "e2e %.1f M" % (d["e2e"]["value"]/1e6), "launches/step", d["launches_per_step"], "sustained", (d.get("sustained") or {}).get("value"), (d.get("cuda_graph") or {}).get("ms_per_step"))
except Exception as ex: print("$args FAILED", ex); print(open("$O/r2r_$tag.err").read()[-700:])
PY
done
timeout 200 python examples/mappo_rollout_demo.py --envs 4096 --steps 16 --iters 2 2>&1 | tail -2
