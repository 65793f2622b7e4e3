#!/bin/bash
O=gpurun_out
timeout 300 python bench.py --workload c5step --steps 200 --warmup 20 --no-cpu --no-secondary --sustained-episodes 0 > $O/r2w_c5step.json 2> $O/r2w_c5step.err; echo rc=$?
timeout 300 python bench.py --workload idq --steps 100 --warmup 5 --no-secondary > $O/r2w_idq.json 2> $O/r2w_idq.err; echo rc=$?
timeout 300 python bench.py --workload c2 --steps 400 --warmup 20 --no-cpu --no-secondary --graph > $O/r2w_c2.json 2> $O/r2w_c2.err; echo rc=$?
python - <<PY
import json
for n in ["c5step","idq","c2"]:
    d=json.loads(open(f"$O/r2w_{n}.json").read().strip().splitlines()[-1])
    print(n, round(d["value"]/1e6,2), "ms", d["ms_per_step"], "frac", d["roofline"]["frac"], "single", d["roofline"]["single_launch"]["frac"], "e2e", d["e2e"]["value"]/1e6, d.get("cuda_graph"))
PY
tail -2 $O/r2w_c5step.err
