#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --workload c4 --steps 400 --warmup 10 > $O/bench_8gpu_r1j_c4.json 2> $O/bench_8gpu_r1j_c4.err
python - <<PY
import json
d=json.loads([l for l in open("$O/bench_8gpu_r1j_c4.json").read().splitlines() if l.startswith("{")][-1])
print("c4 x8", round(d["value"]/1e6,2), "M frac", round(d["roofline"]["frac"],4), d["n_gpus"], d["config"]["envs_total"], d["episode_stats"])
PY
