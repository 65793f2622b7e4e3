#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out
N=$1
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus $N --steps 600 --warmup 20 > $O/bench_${N}gpu_r1j.json 2> $O/bench_${N}gpu_r1j.err
python - <<PY
import json
d=json.loads([l for l in open("$O/bench_${N}gpu_r1j.json").read().splitlines() if l.startswith("{")][-1])
print("c3 x$N", round(d["value"]/1e6,2), "M frac", round(d["roofline"]["frac"],4), "e2e", round(d["e2e"]["value"]/1e6,1), round(d["e2e"]["pipelined_value"]/1e6,1))
PY
