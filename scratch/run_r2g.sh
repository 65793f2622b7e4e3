#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out
T=${TAG:-r2g}
timeout 300 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > $O/${T}_ref1.json 2> $O/${T}_ref1.err
timeout 400 python bench.py --gpus 1 --steps 20 --warmup 5 > $O/${T}_bench_n1.json 2> $O/${T}_bench_n1.err
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 20 --warmup 5 > $O/${T}_bench_n2.json 2> $O/${T}_bench_n2.err
python - <<PY
import json
v={}
for n in (1,2):
    try:
        d=json.loads([l for l in open("$O/${T}_bench_n%d.json"%n).read().splitlines() if l.startswith("{")][-1])
        v[n]=d["value"]; print(n, "value %.1f M  ms/step %.4f  coll_ms %.4f  frac %.4f  e2e %.1f M sustained %.1f M" % (d["value"]/1e6, d["ms_per_step"], d["collective_ms"], d["roofline"]["frac"], d["e2e"]["value"]/1e6, d.get("sustained",{}).get("value",0)/1e6), d["config"]["host_cores_per_rank"])
    except Exception as ex: print(n, "failed", ex)
if 1 in v and 2 in v: print("efficiency at 2:", v[2]/(2*v[1]))
PY
tail -c 600 $O/${T}_bench_n2.err
