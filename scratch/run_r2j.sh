#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out
T=r2j
for n in 8 4; do
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2953$n bench.py --gpus $n --steps 20 --warmup 5 > $O/${T}_bench_n$n.json 2> $O/${T}_bench_n$n.err
done
timeout 400 python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu --no-secondary > $O/${T}_bench_n1.json 2> $O/${T}_bench_n1.err
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 8 --workload c5 --steps 40 --warmup 5 --no-cpu > $O/${T}_bench_c5_n8.json 2> $O/${T}_bench_c5_n8.err
python - <<PY
import json
v={}
for n in (1,4,8):
    try:
        d=json.loads([l for l in open("$O/${T}_bench_n%d.json"%n).read().splitlines() if l.startswith("{")][-1])
        v[n]=d["value"]; print(n, "value %.1f M  ms/step %.4f  coll_ms %.4f  frac %.4f  e2e %.1f M sustained %.1f M cores/rank %s" % (d["value"]/1e6, d["ms_per_step"], d["collective_ms"], d["roofline"]["frac"], d["e2e"]["value"]/1e6, d.get("sustained",{}).get("value",0)/1e6, d["config"]["host_cores_per_rank"]))
    except Exception as ex: print(n, "failed", ex)
for n in (4,8):
    if 1 in v and n in v: print("efficiency at", n, v[n]/(n*v[1]))
try:
    d=json.loads([l for l in open("$O/${T}_bench_c5_n8.json").read().splitlines() if l.startswith("{")][-1])
    print("C5 x8: %.1f M env-steps/s, %.2f G robot-steps/s, frac %.3f, sustained %s" % (d["value"]/1e6, d["robot_steps_per_s"]/1e9, d["roofline"]["frac"], d.get("sustained",{}).get("value")), d["config"]["kernel_build"])
except Exception as ex: print("c5 n8 failed", ex)
PY
tail -c 400 $O/${T}_bench_n8.err
