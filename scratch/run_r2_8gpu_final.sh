#!/bin/bash
O=gpurun_out
T=r2f8
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29561 bench.py --gpus 8 --steps 20 --warmup 5 > $O/${T}_bench_n8.json 2> $O/${T}_bench_n8.err
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29562 bench.py --gpus 8 --workload c5 --steps 40 --warmup 5 --no-cpu --sustained-episodes 0 > $O/${T}_c5_n8.json 2> $O/${T}_c5_n8.err
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29563 bench.py --gpus 8 --workload c4 --steps 40 --warmup 5 --no-cpu --sustained-episodes 0 > $O/${T}_c4_n8.json 2> $O/${T}_c4_n8.err
python - <<PY
import json
for n in ["bench_n8","c5_n8","c4_n8"]:
    try:
        d=json.loads([l for l in open("$O/${T}_%s.json"%n).read().splitlines() if l.startswith("{")][-1])
        print(n, "value %.1f M  ms/step %.4f  coll_ms %.4f  frac %.4f  e2e %.1f M" % (d["value"]/1e6, d["ms_per_step"], d["collective_ms"], d["roofline"]["frac"], d["e2e"]["value"]/1e6), d["config"].get("envs_per_gpu"))
    except Exception as ex: print(n, "failed", ex)
PY
