#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out
T=r2n
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29561 bench.py --gpus 8 --workload c4 --steps 40 --warmup 5 --no-cpu > $O/${T}_bench_c4_n8.json 2> $O/${T}_bench_c4_n8.err
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29562 bench.py --gpus 8 --steps 20 --warmup 5 > $O/${T}_bench_n8.json 2> $O/${T}_bench_n8.err
timeout 400 python bench.py --steps 20 --warmup 5 --no-cpu --no-secondary > $O/${T}_bench_n1.json 2> $O/${T}_bench_n1.err
python - <<PY
import json
for f in ("bench_c4_n8","bench_n8","bench_n1"):
    try:
        d=json.loads([l for l in open("$O/${T}_%s.json"%f).read().splitlines() if l.startswith("{")][-1])
        r=d["roofline"]
        print(f, "%.1f M env-steps/s  %.2f G robot-steps/s  frac %.4f  e2e %.1f M  coll %.4f" % (d["value"]/1e6, d["robot_steps_per_s"]/1e9, r["frac"], d["e2e"]["value"]/1e6, d["collective_ms"]), r.get("overlapped_launch"), d["config"]["kernel_build"])
    except Exception as ex: print(f, "failed", ex); print(open("$O/${T}_%s.err"%f).read()[-600:])
PY
