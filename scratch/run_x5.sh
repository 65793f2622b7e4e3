#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out
show() { python - <<PY
import json
try:
    d=json.load(open("$1")); g=d.get("cuda_graph",{}); print("$2", round(d["value"]/1e6,2), "M  frac", round(d["roofline"]["frac"],4), "launch_us", round(d["roofline"]["mean_launch_ms"]*1e3,2), "graph", round(g.get("value",0)/1e6,1), round(g.get("ms_per_step",0)*1e3,2))
except Exception as ex: print("$2 failed", ex)
PY
}
for v in fat4 fat5 fat6; do
  MDX_NOFAT=-1 MDX_LIB_PATH=$PWD/scratch/libmdx_$v.so python bench.py --workload c4 --steps 600 --warmup 10 --no-cpu > $O/x5_c4_$v.json 2> $O/x5_c4_$v.err; show $O/x5_c4_$v.json "c4 forced $v"
  MDX_LIB_PATH=$PWD/scratch/libmdx_$v.so python bench.py --workload c5 --steps 300 --warmup 10 --no-cpu > $O/x5_c5_$v.json 2> $O/x5_c5_$v.err; show $O/x5_c5_$v.json "c5 $v"
done
for lpe in 0 8 16; do
  MDX_LPE=$lpe python bench.py --workload c2 --graph --steps 1000 --warmup 20 --no-cpu > $O/x5_c2_lpe$lpe.json 2> $O/x5_c2_lpe$lpe.err; show $O/x5_c2_lpe$lpe.json "c2 4096 lpe$lpe"
  MDX_LPE=$lpe python bench.py --workload c2 --graph --envs-per-gpu 65536 --steps 1000 --warmup 20 --no-cpu > $O/x5_c2big_lpe$lpe.json 2> $O/x5_c2big_lpe$lpe.err; show $O/x5_c2big_lpe$lpe.json "c2 65536 lpe$lpe"
done
