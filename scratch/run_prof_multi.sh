#!/bin/bash
# full ncu captures (with source) of the run-time-shape variants: C5, C4, C3-K100, and the IDQ featuriser
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out
T=${TAG:-r1j}
ncu --set full --clock-control none --import-source on -k regex:mdx_env_kernel -s 120 -c 1 -f -o $O/prof_${T}_c5 python bench.py --workload c5 --steps 130 --warmup 3 --no-cpu > $O/ncu_${T}_c5.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:mdx_env_kernel -s 120 -c 1 -f -o $O/prof_${T}_c4 python bench.py --workload c4 --steps 130 --warmup 3 --no-cpu > $O/ncu_${T}_c4.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:mdx_env_kernel -s 120 -c 1 -f -o $O/prof_${T}_k100 python bench.py --workload c3k100 --steps 130 --warmup 3 --no-cpu > $O/ncu_${T}_k100.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:mdx_alt_featurize -s 60 -c 1 -f -o $O/prof_${T}_idq python bench.py --workload idq --steps 70 --warmup 3 --no-cpu > $O/ncu_${T}_idq.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:mdx_gae -s 3 -c 1 -f -o $O/prof_${T}_gae python profiles/bench_gae.py > $O/ncu_${T}_gae.log 2>&1
for n in 131072 262144; do python bench.py --envs-per-gpu $n --steps 300 --warmup 10 --no-cpu > $O/${T}_c3_n$n.json 2> $O/${T}_c3_n$n.err; done
ls -la $O/prof_${T}_*.ncu-rep
python - <<PY
import json
for n in (131072, 262144):
    d=json.load(open("$O/${T}_c3_n%d.json"%n)); print(n, round(d["value"]/1e6,2), "M frac", round(d["roofline"]["frac"],4), d["roofline"]["mean_launch_ms"])
PY
