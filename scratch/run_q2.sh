#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out
python -m pytest tests/test_gpu_parity.py tests/test_gpu_scale.py -x -q 2>&1 | tail -4 > $O/q2_tests.log
for wl in c3k100 c4 c5; do
python bench.py --workload $wl --steps 600 --warmup 10 --no-cpu > $O/q2_$wl.json 2> $O/q2_$wl.err
done
ncu --set full --clock-control none --import-source on -k regex:mdx_env_kernel -s 200 -c 1 -f -o $O/prof_q2_k100 python bench.py --workload c3k100 --steps 220 --warmup 3 --no-cpu > $O/ncu_q2_k100.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:mdx_env_kernel -s 200 -c 1 -f -o $O/prof_q2_c4 python bench.py --workload c4 --steps 220 --warmup 3 --no-cpu > $O/ncu_q2_c4.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:mdx_env_kernel -s 400 -c 1 -f -o $O/prof_q2_c5 python bench.py --workload c5 --steps 420 --warmup 3 --no-cpu > $O/ncu_q2_c5.log 2>&1
cat $O/q2_tests.log; python - <<PY
import json
for wl in ("c3k100","c4","c5"):
    d=json.load(open("$O/q2_%s.json"%wl)); print(wl, round(d["value"]/1e6,2), "M frac", round(d["roofline"]["frac"],4))
PY
