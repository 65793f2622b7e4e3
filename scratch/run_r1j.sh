#!/bin/bash
# r1j measurement pass (run on the GPU box through gpurun)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q --timeout 120 2>&1 | tail -5 > $O/gpu_tests_r1j.log
timeout 300 python bench.py --steps 1000 --warmup 20 > $O/bench_r1j.json 2> $O/bench_r1j.err
timeout 200 python bench.py --workload c2 --graph --steps 1000 --warmup 20 > $O/bench_r1j_c2.json 2> $O/bench_r1j_c2.err
timeout 200 python bench.py --workload c3k100 --steps 600 --warmup 10 > $O/bench_r1j_c3k100.json 2> $O/bench_r1j_c3k100.err
timeout 200 python bench.py --workload c4 --steps 600 --warmup 10 > $O/bench_r1j_c4.json 2> $O/bench_r1j_c4.err
timeout 300 python bench.py --workload c5 --steps 300 --warmup 10 --cpu-envs 1 --cpu-steps 100 > $O/bench_r1j_c5.json 2> $O/bench_r1j_c5.err
MDX_NO_EARLY=1 timeout 200 python bench.py --workload c5 --steps 300 --warmup 10 --no-cpu > $O/bench_r1j_c5_noearly.json 2> $O/bench_r1j_c5_noearly.err
timeout 200 python bench.py --workload greedy --steps 1000 --warmup 20 > $O/bench_r1j_greedy.json 2> $O/bench_r1j_greedy.err
timeout 200 python bench.py --workload idq --steps 600 --warmup 10 > $O/bench_r1j_idq.json 2> $O/bench_r1j_idq.err
timeout 100 python profiles/bench_gae.py > $O/bench_r1j_gae.json 2> $O/bench_r1j_gae.err
timeout 200 python bench.py --impl reference --steps 3 --warmup 3 > $O/bench_r1j_ref.json 2> $O/bench_r1j_ref.err
# ncu: launch list of the default bench command, then one full capture of the headline kernel
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $O/launches_r1j.csv python bench.py --steps 20 --warmup 3 --no-cpu > $O/ncu_launch_r1j.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:mdx_env_kernel -s 200 -c 1 -f -o $O/prof_r1j python bench.py --steps 220 --warmup 3 --no-cpu > $O/ncu_full_r1j.log 2>&1
timeout 200 ncu --set full --clock-control none --import-source on -k regex:mdx_env_kernel -s 60 -c 1 -f -o $O/prof_r1j_idqf python bench.py --workload idq --steps 70 --warmup 3 --no-cpu > $O/ncu_full_r1j_idqf.log 2>&1
timeout 100 ncu --set full --clock-control none --import-source on -k regex:mdx_gae -s 3 -c 1 -f -o $O/prof_r1j_gaetma python profiles/bench_gae.py > $O/ncu_full_r1j_gaetma.log 2>&1
tail -2 $O/gpu_tests_r1j.log; for f in $O/bench_r1j*.json; do echo $f; python - <<PY
import json
try:
    d=json.loads([l for l in open("$f").read().splitlines() if l.startswith("{")][-1])
    print("  ", round(d.get("value",0)/1e6,3), "M", (d.get("roofline") or {}).get("frac"), "e2e", (d.get("e2e") or {}).get("value"), (d.get("e2e") or {}).get("pipelined_value"), "cpu", (d.get("cpu_baseline") or {}).get("value"), (d.get("cpu_baseline_c") or {}).get("value"), d.get("ms_per_launch"))
except Exception as ex: print("  failed", ex)
PY
done
