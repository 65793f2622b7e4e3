#!/bin/bash
# r1h measurement pass (run on the GPU box through gpurun)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > $O/gpu_tests_r1h.log
python bench.py --steps 1000 --warmup 20 > $O/bench_r1h.json 2> $O/bench_r1h.err
python bench.py --workload c2 --graph --steps 1000 --warmup 20 > $O/bench_r1h_c2.json 2> $O/bench_r1h_c2.err
python bench.py --workload c3k100 --steps 600 --warmup 10 > $O/bench_r1h_c3k100.json 2> $O/bench_r1h_c3k100.err
python bench.py --workload c4 --steps 600 --warmup 10 > $O/bench_r1h_c4.json 2> $O/bench_r1h_c4.err
python bench.py --workload c5 --steps 300 --warmup 10 --cpu-envs 1 --cpu-steps 100 > $O/bench_r1h_c5.json 2> $O/bench_r1h_c5.err
python bench.py --workload greedy --steps 1000 --warmup 20 > $O/bench_r1h_greedy.json 2> $O/bench_r1h_greedy.err
python bench.py --workload idq --steps 600 --warmup 10 > $O/bench_r1h_idq.json 2> $O/bench_r1h_idq.err
python profiles/bench_gae.py > $O/bench_r1h_gae.json 2> $O/bench_r1h_gae.err
python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_r1h_ref.json 2> $O/bench_r1h_ref.err
# ncu: launch list of the default bench command, then one full capture of the headline kernel
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $O/launches_r1h.csv python bench.py --steps 20 --warmup 3 --no-cpu > $O/ncu_launch_r1h.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:mdx_env_kernel -s 10 -c 1 -f -o $O/prof_r1h python bench.py --steps 20 --warmup 3 --no-cpu > $O/ncu_full_r1h.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:mdx_env_kernel -s 10 -c 1 -f -o $O/prof_r1h_c5 python bench.py --workload c5 --steps 20 --warmup 3 --no-cpu > $O/ncu_full_r1h_c5.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:mdx_env_kernel -s 10 -c 1 -f -o $O/prof_r1h_k100 python bench.py --workload c3k100 --steps 20 --warmup 3 --no-cpu > $O/ncu_full_r1h_k100.log 2>&1
tail -2 $O/gpu_tests_r1h.log; for f in $O/bench_r1h*.json; do echo $f; cut -c1-400 $f; echo; done
