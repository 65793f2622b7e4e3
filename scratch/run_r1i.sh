#!/bin/bash
# r1i measurement pass (run on the GPU box through gpurun)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > $O/gpu_tests_r1i.log
python bench.py --steps 1000 --warmup 20 > $O/bench_r1i.json 2> $O/bench_r1i.err
python bench.py --workload c2 --graph --steps 1000 --warmup 20 > $O/bench_r1i_c2.json 2> $O/bench_r1i_c2.err
python bench.py --workload c3k100 --steps 600 --warmup 10 > $O/bench_r1i_c3k100.json 2> $O/bench_r1i_c3k100.err
python bench.py --workload c4 --steps 600 --warmup 10 > $O/bench_r1i_c4.json 2> $O/bench_r1i_c4.err
python bench.py --workload c5 --steps 300 --warmup 10 --cpu-envs 1 --cpu-steps 100 > $O/bench_r1i_c5.json 2> $O/bench_r1i_c5.err
python bench.py --workload greedy --steps 1000 --warmup 20 > $O/bench_r1i_greedy.json 2> $O/bench_r1i_greedy.err
python bench.py --workload idq --steps 600 --warmup 10 > $O/bench_r1i_idq.json 2> $O/bench_r1i_idq.err
python profiles/bench_gae.py > $O/bench_r1i_gae.json 2> $O/bench_r1i_gae.err
python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_r1i_ref.json 2> $O/bench_r1i_ref.err
# ncu: launch list of the default bench command, then one full capture of the headline kernel
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $O/launches_r1i.csv python bench.py --steps 20 --warmup 3 --no-cpu > $O/ncu_launch_r1i.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:mdx_env_kernel -s 200 -c 1 -f -o $O/prof_r1i python bench.py --steps 220 --warmup 3 --no-cpu > $O/ncu_full_r1i.log 2>&1
tail -2 $O/gpu_tests_r1i.log; for f in $O/bench_r1i*.json; do echo $f; cut -c1-400 $f; echo; done
