#!/bin/bash
# round-2 evidence: bench lines + ncu launch list of the default command + full captures of the dominant kernels -> gpurun_out/ (copied to profiles/ afterwards)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out
T=${TAG:-r2z}
timeout 500 python bench.py --steps 20 --warmup 5 > $O/${T}_bench20.json 2> $O/${T}_bench20.err
timeout 500 python bench.py > $O/${T}_bench_default.json 2> $O/${T}_bench_default.err
timeout 300 python bench.py --impl reference --steps 20 --warmup 5 > $O/${T}_ref.json 2> $O/${T}_ref.err
timeout 200 python profiles/dropin_latency.py > $O/${T}_dropin.json 2> $O/${T}_dropin.err
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/${T}_launches.csv python bench.py --steps 20 --warmup 5 --no-cpu --no-secondary --sustained-episodes 0 --no-feat-host > $O/${T}_ncu_launch.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:mdx_env_kernel -s 30 -c 1 -f -o $O/prof_${T}_c3 python bench.py --steps 20 --warmup 5 --no-cpu --no-secondary --sustained-episodes 0 --no-feat-host > $O/${T}_ncu_full_c3.log 2>&1
for wl in c5 c4 c3k100; do
  timeout 400 ncu --set full --clock-control none --import-source on -k regex:mdx_env_kernel -s 30 -c 1 -f -o $O/prof_${T}_$wl python bench.py --workload $wl --steps 40 --warmup 3 --no-cpu --no-secondary --split 1 --sustained-episodes 0 --no-feat-host > $O/${T}_ncu_full_$wl.log 2>&1
done
ls -la $O/prof_${T}_*.ncu-rep; head -c 600 $O/${T}_bench20.json; echo; tail -c 500 $O/${T}_bench20.err
