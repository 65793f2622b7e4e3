#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out
T=r2d
timeout 900 python -m pytest tests/test_gpu_r2.py -x -q --timeout 600 2>&1 | tail -30 > $O/${T}_test_gpu_r2.log
timeout 400 python bench.py --steps 20 --warmup 5 --no-cpu > $O/${T}_bench20.json 2> $O/${T}_bench20.err
timeout 400 python bench.py --steps 200 --warmup 20 --no-cpu --no-secondary > $O/${T}_bench200.json 2> $O/${T}_bench200.err
timeout 400 python bench.py --steps 200 --warmup 20 --no-cpu --no-secondary --split 4 > $O/${T}_bench200_s4.json 2> $O/${T}_bench200_s4.err
timeout 200 python profiles/dropin_latency.py > $O/${T}_dropin.json 2> $O/${T}_dropin.err
# ncu: full captures of the run-time-shape kernels and the headline kernel
for wl in c5 c4 c3k100 c3; do
  timeout 400 ncu --set full --clock-control none --import-source on -k regex:mdx_env_kernel -s 40 -c 1 -f -o $O/prof_${T}_$wl python bench.py --workload $wl --steps 50 --warmup 3 --no-cpu --no-secondary --split 1 --sustained-episodes 0 --no-feat-host > $O/ncu_${T}_$wl.log 2>&1
done
ls -la $O/prof_${T}_*.ncu-rep
tail -15 $O/${T}_test_gpu_r2.log; cat $O/${T}_dropin.json; tail -c 800 $O/${T}_bench20.err
