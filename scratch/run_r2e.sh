#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out
T=r2e
for f in test_gpu_parity test_gpu_scale test_gpu_next test_gpu_alt test_gpu_stale test_gpu_r2 test_gpu_sampler; do
  timeout 900 python -m pytest tests/$f.py -x -q --timeout 600 2>&1 | tail -25 > $O/${T}_$f.log
done
timeout 500 python bench.py --steps 20 --warmup 5 --no-cpu > $O/${T}_bench20.json 2> $O/${T}_bench20.err
for wl in c5 c4; do
  timeout 400 ncu --set full --clock-control none --import-source on -k regex:mdx_env_kernel -s 40 -c 1 -f -o $O/prof_${T}_$wl python bench.py --workload $wl --steps 50 --warmup 3 --no-cpu --no-secondary --split 1 --sustained-episodes 0 --no-feat-host > $O/ncu_${T}_$wl.log 2>&1
done
for f in test_gpu_parity test_gpu_scale test_gpu_next test_gpu_alt test_gpu_stale test_gpu_r2 test_gpu_sampler; do echo "== $f"; tail -4 $O/${T}_$f.log; done
tail -c 800 $O/${T}_bench20.err
