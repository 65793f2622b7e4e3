#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out
for f in test_gpu_stale test_gpu_r2 test_gpu_sampler; do
  timeout 900 python -m pytest tests/$f.py -x -q --timeout 600 2>&1 | tail -40 > $O/r2c_$f.log
done
timeout 200 python profiles/dropin_latency.py > $O/r2c_dropin.json 2> $O/r2c_dropin.err
for f in test_gpu_stale test_gpu_r2 test_gpu_sampler; do echo "== $f"; tail -30 $O/r2c_$f.log; done
cat $O/r2c_dropin.json
