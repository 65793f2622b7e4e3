#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out
# the whole GPU suite with EVERY instance on a run-time-compiled fused kernel (MDX_JIT=1)
export MDX_JIT=1
for f in test_gpu_parity test_gpu_scale test_gpu_next test_gpu_alt test_gpu_r2 test_gpu_sampler; do
  timeout 1500 python -m pytest tests/$f.py -q --timeout 900 2>&1 | tail -12 > $O/r2o_$f.log
  echo "== $f (MDX_JIT=1)"; tail -4 $O/r2o_$f.log
done
ls /tmp/mdx_jit_cache | wc -l
