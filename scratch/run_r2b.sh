#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out
timeout 1200 python -m pytest tests/test_gpu_stale.py tests/test_gpu_r2.py tests/test_gpu_sampler.py -q --timeout 600 2>&1 | tail -60 > $O/r2b_new_tests.log
timeout 1200 python -m pytest tests -m gpu -q --timeout 300 --deselect tests/test_gpu_sampler.py --deselect tests/test_gpu_r2.py --deselect tests/test_gpu_stale.py 2>&1 | tail -30 > $O/r2b_gpu_tests.log
timeout 200 python profiles/dropin_latency.py > $O/r2b_dropin.json 2> $O/r2b_dropin.err
timeout 400 python bench.py --steps 20 --warmup 5 --no-cpu > $O/r2b_bench20.json 2> $O/r2b_bench20.err
tail -25 $O/r2b_new_tests.log; tail -5 $O/r2b_gpu_tests.log; tail -c 1500 $O/r2b_bench20.err; cat $O/r2b_dropin.json; tail -c 600 $O/r2b_dropin.err
