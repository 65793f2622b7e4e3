#!/bin/bash
# r2a: first GPU pass of round 2 (device sampler, packed rows, bench)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_sampler.py tests/test_gpu_r2.py -q --timeout 300 2>&1 | tail -40 > $O/r2a_new_tests.log
timeout 1200 python -m pytest tests -m gpu -q --timeout 300 --deselect tests/test_gpu_sampler.py --deselect tests/test_gpu_r2.py 2>&1 | tail -30 > $O/r2a_gpu_tests.log
timeout 400 python bench.py --steps 20 --warmup 5 > $O/r2a_bench20.json 2> $O/r2a_bench20.err
timeout 400 python bench.py --no-cpu --no-secondary > $O/r2a_bench.json 2> $O/r2a_bench.err
timeout 200 python profiles/dropin_latency.py > $O/r2a_dropin.json 2> $O/r2a_dropin.err
timeout 300 python bench.py --impl reference --steps 20 --warmup 5 > $O/r2a_ref.json 2> $O/r2a_ref.err
tail -5 $O/r2a_new_tests.log; tail -5 $O/r2a_gpu_tests.log; tail -c 1500 $O/r2a_bench20.err; cat $O/r2a_dropin.json; tail -c 600 $O/r2a_dropin.err
