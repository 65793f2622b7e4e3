#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out
T=r2m
timeout 900 python -m pytest tests/test_gpu_next.py tests/test_gpu_r2.py -x -q --timeout 600 2>&1 | tail -8
for n in 4096 32768; do for p in 1 2; do echo "envs $n parts $p"; timeout 300 python examples/mappo_rollout_demo.py --envs $n --steps 32 --iters 4 --parts $p 2>&1 | tail -2; done; done
python scratch/c2_launch.py
