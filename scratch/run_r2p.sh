#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_r2.py -x -q --timeout 600 2>&1 | tail -6
MDX_JIT=1 timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_gpu_next.py -q --timeout 900 2>&1 | tail -6
