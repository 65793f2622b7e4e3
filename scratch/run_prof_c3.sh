#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out
ncu --set full --clock-control none --import-source on -k regex:mdx_env_kernel -s 200 -c 1 -f -o $O/prof_r1i python bench.py --steps 220 --warmup 3 --no-cpu > $O/ncu_full_r1i.log 2>&1
ls -la $O/prof_r1i.ncu-rep
