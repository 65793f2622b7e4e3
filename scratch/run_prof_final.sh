#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out
timeout 200 ncu --set full --clock-control none --import-source on -k regex:mdx_env_kernel -s 60 -c 1 -f -o $O/prof_r1j_c5_final python bench.py --workload c5 --steps 70 --warmup 3 --no-cpu > $O/ncu_r1j_c5_final.log 2>&1
timeout 200 ncu --set full --clock-control none --import-source on -k regex:mdx_env_kernel -s 60 -c 1 -f -o $O/prof_r1j_c4_final python bench.py --workload c4 --steps 70 --warmup 3 --no-cpu > $O/ncu_r1j_c4_final.log 2>&1
ls -la $O/prof_r1j_c5_final.ncu-rep $O/prof_r1j_c4_final.ncu-rep
