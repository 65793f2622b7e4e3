#!/bin/bash
O=gpurun_out
timeout 400 ncu --set full --clock-control none --import-source on -k regex:mdx_env_kernel -s 30 -c 1 -f -o $O/prof_r2u_c5 python bench.py --workload c5 --no-specialize --steps 40 --warmup 3 --no-cpu --no-secondary --split 1 --sustained-episodes 0 --no-feat-host > $O/r2u_ncu_c5.log 2>&1
echo rc=$?; ls -la $O/prof_r2u_c5.ncu-rep
