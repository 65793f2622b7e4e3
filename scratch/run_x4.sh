#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out
timeout 300 python -m pytest tests/test_gpu_next.py -x -q -k gae 2>&1 | tail -4
timeout 120 python profiles/bench_gae.py 2>&1 | tail -1
MDX_GAE_NO_TMA=1 timeout 120 python profiles/bench_gae.py 2>&1 | tail -1
for v in $(ls scratch/libmdx_gaetma_*.so 2>/dev/null); do echo -n "$v: "; MDX_LIB_PATH=$PWD/$v timeout 120 python profiles/bench_gae.py 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(round(d['ms_per_launch']*1e3,1),'us frac',round(d['roofline']['frac'],4))"; done
