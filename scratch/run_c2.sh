#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
timeout 100 python -m pytest tests/test_gpu_parity.py tests/test_gpu_next.py -x -q -k "kat_a or kat_b or host_pipeline or collect_rollouts or checkpoint" 2>&1 | tail -2
timeout 100 python bench.py --workload c2 --graph --steps 2000 --warmup 50 --no-cpu | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('c2', round(d['value']/1e6,1), 'M', round(d['ms_per_step']*1e3,2), 'us/launch; graph', round(d['cuda_graph']['value']/1e6,1), 'e2e', round(d['e2e']['value']/1e6,1))"
