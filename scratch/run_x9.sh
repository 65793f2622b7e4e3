#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out
show() { python - <<PY
import json
try:
    d=json.load(open("$1")); print("$2", round(d["value"]/1e6,2), "M  frac", round(d["roofline"]["frac"],4), "launch_us", round(d["roofline"]["mean_launch_ms"]*1e3,2), "e2e", round(d["e2e"]["value"]/1e6,1))
except Exception as ex: print("$2 failed", ex)
PY
}
python - <<PY
import torch
p=torch.cuda.get_device_properties(0)
print("L2", p.L2_cache_size, getattr(p,"persisting_l2_cache_max_size",None), getattr(p,"access_policy_max_window_size",None))
PY
timeout 200 python bench.py --steps 1000 --warmup 20 --no-cpu > $O/x9_c3.json 2> $O/x9_c3.err; show $O/x9_c3.json "c3 persist"
MDX_L2_PERSIST=0 timeout 200 python bench.py --steps 1000 --warmup 20 --no-cpu > $O/x9_c3_np.json 2> $O/x9_c3_np.err; show $O/x9_c3_np.json "c3 no-persist"
MDX_L2_PERSIST_MB=64 timeout 200 python bench.py --steps 1000 --warmup 20 --no-cpu > $O/x9_c3_64.json 2> $O/x9_c3_64.err; show $O/x9_c3_64.json "c3 persist 64MB"
timeout 200 python bench.py --steps 1000 --warmup 20 --no-cpu > $O/x9_c3b.json 2> $O/x9_c3b.err; show $O/x9_c3b.json "c3 persist (2)"
timeout 120 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct --clock-control none -k regex:mdx_env_kernel -s 100 -c 2 python bench.py --steps 110 --warmup 3 --no-cpu 2>&1 | grep -E "dram__|gpu__time|lts__" 
for wl in c5 c4 c3k100 idq; do
timeout 200 python bench.py --workload $wl --steps 300 --warmup 10 --no-cpu > $O/x9_$wl.json 2> $O/x9_$wl.err; show $O/x9_$wl.json "$wl persist"
MDX_L2_PERSIST=0 timeout 200 python bench.py --workload $wl --steps 300 --warmup 10 --no-cpu > $O/x9_${wl}_np.json 2> $O/x9_${wl}_np.err; show $O/x9_${wl}_np.json "$wl no-persist"
done
