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
python -m pytest tests/test_gpu_alt.py tests/test_gpu_next.py -x -q 2>&1 | tail -4
python profiles/bench_gae.py 2>&1 | tail -1
python bench.py --workload idq --steps 600 --warmup 10 --no-cpu > $O/x1_idq.json 2> $O/x1_idq.err; show $O/x1_idq.json idq
for v in base t64b18 t96b12 t32b32 t128b10 t64b20; do
  MDX_LIB_PATH=$PWD/scratch/libmdx_$v.so python bench.py --steps 1000 --warmup 20 --no-cpu > $O/x1_c3_$v.json 2> $O/x1_c3_$v.err; show $O/x1_c3_$v.json "c3 $v"
done
for n in 63936 69264 58608; do
  python bench.py --envs-per-gpu $n --steps 1000 --warmup 20 --no-cpu > $O/x1_c3_n$n.json 2> $O/x1_c3_n$n.err; show $O/x1_c3_n$n.json "c3 N=$n"
done
