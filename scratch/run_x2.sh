#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out
show() { python - <<PY
import json
try:
    d=json.load(open("$1")); print("$2", round(d["value"]/1e6,2), "M  frac", round(d["roofline"]["frac"],4), "launch_us", round(d["roofline"]["mean_launch_ms"]*1e3,2), "e2e", round(d["e2e"]["value"]/1e6,1), "pipe", round(d["e2e"].get("pipelined_value",0)/1e6,1))
except Exception as ex: print("$2 failed", ex)
PY
}
python -m pytest tests -m gpu -x -q 2>&1 | tail -4
for v in gae_t128c16db0 gae_t128c16db1 gae_t32c16db0 gae_t256c16db1 gae_t128c32db0 gae_t128c8db1 gae_t64c16db1 gae_t256c16db0 gae_t512c8db1; do
  echo -n "$v: "; MDX_LIB_PATH=$PWD/scratch/libmdx_$v.so python profiles/bench_gae.py 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(round(d['ms_per_launch']*1e3,1),'us frac',round(d['roofline']['frac'],4))"
done
python bench.py --steps 1000 --warmup 20 --no-cpu > $O/x2_c3.json 2> $O/x2_c3.err; show $O/x2_c3.json c3
python bench.py --workload c5 --steps 300 --warmup 10 --no-cpu > $O/x2_c5_fat.json 2> $O/x2_c5_fat.err; show $O/x2_c5_fat.json "c5 fat stage-default"
MDX_STAGE_BYTES=3000 python bench.py --workload c5 --steps 300 --warmup 10 --no-cpu > $O/x2_c5_fat3000.json 2> $O/x2_c5_fat3000.err; show $O/x2_c5_fat3000.json "c5 fat stage3000"
MDX_NOFAT=1 python bench.py --workload c5 --steps 300 --warmup 10 --no-cpu > $O/x2_c5_nofat.json 2> $O/x2_c5_nofat.err; show $O/x2_c5_nofat.json "c5 nofat stage-default"
MDX_NOFAT=1 MDX_STAGE_BYTES=3000 python bench.py --workload c5 --steps 300 --warmup 10 --no-cpu > $O/x2_c5_nofat3000.json 2> $O/x2_c5_nofat3000.err; show $O/x2_c5_nofat3000.json "c5 nofat stage3000"
python bench.py --workload c4 --steps 600 --warmup 10 --no-cpu > $O/x2_c4.json 2> $O/x2_c4.err; show $O/x2_c4.json c4
