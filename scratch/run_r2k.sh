#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out
T=r2k
timeout 900 python -m pytest tests/test_gpu_r2.py -x -q --timeout 600 2>&1 | tail -25 > $O/${T}_test_gpu_r2.log; tail -5 $O/${T}_test_gpu_r2.log
run() { # tag workload extra
  timeout 200 python bench.py --workload $2 --steps 60 --warmup 5 --no-cpu --no-secondary --sustained-episodes 0 --no-feat-host $3 > $O/${T}_$1_$2.json 2> $O/${T}_$1_$2.err
  python - <<PY
import json
try:
    d=json.loads([l for l in open("$O/${T}_$1_$2.json").read().splitlines() if l.startswith("{")][-1])
    r=d["roofline"]; print("$1 $2", "split_frac %.4f step_ms %.4f | single_frac %.4f launch_ms %.4f |" % (r["frac"], r["step_ms"], r["single_launch"]["frac"], r["single_launch"]["mean_launch_ms"]), d["config"]["kernel_build"])
except Exception as ex: print("$1 $2 failed", ex); print(open("$O/${T}_$1_$2.err").read()[-800:])
PY
}
run baked c3 ""
MDX_JIT_FORCE=1 run jit_interleave c3 "--specialize"
MDX_JIT_FORCE=1 MDX_JIT_NO_INTERLEAVE=1 run jit_plain c3 "--specialize"
MDX_JIT_FORCE=1 MDX_JIT_BLOCKS=8 run jit_interleave_b8 c3 "--specialize"
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu > $O/${T}_bench20.json 2> $O/${T}_bench20.err
python - <<PY
import json
d=json.loads([l for l in open("$O/${T}_bench20.json").read().splitlines() if l.startswith("{")][-1])
print("headline", d["value"]/1e6, d["roofline"]["frac"], "e2e", d["e2e"]["value"]/1e6)
for k,v in d["secondary_workloads"].items():
    print(k, v.get("value",0)/1e6, v.get("ms_per_step"), (v.get("roofline") or {}).get("frac"), v.get("kernel_build"), {x:(v.get(x) or {}).get("ms_per_step") for x in ("stepper","step_n","cuda_graph")}, v.get("unavailable"))
PY
