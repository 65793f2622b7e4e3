#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out
T=r2h
timeout 900 python -m pytest tests/test_gpu_r2.py -x -q --timeout 600 -k "specialised or split or guard" 2>&1 | tail -25 > $O/${T}_test_gpu_r2.log
tail -6 $O/${T}_test_gpu_r2.log
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
for wl in c4 c5 c3k100; do run jit $wl ""; run pre $wl "--no-specialize"; done
run jit c3 "--specialize"; run pre c3 ""
for blk in 5 6 8 9; do MDX_JIT_BLOCKS=$blk run jitb$blk c3k100 ""; done
for blk in 6 7; do MDX_JIT_BLOCKS=$blk run jitb$blk c4 ""; done
MDX_JIT_BLOCKS=4 run jitb4 c5 ""
