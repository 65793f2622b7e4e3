#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out
T=r2i
for f in test_gpu_r2 test_gpu_parity test_gpu_scale test_gpu_next test_gpu_alt test_gpu_stale test_gpu_sampler; do
  timeout 900 python -m pytest tests/$f.py -x -q --timeout 600 2>&1 | tail -25 > $O/${T}_$f.log
  echo "== $f"; tail -4 $O/${T}_$f.log
done
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
run jit c5 ""; run pre c5 "--no-specialize"
MDX_GVEC_CHUNK=100 run jit_nochunk c5 ""
MDX_JIT_BLOCKS=5 run jit_b5 c5 ""
run jit c4 ""; run jit c3k100 ""; run pre c3 ""
