#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out
T=r2l
for f in test_gpu_r2 test_gpu_parity test_gpu_next test_gpu_scale test_gpu_sampler test_gpu_stale test_gpu_alt; do
  timeout 900 python -m pytest tests/$f.py -x -q --timeout 600 2>&1 | tail -25 > $O/${T}_$f.log
  echo "== $f"; tail -3 $O/${T}_$f.log
done
for pdl in 1 0 2; do
MDX_PDL=$pdl timeout 200 python bench.py --workload c2 --steps 400 --warmup 20 --no-cpu --no-secondary --graph > $O/${T}_c2_pdl$pdl.json 2> $O/${T}_c2_pdl$pdl.err
python - <<PY
import json
try:
    d=json.loads([l for l in open("$O/${T}_c2_pdl$pdl.json").read().splitlines() if l.startswith("{")][-1])
    print("c2 pdl=$pdl ms/step %.5f value %.1f M graph %s" % (d["ms_per_step"], d["value"]/1e6, (d.get("cuda_graph") or {}).get("ms_per_step")))
except Exception as ex: print("failed", ex); print(open("$O/${T}_c2_pdl$pdl.err").read()[-600:])
PY
done
for pdl in 1 2; do
MDX_PDL=$pdl timeout 200 python bench.py --steps 60 --warmup 5 --no-cpu --no-secondary --sustained-episodes 0 --no-feat-host > $O/${T}_c3_pdl$pdl.json 2> $O/${T}_c3_pdl$pdl.err
python - <<PY
import json
d=json.loads([l for l in open("$O/${T}_c3_pdl$pdl.json").read().splitlines() if l.startswith("{")][-1])
r=d["roofline"]; print("c3 pdl=$pdl", "split_frac %.4f step_ms %.4f | single_frac %.4f launch_ms %.4f" % (r["frac"], r["step_ms"], r["single_launch"]["frac"], r["single_launch"]["mean_launch_ms"]), "single_stream", d["single_stream"]["ms_per_step"])
PY
done
