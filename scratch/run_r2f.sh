#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out
T=r2f
run() { # name lib workload
  MDX_LIB_PATH=$2 timeout 200 python bench.py --workload $3 --steps 60 --warmup 5 --no-cpu --no-secondary --sustained-episodes 0 --no-feat-host $4 > $O/${T}_$1_$3.json 2> $O/${T}_$1_$3.err
  python - <<PY
import json
try:
    d=json.loads([l for l in open("$O/${T}_$1_$3.json").read().splitlines() if l.startswith("{")][-1])
    r=d["roofline"]; print("$1 $3 $4", "split_frac %.4f step_ms %.4f | single_frac %.4f launch_ms %.4f" % (r["frac"], r["step_ms"], r["single_launch"]["frac"], r["single_launch"]["mean_launch_ms"]))
except Exception as ex: print("$1 $3 failed", ex)
PY
}
for v in cur old5 old7 new6 new7 new8; do
  lib=scratch/libmdx_$v.so; [ $v = cur ] && lib=marl-delivery_b200/csrc/libmdx.so
  for wl in c4 c5; do run $v $lib $wl; done
done
for v in cur old5; do lib=scratch/libmdx_$v.so; [ $v = cur ] && lib=marl-delivery_b200/csrc/libmdx.so; run $v $lib c3k100; done
MDX_NOFAT=1 run cur_nofat marl-delivery_b200/csrc/libmdx.so c4
MDX_NOFAT=1 run cur_nofat marl-delivery_b200/csrc/libmdx.so c5
MDX_NOFAT=1 run old_nofat scratch/libmdx_old5.so c4
run cur marl-delivery_b200/csrc/libmdx.so c3
run fixh scratch/libmdx_fixh.so c3
# compute-sanitizer retry (VERDICT r1 task 9)
timeout 300 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_sampler.py -x -q -k "reference_reset_blocks" > $O/${T}_sanitizer.log 2>&1; tail -5 $O/${T}_sanitizer.log
