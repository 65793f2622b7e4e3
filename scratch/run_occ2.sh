#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out
show() { python - <<PY
import json
try:
    d=json.load(open("$1")); print("$2", round(d["value"]/1e6,2), "M  frac", round(d["roofline"]["frac"],4))
except Exception as ex: print("$2 failed", ex)
PY
}
export MDX_CARVEOUT=100
for v in g7_f10 g7_f11 g7_f12; do
  MDX_LIB_PATH=$PWD/scratch/libmdx_$v.so python bench.py --steps 600 --warmup 20 --no-cpu > $O/occ2_c3_$v.json 2> $O/occ2_c3_$v.err; show $O/occ2_c3_$v.json "c3 $v carve100"
done
unset MDX_CARVEOUT
python bench.py --steps 600 --warmup 20 --no-cpu > $O/occ2_c3_base.json 2>$O/occ2_c3_base.err; show $O/occ2_c3_base.json "c3 base(g7_f9)"
for wl in c4 c5 c3k100 c2; do
  python bench.py --workload $wl --steps 300 --warmup 10 --no-cpu > $O/occ2_${wl}_base.json 2>$O/occ2_${wl}_base.err; show $O/occ2_${wl}_base.json "$wl g7"
  for v in g8_f9 g9_f9 g10_f9; do
    MDX_LIB_PATH=$PWD/scratch/libmdx_$v.so python bench.py --workload $wl --steps 300 --warmup 10 --no-cpu > $O/occ2_${wl}_$v.json 2> $O/occ2_${wl}_$v.err; show $O/occ2_${wl}_$v.json "$wl $v"
  done
done
