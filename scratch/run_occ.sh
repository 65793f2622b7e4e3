#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out
for v in base b8 b9 b10; do
  for c in -1 100; do
    if [ $v = base ]; then unset MDX_LIB_PATH; else export MDX_LIB_PATH=$PWD/scratch/libmdx_$v.so; fi
    if [ $c = -1 ]; then unset MDX_CARVEOUT; else export MDX_CARVEOUT=$c; fi
    python bench.py --steps 600 --warmup 20 --no-cpu > $O/occ_${v}_$c.json 2> $O/occ_${v}_$c.err
    python - <<PY
import json
try:
    d=json.load(open("$O/occ_${v}_$c.json")); print("$v carve=$c", round(d["value"]/1e6,1), "M  frac", round(d["roofline"]["frac"],4))
except Exception as ex: print("$v $c failed", ex)
PY
  done
done
