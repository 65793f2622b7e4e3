#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out
python -m pytest tests/test_gpu_parity.py tests/test_gpu_scale.py -x -q 2>&1 | tail -4 > $O/q3_tests.log
for wl in c3 c4 c5 c3k100; do
python bench.py --workload $wl --steps 600 --warmup 10 --no-cpu > $O/q3_$wl.json 2> $O/q3_$wl.err
done
cat $O/q3_tests.log; python - <<PY
import json
for wl in ("c3","c4","c5","c3k100"):
    try:
        d=json.load(open("$O/q3_%s.json"%wl)); print(wl, round(d["value"]/1e6,2), "M frac", round(d["roofline"]["frac"],4), "e2e", round(d["e2e"]["value"]/1e6,2))
    except Exception as ex: print(wl, "failed", ex, open("$O/q3_%s.err"%wl).read()[-800:])
PY
