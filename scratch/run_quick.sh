#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -4 > $O/quick_tests.log
python bench.py --steps 1000 --warmup 20 --no-cpu > $O/quick_c3.json 2> $O/quick_c3.err
python profiles/bench_gae.py > $O/quick_gae.json 2>&1
cat $O/quick_tests.log; python - <<PY
import json
d=json.load(open("$O/quick_c3.json")); print("c3", round(d["value"]/1e6,2), "M frac", round(d["roofline"]["frac"],4), "e2e", round(d["e2e"]["value"]/1e6,1))
print(open("$O/quick_gae.json").read())
PY
