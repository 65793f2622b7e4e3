#!/bin/bash
# driver-style final pass: smoke, the whole GPU suite, the driver's bench command, the reference arm
O=gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $O/r2x_smoke.log 2>&1; echo "smoke rc=$?"
timeout 1500 python -m pytest tests/ -x -q -m gpu > $O/r2x_tests.log 2>&1; echo "tests rc=$?"; tail -3 $O/r2x_tests.log
timeout 600 python bench.py --steps 20 --warmup 5 > $O/r2x_bench20.json 2> $O/r2x_bench20.err; echo "bench rc=$?"
timeout 300 python bench.py --impl reference --steps 20 --warmup 5 > $O/r2x_ref.json 2> $O/r2x_ref.err; echo "ref rc=$?"
python - <<PY
import json
d=json.loads(open("$O/r2x_bench20.json").read().strip().splitlines()[-1])
print("value", d["value"]/1e6, "frac", d["roofline"]["frac"], "e2e", d["e2e"]["value"]/1e6, "sustained", d.get("sustained",{}).get("value",0)/1e6, "clocks", d["clocks"])
for k,v in d.get("secondary_workloads",{}).items(): print(k, {a:(round(b,4) if isinstance(b,float) else b) for a,b in v.items() if a in ("value","frac","single_launch_frac","ms_per_step","kernel_build")})
r=json.loads(open("$O/r2x_ref.json").read().strip().splitlines()[-1]); print("ref", r["value"], r["cpu_baseline"])
PY
